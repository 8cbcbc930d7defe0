"""Measure what the box's PCIe link gives (pinned host memory), alone and in both directions at once."""
import time
import torch
n = 512 << 20
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.time() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def chunks(k):
    def f():
        c = n // k
        with torch.cuda.stream(s1):
            for i in range(k): d1[i*c:(i+1)*c].copy_(h1[i*c:(i+1)*c], non_blocking=True)
    return f
print("H2D GB/s", n / run(h2d) / 1e9)
print("D2H GB/s", n / run(d2h) / 1e9)
t = run(both); print("both: each direction GB/s", n / t / 1e9)
print("H2D in 3.4MB chunks GB/s", n / run(chunks(150)) / 1e9)
def chunks_d2h(k):
    def f():
        c = n // k
        with torch.cuda.stream(s2):
            for i in range(k): h2[i*c:(i+1)*c].copy_(d2[i*c:(i+1)*c], non_blocking=True)
    return f
def both_chunks(k):
    a, b = chunks(k), chunks_d2h(k)
    def f():
        a(); b()
    return f
t = run(both_chunks(150)); print("both directions in 3.4MB chunks: each GB/s", n / t / 1e9)
t = run(both_chunks(16)); print("both directions in 32MB chunks: each GB/s", n / t / 1e9)
