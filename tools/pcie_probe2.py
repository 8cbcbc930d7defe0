"""Picture-batch shaped PCIe probe: k pictures per batch, H2D chunk size hc per picture and D2H chunk 3.13 MB per picture,
each direction on its own stream; per-direction time from CUDA events.  merged = the whole batch as ONE transfer."""
import sys, torch
k, reps = 128, 6
D = 3133440
def probe(hc, merge_h, merge_d, label):
    hs = torch.empty(k * hc, dtype=torch.uint8, pin_memory=True); ds = torch.empty(k * hc, dtype=torch.uint8, device="cuda")
    hd = torch.empty(k * D, dtype=torch.uint8, pin_memory=True); dd = torch.empty(k * D, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    def batch():
        with torch.cuda.stream(s1):
            if merge_h: ds.copy_(hs, non_blocking=True)
            else:
                for i in range(k): ds[i*hc:(i+1)*hc].copy_(hs[i*hc:(i+1)*hc], non_blocking=True)
        with torch.cuda.stream(s2):
            if merge_d: hd.copy_(dd, non_blocking=True)
            else:
                for i in range(k): hd[i*D:(i+1)*D].copy_(dd[i*D:(i+1)*D], non_blocking=True)
    batch(); torch.cuda.synchronize()
    ev[0].record(s1); ev[2].record(s2)
    for _ in range(reps): batch()
    ev[1].record(s1); ev[3].record(s2)
    torch.cuda.synchronize()
    th, td = ev[0].elapsed_time(ev[1]) / reps, ev[2].elapsed_time(ev[3]) / reps
    print(f"{label}: H2D {k*hc/1e6:.0f} MB in {th:.2f} ms = {k*hc/th/1e6:.1f} GB/s | D2H {k*D/1e6:.0f} MB in {td:.2f} ms = {k*D/td/1e6:.1f} GB/s | batch rate {1000/max(th,td)*k:.0f} pictures/s")
for hc, name in ((3400000, "dense 3.4MB"), (1650000, "packed 1.65MB"), (450000, "0.45MB")):
    probe(hc, False, False, f"H2D {name} per picture, D2H per picture")
    probe(hc, True, False, f"H2D {name} merged per batch, D2H per picture")
    probe(hc, True, True, f"H2D {name} merged, D2H merged")
