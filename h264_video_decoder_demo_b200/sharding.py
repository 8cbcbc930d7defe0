"""Multi-GPU partitioning: independent bitstreams (or closed GOPs) are dealt round-robin to ranks.

The path shards with NO data-path collective: every rank owns its streams' decoded picture
buffers; the only cross-rank traffic is the 8-byte frame checksums and the timing max, carried by
torch.distributed (NCCL on the GPU box, gloo in the CPU tests)."""
from typing import Dict, List


def shard(n_units: int, rank: int, world: int) -> List[int]:
    """Unit i (stream or GOP index) -> rank i % world."""
    return list(range(rank, n_units, world))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_checksums(local: Dict[int, int], n_units: int) -> List[int]:
    """All ranks contribute {unit: u64 checksum}; returns the full list on every rank."""
    import torch
    d = _dist()
    # two int32 halves per checksum: exact under SUM on every backend
    t = torch.zeros(n_units, 2, dtype=torch.int64, device=_device())
    for u, s in local.items():
        t[u, 0] = s & 0xFFFFFFFF
        t[u, 1] = (s >> 32) & 0xFFFFFFFF
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    t = t.cpu()
    return [int(t[u, 0]) | (int(t[u, 1]) << 32) for u in range(n_units)]


def max_over_ranks(value: float) -> float:
    import torch
    d = _dist()
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    d = _dist()
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def barrier():
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.barrier()


def closed_gop_starts(rp) -> List[int]:
    """Decode indices at which a closed GOP starts (an IDR picture: the DPB is wiped, RPL:1506-1518, the output
    queue is flushed, VD:389-414, POC restarts, RPL:49-53), so that [start_k, start_k+1) can be reconstructed
    independently of everything before it.  The replay container does not carry nal_unit_type; an IDR is
    recognised as an I picture whose PicOrderCnt is 0 and which no later picture precedes in output order."""
    starts = []
    # output position of every decode index; a picture that is never output (truncated stream) sorts last
    pos = {d: k for k, d in enumerate(getattr(rp, "out_order", []) or [])}
    n = len(rp.pictures)
    big = len(pos) + n
    opos = [pos.get(p.decode_idx, big + i) for i, p in enumerate(rp.pictures)]
    # suffix minimum of the output positions: picture i opens a closed GOP only if everything decoded before it is output before
    # everything decoded from it on (a non-IDR I picture with POC 0, e.g. after MMCO5, fails this when B pictures straddle it)
    suf = [big + n] * (n + 1)
    for i in range(n - 1, -1, -1):
        suf[i] = min(suf[i + 1], opos[i])
    pre = -1
    for i, p in enumerate(rp.pictures):
        if p.slice_type % 5 == 2 and p.poc == 0 and p.has_inter == 0 and pre < suf[i]:
            starts.append(i)
        pre = max(pre, opos[i])
    if not starts or starts[0] != 0:
        starts.insert(0, 0)
    return starts
