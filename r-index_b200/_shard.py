"""Multi-GPU plumbing for the pattern-parallel path: the index is replicated on
every GPU, patterns are cut into contiguous shards, and there is NO collective on the search path.
The optional collation below (all-gather of the fixed-size range records, all-gather-v of the
occurrence buffers) runs over torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU
tests. Same shard formula as the C++ fan-out (host/cli_common.hpp: GpuFleet)."""
import numpy as np


def shard_bounds(N, world, rank):
    """Contiguous shard [a, b) of rank `rank`: concatenating shard outputs in rank order reproduces
    the single-GPU output order."""
    return N * rank // world, N * (rank + 1) // world


def balanced_cuts(nocc, world, per_pattern_cost=64):
    """SURVEY.md §8e re-balancing: cut the patterns [0, N) into `world` CONTIGUOUS shards of near-equal work,
    work(p) = n_occ(p) + per_pattern_cost (the backward search of a pattern costs about as much as a few dozen
    occurrences of expansion). nocc: per-pattern occurrence counts of the WHOLE batch (after the count phase and
    the all-gather of the counts: 8 bytes per pattern). Returns world + 1 ascending cut points, cuts[0] = 0,
    cuts[world] = N; rank k takes [cuts[k], cuts[k+1]). Contiguity keeps the output order = concatenation of the
    shard outputs. Integer arithmetic only — the same rule, bit for bit, as the device kernel
    (csrc/post_kernels.cuh: balanced_cuts_kernel) and the C++ fan-out (host/cli_common.hpp):
      i = the first pattern with cum(i) * world >= total * k;  c = i + 1;
      if cum(i) * world - total * k > total * k - cum(i-1) * world: c = i      (the crossing pattern goes to the side
      that leaves the smaller excess);  cuts[k] = clamp(c, cuts[k-1], N)."""
    w = np.asarray(nocc, dtype=np.uint64) + np.uint64(per_pattern_cost)
    N = int(w.size)
    cum = np.cumsum(w, dtype=np.uint64)
    total = int(cum[-1]) if N else 0
    cuts = [0]
    for k in range(1, world):
        c = N
        if N and total:
            t = total * k
            i = int(np.searchsorted(cum, np.uint64(-(-t // world)), side="left"))   # first i with cum[i] >= ceil(t / world)
            if i < N:
                ci, cp = int(cum[i]), (int(cum[i - 1]) if i else 0)
                c = i if (ci * world - t) > (t - cp * world) else i + 1
        cuts.append(min(max(c, cuts[-1]), N))
    cuts.append(N)
    return cuts


def balanced_cuts_from_offsets(off, world, per_pattern_cost=64):
    """The same cut points from the EXCLUSIVE PREFIX of the occurrence counts (off[0..N], off[N] = total) that a
    locate-mode search leaves behind — the arithmetic of csrc/post_kernels.cuh: cuts_from_offsets_kernel
    (rig_plan_batch_dev), restated for the CPU tests: cum(i) = off[i + 1] + cost * (i + 1) is monotone, so every
    target is a binary search instead of a pass over the batch."""
    off = [int(x) for x in off]
    N = len(off) - 1
    total = off[N] + per_pattern_cost * N if N else 0
    raw = []
    for k in range(world + 1):
        c = N
        if k == 0:
            c = 0
        elif k < world and N and total:
            t = total * k
            lo, hi = 0, N
            while lo < hi:
                mid = (lo + hi) >> 1
                if (off[mid + 1] + per_pattern_cost * (mid + 1)) * world >= t:
                    hi = mid
                else:
                    lo = mid + 1
            i = lo
            if i < N:
                cum, prev = off[i + 1] + per_pattern_cost * (i + 1), off[i] + per_pattern_cost * i
                c = i if (cum * world - t) > (t - prev * world) else i + 1
        raw.append(c)
    cuts, last = [], 0
    for k, c in enumerate(raw):
        c = min(max(c, last), N)
        if k == world:
            c = N
        cuts.append(c)
        last = c
    return cuts


def _torch():
    import torch
    import torch.distributed as dist
    return torch, dist


def collate_ranges(lo, hi, N, device=None):
    """all-gather of the per-pattern (lo, hi) records. lo/hi: this rank's shard (numpy u64).
    Returns full-length numpy arrays on every rank."""
    torch, dist = _torch()
    world, rank = dist.get_world_size(), dist.get_rank()
    per = max(shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world))
    buf = torch.zeros(2 * per, dtype=torch.int64, device=device)
    a, b = shard_bounds(N, world, rank)
    buf[: b - a] = torch.from_numpy(lo.view(np.int64)).to(buf.device)
    buf[per: per + b - a] = torch.from_numpy(hi.view(np.int64)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    glo = np.zeros(N, dtype=np.uint64)
    ghi = np.zeros(N, dtype=np.uint64)
    for r in range(world):
        ra, rb = shard_bounds(N, world, r)
        t = out[r].cpu().numpy().view(np.uint64)
        glo[ra:rb] = t[: rb - ra]
        ghi[ra:rb] = t[per: per + rb - ra]
    return glo, ghi


def collate_occurrences(off_local, occ_local, device=None):
    """all-gather-v of the occurrence buffers (padded all_gather). off_local: shard-local offsets
    (shard patterns + 1), occ_local: this shard's occurrences. Returns (global offsets, global occ)."""
    torch, dist = _torch()
    world = dist.get_world_size()
    sizes = torch.tensor([occ_local.size, off_local.size - 1], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [tuple(int(x) for x in s.cpu().tolist()) for s in all_sizes]
    max_occ = max(s[0] for s in all_sizes)
    max_pat = max(s[1] for s in all_sizes)
    buf = torch.zeros(max_occ + max_pat + 1, dtype=torch.int64, device=device)
    buf[: occ_local.size] = torch.from_numpy(occ_local.view(np.int64)).to(buf.device)
    buf[max_occ: max_occ + off_local.size] = torch.from_numpy(off_local.view(np.int64)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    occs, offs, base = [], [np.zeros(1, dtype=np.uint64)], 0
    for r in range(world):
        t = out[r].cpu().numpy().view(np.uint64)
        n_occ, n_pat = all_sizes[r]
        occs.append(t[:n_occ])
        offs.append(t[max_occ + 1: max_occ + 1 + n_pat] + np.uint64(base))
        base += n_occ
    return np.concatenate(offs), np.concatenate(occs)
