#!/usr/bin/env python
"""Concurrent host<->device copy probe: what is THIS box's aggregate pinned-memory ceiling?

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py

Every rank owns one GPU and one pinned host buffer (cudaHostAlloc through torch). For k = 1, 2, 4, .., world the
first k ranks copy 1 GiB device->host (then host->device) at the same time, `reps` times back to back; the other
ranks idle. Per k the line gives the per-rank GB/s (CUDA events on the copying stream) and the aggregate (total
bytes / the slowest rank's time). Variants: default pinned; pinned after MADV_HUGEPAGE + cudaHostRegister (fewer
IOMMU / TLB entries per byte); a kernel writing straight into mapped host memory (no copy engine).
Rank 0 then drives ALL GPUs from one process (one stream per device) — the one-process-8-threads question.
Output: one JSON line per measurement on rank 0's stdout; bench.py reads the committed summary
(profiles/r2_d2h_probe.json) as the denominator of its multi-GPU `e2e` fraction."""
import ctypes
import json
import mmap
import os
import sys
import time

import torch
import torch.distributed as dist

GIB = 1 << 30


def pinned_plain(nbytes):
    return torch.empty(nbytes, dtype=torch.uint8).pin_memory()


def pinned_huge(nbytes):
    """anonymous mmap + MADV_HUGEPAGE, touched, then cudaHostRegister: 2 MiB pages where THP allows."""
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except Exception:
        pass
    t = torch.frombuffer(mm, dtype=torch.uint8)
    t.zero_()
    rt = torch.cuda.cudart()
    rc = rt.cudaHostRegister(t.data_ptr(), nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError("cudaHostRegister rc=%s" % rc)
    t._mm = mm
    return t


def timed_copies(dst, src, reps, stream):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        ev1.record()
    ev1.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    nbytes, reps = GIB, 4
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.zero_()
    st = torch.cuda.Stream(device=dev)
    out = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def sweep(tag, h):
        ks = [k for k in (1, 2, 4, 8, 16) if k <= world]
        for direction in ("d2h", "h2d"):
            for k in ks:
                timed_copies(h, d, 1, st) if direction == "d2h" else timed_copies(d, h, 1, st)
                barrier()
                secs = 0.0
                if rank < k:
                    secs = timed_copies(h, d, reps, st) if direction == "d2h" else timed_copies(d, h, reps, st)
                barrier()
                t = torch.tensor([secs], dtype=torch.float64)
                if world > 1:
                    allt = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
                    dist.all_gather(allt, t)
                    secs_all = [float(x) for x in allt][:k]
                else:
                    secs_all = [secs]
                if rank == 0:
                    per = [reps * nbytes / s / 1e9 for s in secs_all]
                    out.append({"probe": tag, "dir": direction, "ranks_active": k, "per_rank_GBps": [round(x, 1) for x in per],
                                "aggregate_GBps": round(k * reps * nbytes / max(secs_all) / 1e9, 1)})
                    print(json.dumps(out[-1]), flush=True)

    h = pinned_plain(nbytes)
    sweep("cudaHostAlloc (torch pin_memory), one process per GPU", h)
    del h
    try:
        h2 = pinned_huge(nbytes)
        sweep("mmap + MADV_HUGEPAGE + cudaHostRegister, one process per GPU", h2)
        torch.cuda.cudart().cudaHostUnregister(h2.data_ptr())
        del h2
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"probe": "hugepage", "error": str(e)}), flush=True)
    barrier()
    # one process, all GPUs (rank 0 only; the others wait)
    if rank == 0:
        ng = torch.cuda.device_count()
        bufs = []
        for g in range(ng):
            with torch.cuda.device(g):
                bufs.append((torch.zeros(nbytes, dtype=torch.uint8, device="cuda:%d" % g), pinned_plain(nbytes),
                             torch.cuda.Stream(device="cuda:%d" % g)))
        for k in [k for k in (1, 2, 4, 8) if k <= ng]:
            for g in range(k):
                with torch.cuda.device(g):
                    torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                for g in range(k):
                    dd, hh, ss = bufs[g]
                    with torch.cuda.device(g), torch.cuda.stream(ss):
                        hh.copy_(dd, non_blocking=True)
            for g in range(k):
                with torch.cuda.device(g):
                    torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out.append({"probe": "one process driving k GPUs (wall clock)", "dir": "d2h", "ranks_active": k,
                        "aggregate_GBps": round(k * reps * nbytes / dt / 1e9, 1)})
            print(json.dumps(out[-1]), flush=True)
        try:
            info = {"cpus": os.cpu_count(), "numa_nodes": sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")),
                    "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()}
        except Exception as e:  # noqa: BLE001
            info = {"error": str(e)}
        print(json.dumps({"probe": "host", **info}), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
