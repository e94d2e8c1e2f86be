#!/usr/bin/env python
"""bench.py — the hot path's headline measurement (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5|...]

A "step" = one pass of the locate hot path (backward search + toehold, offsets, Phi expansion) over ONE JOB = one
batch of synthetic patterns.

  N = 1   default workload = BASELINE.json configs[1] (C2): ri-locate, 100 MB synthetic repetitive DNA (sigma=4),
          100k patterns of length 20.
  N > 1   default workload = BASELINE.json configs[4] (C5): ri-locate, 4 GB synthetic DNA, 100k patterns of length 15,
          ~10k occurrences each (~1e9 occurrences per job) — the config BASELINE names for 1 -> 8 GPU scaling. STRONG
          scaling: the job is FIXED, the index replicated in every GPU's HBM, the patterns sharded. Per step every
          rank (1) learns the per-pattern occurrence counts — `--plan replicate`: it counts the whole batch itself (no
          collective; the automatic choice up to 8 M LF steps per batch, a count pass then costs less than a
          collective's latency); `--plan gather`: it counts its equal-count shard and the ranks all-gather the counts
          (8 B per pattern: NCCL for the device-resident number, gloo for the host-buffer one) —, (2) cuts the batch
          into contiguous shards of equal OCCURRENCE mass (the same integer rule on every rank), (3)
          locates its shard. No collective on the search path. In the same run rank 0 also runs the WHOLE job alone (`strong_scaling.n1`), so every line
          carries the one-GPU figure of its own job.

  value      occurrences/s of the whole job, device-resident inputs and outputs, CUDA-event timed, max over ranks
  e2e        occurrences/s through the host-buffer C-ABI calls (rig_count_batch / rig_locate_batch) with pinned host
             buffers: H2D of the patterns and D2H of ranges, offsets and every occurrence inside the timed region
  roofline   dominant kernel (phi_fused_kernel: seed hops + window fill in one persistent kernel): bytes that must
             cross HBM per launch (8 B per occurrence written + one pass over the Phi tables + per item 32 B of list
             traffic and one 64 B seed record) / its CUDA-event duration, against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json); `traffic` = ncu DRAM bytes of the same (workload, kernel) from
             profiles/ncu_traffic.json, null when no capture of that pair exists
  cpu_baseline  the reference's own code (oracle/_ref) on the box's host cores, bounded sample (N = 1 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

CACHE = os.path.join(ROOT, ".cache")

# SURVEY.md §8d configs, concretised. snps: tuned so that r ~ n/1000 for C2 (measured: see DESIGN.md).
WORKLOADS = {
    # name: (kind, n, p0, p1, text_seed, N, m, patt_seed, start_limit, description)
    "c2": ("dna_drift", 100_000_000, 50_000, 3, 0xB2000002, 100_000, 20, 0xB2001002, 0,
           "C2 ri-locate: 100MB synthetic DNA sigma=4 repetitive, 100k len-20 patterns"),
    "c2x4": ("dna_drift", 100_000_000, 50_000, 3, 0xB2000002, 400_000, 20, 0xB2001002, 0,
             "C2 text with a 4x larger batch (400k len-20 patterns): parallelism sensitivity"),
    "c2s": ("dna_drift", 10_000_000, 50_000, 3, 0xB2000002, 20_000, 20, 0xB2001002, 0,
            "C2 scaled down 10x (smoke/dev)"),
    "c5s": ("dna_indep", 400_000_000, 400_000, 10_000, 0xB2000005, 100_000, 15, 0xB2001005, 400_000,
            "C5 scaled to 400MB (1000 copies): Phi-chain stress, ~1k occ/pattern"),
    "c3s": ("versioned_doc", 200_000_000, 25_000, 96, 0xB2000003, 200_000, 30, 0xB2001003, 0,
            "C3 scaled to 200MB: einstein-like sigma=96 versioned document, 200k len-30 patterns"),
    # edit probability per version 0.05 instead of SURVEY 8d's 0.25: 0.25 gives r = 165k at 40k versions, outside
    # the config's r ~ 50k (measured: r = 21.7k + 14.3 per edit)
    "c3": ("versioned_doc", 1_000_000_000, 25_000, 50_096, 0xB2000003, 125_000, 30, 0xB2001003, 0,
           "C3 ri-locate at FULL size: 1 GB einstein-like sigma=96 versioned document; 125k len-30 patterns per GPU (the config's 1M patterns on 8 GPUs)"),
    "c5": ("dna_indep", 4_000_000_000, 400_000, 10_000, 0xB2000005, 100_000, 15, 0xB2001005, 400_000,
           "C5 ri-locate at FULL size: 4 GB synthetic DNA sigma=4, 10k copies, ~10k occ/pattern, 100k len-15 patterns (n = 4.0e9, just under 2^32: 32-bit words)"),
    "c4": ("pangenome", 10_000_000_000, 10_000_000, 100_000, 0xB2000004, 1_250_000, 100, 0xB2001004, 0,
           "C4 ri-count at FULL size: 10 GB synthetic pan-genome sigma=5 (1000 haplotypes); 1.25M len-100 reads per GPU (the config's 10M reads on 8 GPUs)"),
    "c4s": ("pangenome", 1_000_000_000, 10_000_000, 100_000, 0xB2000004, 1_000_000, 100, 0xB2001004, 0,
            "C4 scaled to 1GB: synthetic pan-genome sigma=5 (100 haplotypes), 1M len-100 reads (count; index >> L2)"),
}
# How a workload's JOB grows with the number of GPUs: "strong" = the job is WORKLOADS' N patterns whatever the GPU
# count (C5: BASELINE's scaling config; C2); "weak" = N patterns per GPU, one job of N x gpus patterns (C3 and C4 are
# defined on 8 GPUs: 125k x 8 = the config's 1M patterns, 1.25M x 8 = its 10M reads; one GPU cannot hold C3's output).
STRONG = {"c2", "c2s", "c2x4", "c5", "c5s"}
B_PHI = 264          # SURVEY §8d's touched-bytes figure per occurrence for the REFERENCE structure (4 x 64 B blocks + 8 B store); printed as survey_touched, not used for `roofline.frac`
B_RANK = lambda ell: 64 * (3 + ell)  # noqa: E731  per rank query
FALLBACK_HBM_GBS = 6650.0
L2_BYTES = 126 << 20
CONFIG_NOTE = ("GPU arm: 256 MiB memset between timed steps (L2 flush, outside the per-step event pair), CUDA events on the "
               "launch stream, max over ranks; CPU arm: steady_clock around the query loop")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def job_size(workload, world):
    N = WORKLOADS[workload][5]
    return N if workload in STRONG else N * world


def base_of(workload):
    return {"c2x4": "c2"}.get(workload, workload)  # workloads that share a text share its index


def shared_config(workload, world, n, r):
    """`config` of the JSON line: identical keys and values in both arms (ours / --impl reference)."""
    kind, _, _, _, _, _, m, _, _, desc = WORKLOADS[workload]
    return {"workload": desc, "name": workload, "n": int(n), "r": int(r), "job_patterns": job_size(workload, world),
            "pattern_length": m, "gpus": world,
            "scaling": ("strong: one fixed job" if workload in STRONG else "weak: %d patterns per GPU, one job" % WORKLOADS[workload][5]),
            "notes": CONFIG_NOTE}


def job_inputs(workload, world, rank, need_text=False, barrier=None, need_index=True):
    """The job's pattern array (the same on every rank) and this repo's logical index. Rank 0 builds what .cache/
    lacks (text -> prefix-free-parsing builder; patterns drawn from the text) and the others load it after the
    barrier: one build per box, not one per rank. Returns (text or None, patt, N, m, host). need_index=False (the
    reference arm with its own cached index): nothing of this repo is loaded when the cached files exist."""
    rib = None
    kind, n, p0, p1, tseed, _, m, pseed, limit, desc = WORKLOADS[workload]
    N = job_size(workload, world)
    os.makedirs(CACHE, exist_ok=True)
    rpath = os.path.join(CACHE, "%s.rib" % base_of(workload))
    ppath = os.path.join(CACHE, "%s.N%d.m%d.patt.npy" % (workload, N, m))
    text = None
    t0 = time.time()
    if rank == 0:
        if need_text or not (os.path.exists(rpath) and os.path.exists(ppath)):
            rib = ge.load_package()
            text = rib.gen_text(kind, n, p0, p1, tseed)
        if not os.path.exists(ppath):
            patt = rib.gen_patterns(text, N, m, pseed, limit)
            np.save(ppath + ".tmp.npy", patt)
            os.replace(ppath + ".tmp.npy", ppath)
        if not os.path.exists(rpath):
            log("[bench] building index for %s (n=%d) ..." % (workload, n))
            host = rib.HostIndex.from_text_auto(text)  # prefix-free parsing above 16 MB: C2 ~1 s, C3 (1 GB) ~6 s, C5 (4 GB) ~24 s
            host.save(rpath + ".tmp")
            os.replace(rpath + ".tmp", rpath)
            del host
    if barrier is not None:
        barrier()
    patt = np.load(ppath)
    host = None
    if need_index:
        rib = rib or ge.load_package()
        host = rib.HostIndex.load(rpath)
    if rank == 0 and host is not None:
        log("[bench] inputs of %s ready in %.1fs: n=%d r=%d n/r=%.1f, %d patterns" % (workload, time.time() - t0, host.n, host.r, host.n / host.r, N))
    return text, patt, N, m, host


def reference_index(workload, host=None):
    """The reference's r_index<> for the workload (oracle/_ref), or the plain-C port when _ref did not travel."""
    ob = ge.load_oracle()
    if ob.have_ref():
        rpath = os.path.join(CACHE, "%s.ref.ri" % base_of(workload))
        if os.path.exists(rpath):
            return ob.RefIndex.load(rpath), "index built by the reference's own constructor (cached .ri, reference serialize/load)"
        if host is None:
            _, _, _, _, host = job_inputs(workload, 1, 0)
        log("[bench] no cached reference index for %s: assembling the reference's structures over this repo's BWT + samples" % workload)
        return ob.RefIndex.from_logical(host.arrays()), "reference structures built by their own constructors over this repo's BWT + samples (ref_from_logical: the suffix sort is bypassed)"
    if not ob.have_port():
        ob.build()
    log("[bench] oracle/_ref missing: CPU baseline falls back to the plain-C port")
    rib = ge.load_package()
    kind, n, p0, p1, tseed = WORKLOADS[workload][:5]
    text = rib.gen_text(kind, n, p0, p1, tseed)
    return ob.PortIndex(text, sa=rib.suffix_array(text)), "plain-C port over a suffix array"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe). NVML is polled from a
    thread every ~2 ms (a timed region of 10 steps lasts ~20 ms: `nvidia-smi -lms` cannot be relied on to land a
    sample inside it, and with 8 GPUs it did not); falls back to an nvidia-smi loop if pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples, self.mask, self.max_mhz = [], 0, None
        self.stop_flag = False
        self.thread = None
        self.p = None
        self.f = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(get_reasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.5)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.samples:
                hi = [x for x in self.samples if x >= 0.5 * max(self.samples)]
                out.update(sm_mhz=statistics.median(hi), sm_max_mhz=self.max_mhz, samples=len(self.samples),
                           reasons=sorted(name for bit, name in self.REASONS.items() if self.mask & bit))
            return out
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # "under load" = samples in the upper half of the observed range
            hi = [x for x in sm if x >= 0.5 * max(sm)]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload` from the committed ncu
    captures (profiles/ncu_traffic.json: {workload: {kernel: bytes}}), or None when no capture of that pair exists."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(kernel)
    except Exception:
        return None


def search_kernel_name():
    """The backward-search kernel the library launches for K = 4 block records (RIG_VARIANT bit 14 / bit 7 select the others)."""
    v = int(os.environ.get("RIG_VARIANT", "0") or 0)
    return "search_kernel" if v & 128 else ("search_lane_kernel" if v & 16384 else "search_pair_kernel")


def d2h_ceiling(world):
    """Aggregate device->host GB/s this box sustains with `world` GPUs copying at once (profiles/r2_d2h_probe.json,
    measured by tools/d2h_probe.py), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_d2h_probe.json")))["d2h_aggregate_GBps_by_ranks"]
        return float(t[str(world)])
    except Exception:
        return None


def regime(index_bytes, seed_bytes=0):
    hot = int(index_bytes) - int(seed_bytes)
    if hot <= L2_BYTES:
        return "A (the tables the dominant kernel reads, %d MB, fit the 126 MB L2: DRAM traffic ~ the output stream)" % (hot >> 20)
    return "B (the tables the dominant kernel reads, %d MB, exceed the 126 MB L2: lookups reach DRAM)" % (hot >> 20)


def ref_locate_sample(ref, patt, N, m, S, threads):
    """Reference CPU path (locate_all per pattern, results dropped as ri-locate does) on the first S patterns."""
    S = max(1, min(N, S))
    _, _, _, occ_total, secs = ref.locate(patt[: S * m], S, m, threads=threads, want=False)
    return occ_total, secs, S


def bounded_sample(N, occ_per_pattern, target_occ):
    """Patterns in a CPU sample of about target_occ occurrences."""
    return int(max(1, min(N, target_occ / max(1.0, occ_per_pattern))))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref) on the box's host threads,
    on the same workload (`config` identical to our arm's); each step a bounded sample of the job."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    ob = ge.load_oracle()
    own_index = ob.have_ref() and os.path.exists(os.path.join(CACHE, "%s.ref.ri" % base_of(args.workload)))
    _, patt, N, m, host = job_inputs(args.workload, world, 0, need_index=not own_index)   # patterns only when the reference has its own index
    ref, how = reference_index(args.workload, host)
    threads = cores if ref.kind == "reference" else 1
    count_mode = args.mode == "count"
    metric, unit = ("count_patterns_per_s", "patterns/s") if count_mode else ("locate_occurrences_per_s", "occ/s")
    if count_mode:
        S = min(N, args.ref_sample or 200_000)
        run = lambda k: (k, ref.count(patt[: k * m], k, m, threads=threads, want=False)[2])  # noqa: E731
    else:
        # size the sample from a probe: about 1.5e8 occurrences per step (~1 s on 16 threads)
        occ_p, _, Sp = ref_locate_sample(ref, patt, N, m, min(N, 2000), threads)
        S = args.ref_sample or bounded_sample(N, occ_p / Sp, 1.5e8)
        run = lambda k: ref_locate_sample(ref, patt, N, m, k, threads)[:2]  # noqa: E731
    for _ in range(args.warmup):
        run(max(1, S // 10))
    work, secs = 0, 0.0
    for _ in range(args.steps):
        w, t = run(S)
        work += w; secs += t
    val = work / secs
    sample = "first %d of the job's %d patterns per step, %s, all %d host threads" % (
        S, N, "count()" if count_mode else "locate_all, results dropped as ri-locate does", threads)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": shared_config(args.workload, world, ref.n, ref.r),
        "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": ref.kind, "sample": sample, "index": how},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
        "repo_libraries_loaded": repo_libraries_loaded()}), flush=True)
    return 0


def repo_libraries_loaded():
    """This repo's native libraries mapped into the process (the reference arm must not need any once its inputs are cached)."""
    try:
        with open("/proc/self/maps") as f:
            return sorted({os.path.basename(l.split()[-1]) for l in f if "librindex_" in l})
    except OSError:
        return None


class Dist:
    """torch.distributed plumbing of one run: NCCL for device tensors, a gloo group for the host-buffer path."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.gloo = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.gloo = dist.new_group(backend="gloo")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.gloo)

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def open_index(rib, host, args, D, kw):
    """The device index of this rank. N > 1 (or --flat): one flatten per box, not one per rank — rank 0 flattens and writes
    the flattened index (rig_index_save_flat) unless the file is already there, every other rank loads that file
    (rig_index_load_flat: a read + an upload; the library checks it against the logical index's digest)."""
    if D.world == 1 and not args.flat:
        return rib.GpuIndex(host, **kw)
    flat = os.path.join(CACHE, "%s.K%d.D%d.S%d.flat" % (base_of(args.workload), kw.get("runs_per_block", 0), kw.get("phi_jump", 0),
                                                         kw.get("seed_jump", 0)))
    gpu = None
    if D.rank == 0:
        gpu = rib.GpuIndex(host, flat=flat, **kw)
        if not gpu.from_flat:
            gpu.save_flat(flat + ".tmp")
            os.replace(flat + ".tmp", flat)
    D.host_barrier()
    if D.rank != 0:
        gpu = rib.GpuIndex(host, flat=flat, **kw)
    return gpu


def run_ours_count(args):
    """--mode count: the ri-count configs (C4). A step = one backward-search pass (rig_count_batch_dev) over the job,
    every rank its contiguous equal-count shard (reads of one length cost the same: no re-balancing); value =
    patterns/s; roofline = the search kernel against HBM (regime B when the flattened index is larger than L2)."""
    import torch
    rib = ge.load_package()
    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    from rindex_b200 import _shard
    text, patt, N, m, host = job_inputs(args.workload, world, rank, barrier=D.host_barrier)
    a, b = _shard.shard_bounds(N, world, rank)
    Ns = b - a
    t0 = time.time()
    gpu = open_index(rib, host, args, D, dict(device=local, runs_per_block=args.runs_per_block, lf_bucket_log2=args.lf_log2,
                                              phi_bucket_log2=args.phi_log2, phi_jump=args.phi_jump or 1, seed_jump=1))  # count only: smallest locate tables
    load_s = time.time() - t0
    info = gpu.info
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    shard = np.ascontiguousarray(patt[a * m: b * m])
    d_patt = torch.from_numpy(shard).to(dev)
    d_lo = torch.empty(max(Ns, 1), dtype=torch.int64, device=dev)
    d_hi = torch.empty(max(Ns, 1), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        gpu.count_dev(d_patt.data_ptr(), Ns, m, d_lo.data_ptr(), d_hi.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    D.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    D.barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(); step(); ev[k][1].record()
        torch.cuda.synchronize()
        t = gpu.timing()
        kern_ms.append(t["search_ms"])
    D.barrier()
    clocks = sampler.stop()
    total_ms = sum(x.elapsed_time(y) for x, y in ev)
    lf_steps = t["lf_steps"]
    # end to end: host buffers (pinned) through rig_count_batch
    h_patt = torch.from_numpy(shard).pin_memory()
    h_lo = torch.empty(max(Ns, 1), dtype=torch.int64).pin_memory()
    h_hi = torch.empty(max(Ns, 1), dtype=torch.int64).pin_memory()
    for _ in range(2):
        gpu.count_raw(h_patt.data_ptr(), Ns, m, h_lo.data_ptr(), h_hi.data_ptr())
    D.barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t1 = time.perf_counter()
    for k in range(e2e_steps):
        gpu.count_raw(h_patt.data_ptr(), Ns, m, h_lo.data_ptr(), h_hi.data_ptr())
    e2e_t = time.perf_counter() - t1
    D.barrier()
    nocc = (h_hi[:Ns] - h_lo[:Ns] + 1).clamp(min=0)
    occ_t = int(nocc[(h_hi[:Ns] >= h_lo[:Ns])].sum())
    total_ms_g, e2e_ms_g = D.reduce([total_ms, e2e_t * 1e3 / e2e_steps * args.steps], "MAX")
    N_g, lf_g, occ_g = D.reduce([float(Ns), float(lf_steps), float(occ_t)], "SUM")
    if rank == 0:
        peak, peak_src = hbm_peak()
        k_ms = statistics.mean(kern_ms)
        # per rank query this layout touches one directory sector (32 B) and one block record (32/64/96 B); 2 queries per LF step
        q_bytes = 32 + int(info.lf_record_bytes)
        alg_bytes = int(lf_steps) * 2 * q_bytes + Ns * (m + 16)
        skern = search_kernel_name()
        ell = max(1, int(np.ceil(np.log2(max(2, info.sigma)))))
        line = {
            "metric": "count_patterns_per_s", "value": N_g * args.steps / (total_ms_g * 1e-3), "unit": "patterns/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_g / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": shared_config(args.workload, world, info.n, info.r),
            "detail": {"sigma": int(info.sigma), "patterns_rank0": Ns, "lf_steps_rank0": int(lf_steps), "total_occurrences": int(occ_g),
                       "index_device_bytes": int(info.device_bytes), "index_load_s": round(load_s, 3),
                       "index_from_flat_file": bool(getattr(gpu, "from_flat", False)),
                       "runs_per_block": int(info.runs_per_block), "parallelism": "contiguous equal-count shards x%d, index replicated" % world},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
            "e2e": {"value": N_g * args.steps / (e2e_ms_g * 1e-3), "unit": "patterns/s", "h2d_bytes_per_step": int(N * m),
                    "d2h_bytes_per_step": int(16 * N), "api": "rig_count_batch (host buffers, pinned)", "ms_per_step": e2e_ms_g / args.steps},
            "gpu_launches": args.steps * world,
            "lf_steps_per_s": lf_g * args.steps / (total_ms_g * 1e-3),
            "roofline": {"bound": "hbm", "kernel": skern, "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(args.workload, skern),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_ms,
                         "algorithmic_bytes": "per LF step 2 rank queries x (32 B directory sector + %d B block record) + m + 16 B per pattern; "
                                              "every touched sector counted as a DRAM fetch: an upper bound when the index fits L2 or reads revisit "
                                              "the same blocks (then `traffic`, where a capture exists, is the DRAM figure)" % int(info.lf_record_bytes),
                         "regime": regime(info.device_bytes),
                         "survey_touched": {"bytes_per_lf_step": 2 * B_RANK(ell),
                                            "achieved": int(lf_steps) * 2 * B_RANK(ell) / (k_ms * 1e-3) / 1e9}},
        }
        if world == 1 and not args.no_cpu:
            ref, how = reference_index(args.workload, host)
            cores = os.cpu_count() or 1
            threads = cores if ref.kind == "reference" else 1
            S = min(N, args.cpu_sample or 200_000)
            ref.count(patt[: max(1, S // 10) * m], max(1, S // 10), m, threads=threads, want=False)
            best = min(ref.count(patt[: S * m], S, m, threads=threads, want=False)[2] for _ in range(3))
            line["cpu_baseline"] = {"value": S / best, "unit": "patterns/s", "cores": threads, "kind": ref.kind, "index": how,
                                    "sample": "%d of %d patterns, reference count() loop, best of 3: %.2fs" % (S, N, best)}
        print(json.dumps(line), flush=True)
    D.close()
    return 0


PLAN_REPLICATE_MAX_LF = 8_000_000   # LF steps (patterns x length) up to which every rank counts the whole batch itself


class LocateJob:
    """One rank's part of a locate job: device-resident step and host-buffer (e2e) step, with the count -> all-gather
    -> re-cut by occurrence mass -> locate sequence when there is more than one rank (`solo` = the whole job alone).
    Equal-count shards are [r * per, (r + 1) * per) with per = ceil(N / world), so the all-gathered count buffer IS
    the batch's count array (plus padding at its very end)."""

    def __init__(self, D, gpu, patt, N, m, stream, solo=False, plan="auto"):
        torch = D.torch
        self.D, self.gpu, self.N, self.m, self.stream = D, gpu, N, m, stream
        self.world, self.rank = (1, 0) if solo else (D.world, D.rank)
        # Planning: "gather" = every rank counts its equal-count shard, the ranks all-gather the counts, every rank then
        # locates (search + expansion) its re-cut shard; "replicate" = every rank SEARCHES the whole batch itself
        # (rig_plan_batch_dev: ranges, toeholds, output offsets, and the cut points by binary search in the offsets) and
        # then only EXPANDS its shard (rig_expand_shard_dev) — no collective at all, and cheaper whenever a search pass
        # over the batch costs less than a collective's latency (C5: 1.5 M LF steps = 0.06 ms against ~0.3 ms of count
        # + all-gather + cuts at 8 ranks). auto: replicate up to 8 M LF steps per batch.
        self.plan = plan if plan != "auto" else ("replicate" if N * m <= PLAN_REPLICATE_MAX_LF else "gather")
        from rindex_b200 import _shard
        self.shard = _shard
        self.patt = patt
        self.d_patt = torch.from_numpy(patt).to(D.dev)          # the whole batch on every GPU (N x m bytes)
        self.per = (N + self.world - 1) // self.world
        self.a, self.b = min(N, self.rank * self.per), min(N, (self.rank + 1) * self.per)
        self.d_lo = torch.empty(N + 1, dtype=torch.int64, device=D.dev)
        self.d_hi = torch.empty(N + 1, dtype=torch.int64, device=D.dev)
        self.d_off = torch.empty(N + 2, dtype=torch.int64, device=D.dev)
        self.d_cnt = torch.zeros(self.per, dtype=torch.int64, device=D.dev)
        self.d_all = torch.zeros(self.world * self.per, dtype=torch.int64, device=D.dev)
        self.d_occ = None
        self.cuts = None
        self.plan_launches = 0

    # -- re-balancing: equal-count count phase, all-gather of the counts, cuts of equal occurrence mass --
    def plan_dev(self):
        """count phase on the device (rig_count_batch_dev + rig_counts_dev), NCCL all-gather of the counts, cut points
        by the library's device kernel (rig_balanced_cuts_dev: one launch, world + 1 words back to the host, where the
        shard bounds are launch parameters). Returns this rank's [c0, c1)."""
        if self.world == 1:
            return 0, self.N
        if self.plan == "replicate":   # search + offsets of the whole batch here, cuts from the offsets (rig_plan_batch_dev)
            self.cuts, _ = self.gpu.plan_dev(self.d_patt.data_ptr(), self.N, self.m, self.d_lo.data_ptr(), self.d_hi.data_ptr(),
                                             self.d_off.data_ptr(), self.world, 64, self.stream)
            self.plan_launches = 3   # prep_kernel, search_pair_kernel, cuts_from_offsets_kernel
            return self.cuts[self.rank], self.cuts[self.rank + 1]
        else:
            n = self.b - self.a
            self.gpu.count_dev(self.d_patt.data_ptr() + self.a * self.m, n, self.m, self.d_lo.data_ptr(), self.d_hi.data_ptr(), self.stream)
            self.gpu.counts_dev(self.d_lo.data_ptr(), self.d_hi.data_ptr(), n, self.d_cnt.data_ptr(), self.stream)
            self.D.dist.all_gather_into_tensor(self.d_all, self.d_cnt)
        self.cuts = self.gpu.balanced_cuts_dev(self.d_all.data_ptr(), self.N, self.world, 64, self.stream)
        self.plan_launches = 4   # prep, search (count), counts, cuts — the all-gather is NCCL's kernel, not counted
        return self.cuts[self.rank], self.cuts[self.rank + 1]

    def counts_all(self):
        if self.world > 1 and self.plan == "replicate":   # after plan_dev: d_off is the exclusive prefix of the counts
            off = self.d_off[: self.N + 1].cpu().numpy()
            return off[1:] - off[:-1]
        return self.d_all[: self.N].cpu().numpy()

    def step_dev(self, mark=None):
        c0, c1 = self.plan_dev()
        if mark is not None:   # end of the planning phase (count, all-gather, cuts: plan_dev ends on a host sync)
            mark.record()
        if self.world > 1 and self.plan == "replicate":   # the batch is searched already: expand this rank's shard
            return self.gpu.expand_shard_dev(self.N, c0, c1, self.d_lo.data_ptr(), self.d_hi.data_ptr(), self.d_off.data_ptr(),
                                             self.d_occ.data_ptr(), self.d_occ.numel(), self.stream)
        return self.gpu.locate_dev(self.d_patt.data_ptr() + c0 * self.m, c1 - c0, self.m, self.d_lo.data_ptr(), self.d_hi.data_ptr(),
                                   self.d_off.data_ptr(), self.d_occ.data_ptr(), self.d_occ.numel(), self.stream)

    def size_output(self):
        """Two-call protocol on this rank's shard; allocates the device occurrence buffer. Returns its occurrences."""
        torch = self.D.torch
        c0, c1 = self.plan_dev()
        try:
            if self.world > 1 and self.plan == "replicate":
                need = self.gpu.expand_shard_dev(self.N, c0, c1, self.d_lo.data_ptr(), self.d_hi.data_ptr(), self.d_off.data_ptr(), None, 0, self.stream)
            else:
                need = self.gpu.locate_dev(self.d_patt.data_ptr() + c0 * self.m, c1 - c0, self.m, self.d_lo.data_ptr(), self.d_hi.data_ptr(),
                                           self.d_off.data_ptr(), None, 0, self.stream)
        except Exception as e:  # noqa: BLE001
            if getattr(e, "code", 0) != -4:
                raise
            need = e.needed
        self.d_occ = torch.empty(max(need, 1), dtype=torch.int64, device=self.D.dev)
        self.c0, self.c1 = c0, c1
        return need

    # -- host-buffer path (the reference-facing calls): pinned buffers, gloo all-gather of the counts --
    def setup_host(self, occ_cap):
        torch = self.D.torch
        self.h_patt = torch.from_numpy(self.patt).pin_memory()
        self.h_lo = torch.empty(self.N + 1, dtype=torch.int64).pin_memory()
        self.h_hi = torch.empty(self.N + 1, dtype=torch.int64).pin_memory()
        self.h_off = torch.empty(self.N + 2, dtype=torch.int64).pin_memory()
        self.h_occ = torch.empty(max(occ_cap, 1), dtype=torch.int64).pin_memory()
        self.h_cnt = torch.zeros(self.per, dtype=torch.int64)
        self.h_all = torch.zeros(self.world * self.per, dtype=torch.int64)

    def step_host(self):
        torch, dist = self.D.torch, self.D.dist
        c0, c1 = 0, self.N
        if self.world > 1:
            if self.plan == "replicate":
                n = self.N
                self.gpu.count_raw(self.h_patt.data_ptr(), n, self.m, self.h_lo.data_ptr(), self.h_hi.data_ptr())
                torch.sub(self.h_hi[:n], self.h_lo[:n], out=self.h_all[:n])
                self.h_all[:n].add_(1).clamp_(min=0)
            else:
                n = self.b - self.a
                self.gpu.count_raw(self.h_patt.data_ptr() + self.a * self.m, n, self.m, self.h_lo.data_ptr(), self.h_hi.data_ptr())
                torch.sub(self.h_hi[:n], self.h_lo[:n], out=self.h_cnt[:n])
                self.h_cnt[:n].add_(1).clamp_(min=0)
                dist.all_gather_into_tensor(self.h_all, self.h_cnt, group=self.D.gloo)
            cuts = self.shard.balanced_cuts(self.h_all[: self.N].numpy(), self.world)
            c0, c1 = cuts[self.rank], cuts[self.rank + 1]
        tot = self.gpu.locate_raw(self.h_patt.data_ptr() + c0 * self.m, c1 - c0, self.m, self.h_lo.data_ptr(), self.h_hi.data_ptr(),
                                  self.h_off.data_ptr(), self.h_occ.data_ptr(), self.h_occ.numel())
        return tot, c1 - c0


def measure_locate(D, job, steps, warmup, e2e_steps, flush, solo=False):
    """Timed region of one job: `steps` device-resident steps (CUDA events per step, L2 flushed in between) and
    `e2e_steps` host-buffer steps (wall clock around the whole loop, barriers on both sides). Returns a dict of this
    rank's figures; the caller reduces over ranks."""
    torch = D.torch
    gpu = job.gpu
    barrier = (lambda: torch.cuda.synchronize()) if solo else D.barrier
    occ_rank = job.size_output()
    for _ in range(warmup):
        job.step_dev()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ph = {k: [] for k in ("search_ms", "scan_ms", "seed_ms", "window_ms", "expand_ms")}
    plan = []
    launches = 0
    barrier()
    for k in range(steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
        mark = torch.cuda.Event(enable_timing=True) if job.world > 1 else None
        ev[k][0].record()
        tot = job.step_dev(mark)
        ev[k][1].record()
        torch.cuda.synchronize()
        t = gpu.timing()
        for key in ph:
            ph[key].append(t[key])
        if mark is not None:
            plan.append(ev[k][0].elapsed_time(mark))
        launches += t["launches"] + (job.plan_launches if job.world > 1 else 0)
    barrier()
    assert tot == occ_rank
    total_ms = sum(x.elapsed_time(y) for x, y in ev)
    out = {"occ_rank": occ_rank, "patterns_rank": job.c1 - job.c0, "total_ms": total_ms, "launches": launches, "expansion_kernels": t["slices"],
           "lf_steps": t["lf_steps"], "chains": t["chains"], "phases": {k: statistics.mean(v) for k, v in ph.items()}}
    if plan:   # N > 1: count on the equal-count shard + all-gather of the counts + cut points, before the locate of the shard
        out["phases"]["plan_ms"] = statistics.mean(plan)
    # end to end through the host-buffer C-ABI calls, pinned host memory. A shard whose occurrences exceed 16 GB (C3 at
    # full size) is not measured end to end: pinning that much host memory per rank is not what this number is about.
    if e2e_steps and occ_rank * 8 <= (16 << 30):
        job.setup_host(occ_rank)
        for _ in range(2):
            job.step_host()
        barrier()
        if not solo:
            D.host_barrier()
        t1 = time.perf_counter()
        for k in range(e2e_steps):
            tot_h, pat_h = job.step_host()
        e2e_t = time.perf_counter() - t1
        barrier()
        assert tot_h == occ_rank and int(job.h_off[pat_h]) == occ_rank   # cheap self-check (parity tests live in tests/)
        n_cnt = 0 if job.world == 1 else (job.N if job.plan == "replicate" else job.b - job.a)   # patterns of the count call
        out.update(e2e_ms=e2e_t * 1e3 / e2e_steps, e2e_h2d=int(n_cnt * job.m + pat_h * job.m),
                   e2e_d2h=int(8 * (2 * n_cnt + 3 * pat_h + 1 + occ_rank)))
    return out


def run_ours(args):
    import torch
    rib = ge.load_package()
    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    need_text = world == 1 and not args.no_post
    text, patt, N, m, host = job_inputs(args.workload, world, rank, need_text=need_text, barrier=D.host_barrier)
    t0 = time.time()
    kw = dict(device=local, runs_per_block=args.runs_per_block, lf_bucket_log2=args.lf_log2, phi_bucket_log2=args.phi_log2,
              expand_threads=args.expand_threads, phi_jump=args.phi_jump, seed_jump=args.seed_jump)
    gpu = open_index(rib, host, args, D, kw)
    load_s = time.time() - t0
    info = gpu.info
    # the library launches on the stream it is given (NULL would mean its own stream): use a real
    # torch stream so that torch.cuda.Event sees the same stream the kernels run on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    e2e_steps = max(1, min(args.steps, 5))

    # ---- the same job on ONE GPU, measured in this run by rank 0 while the others wait (N > 1 only) ----
    solo = None
    solo_digest = None
    if world > 1 and not args.no_solo:
        D.host_barrier()   # every rank has loaded its index: nothing else runs on the host during the one-GPU run
        if rank == 0:
            sj = LocateJob(D, gpu, patt, N, m, stream, solo=True)
            sm = measure_locate(D, sj, max(3, min(args.steps, 5)), 3, 2, flush, solo=True)
            solo_digest = gpu.digest_dev(sj.d_occ.data_ptr(), sm["occ_rank"], stream)   # of the whole job's output, in order
            solo = {"value": sm["occ_rank"] * max(3, min(args.steps, 5)) / (sm["total_ms"] * 1e-3), "ms_per_step": sm["total_ms"] / max(3, min(args.steps, 5)),
                    "e2e": (sm["occ_rank"] / (sm["e2e_ms"] * 1e-3)) if "e2e_ms" in sm else None,
                    "e2e_ms_per_step": sm.get("e2e_ms"), "occurrences": sm["occ_rank"], "phases_ms": sm["phases"],
                    "expansion_kernels": sm["expansion_kernels"]}
            del sj
            torch.cuda.empty_cache()
        # the other ranks wait on the HOST (gloo): a device-side barrier would leave seven NCCL kernels polling over
        # NVLink for the whole one-GPU run (measured: it slows rank 0's kernels by 25-75%)
        D.host_barrier()
        D.barrier()

    job = LocateJob(D, gpu, patt, N, m, stream, plan=args.plan)
    sampler = ClockSampler(local)
    sampler.start()
    M = measure_locate(D, job, args.steps, args.warmup, e2e_steps, flush)
    clocks = sampler.stop()
    occ_rank, lf_steps, chains = M["occ_rank"], M["lf_steps"], M["chains"]

    # count-only pass (patterns/s) on the equal-count shard, device resident
    a, b = job.a, job.b
    cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def count_dev():
        gpu.count_dev(job.d_patt.data_ptr() + a * m, b - a, m, job.d_lo.data_ptr(), job.d_hi.data_ptr(), stream)

    count_dev(); D.barrier()
    for k in range(args.steps):
        flush.zero_()
        cev[k][0].record(); count_dev(); cev[k][1].record()
    D.barrier()
    count_ms = sum(x.elapsed_time(y) for x, y in cev)

    # the same host-buffer call with 32-bit positions (rig_locate_batch32; texts below 4 GiB): half the D2H bytes
    e2e32_ms = None
    e2e32_dev = None
    if world == 1 and int(info.n) <= 0xFFFFFFFF and "e2e_ms" in M:
        h_occ32 = torch.empty(max(occ_rank, 1), dtype=torch.int32).pin_memory()
        call32 = lambda: gpu.locate32_raw(job.h_patt.data_ptr(), N, m, job.h_lo.data_ptr(), job.h_hi.data_ptr(), job.h_off.data_ptr(),  # noqa: E731
                                          h_occ32.data_ptr(), h_occ32.numel())
        for _ in range(2):
            call32()
        t1 = time.perf_counter()
        for k in range(e2e_steps):
            call32()
        e2e32_ms = (time.perf_counter() - t1) * 1e3 / e2e_steps
        t32 = gpu.timing()   # device phases of the last 32-bit call (warm L2: the e2e loop does not flush)
        e2e32_dev = {k: float(t32[k]) for k in ("search_ms", "expand_ms", "seed_ms", "window_ms", "d2h_ms")}
        assert torch.equal(h_occ32[:4096].to(torch.int64) & 0xFFFFFFFF, job.h_occ[:4096])
        del h_occ32

    # ri-locate -o / -c post-processing on the device, timed once on the resident output of the last
    # step: segmented sort of every pattern's occurrences, then the self-check (hash-join brute-force counts over
    # the text + byte comparison of every occurrence). Not part of `value`. One GPU only.
    post = None
    if world == 1 and not args.no_post:
        job.step_dev(); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); gpu.sort_dev(N, job.d_off.data_ptr(), job.d_occ.data_ptr(), occ_rank, stream); s1.record()
        torch.cuda.synchronize()
        gpu.text_attach(text)
        t1 = time.perf_counter()
        rep = gpu.check_dev(job.d_patt.data_ptr(), N, m, job.d_lo.data_ptr(), job.d_hi.data_ptr(), job.d_off.data_ptr(), job.d_occ.data_ptr(),
                            occ_rank, True, stream)
        check_ms = (time.perf_counter() - t1) * 1e3
        assert rep.clean, rep.as_dict()
        post = {"sort_ms": s0.elapsed_time(s1), "sort_keys_per_s": occ_rank / (s0.elapsed_time(s1) * 1e-3),
                "check_ms": check_ms, "check": rep.as_dict(),
                "note": "rig_sort_occurrences_dev + rig_check_dev on the device-resident output (ri-locate -o / -c)"}

    # the sharded job's output IS the one-GPU run's output: order-sensitive digest of the concatenated shards, combined
    # from the per-rank digests (sum and index-weighted sum are additive with the shard's base index), against rank 0's
    # digest of its one-GPU run of the same job
    same_as_n1 = None
    if world > 1:
        job.step_dev(); torch.cuda.synchronize()
        mine = gpu.digest_dev(job.d_occ.data_ptr(), occ_rank, stream)
        piece = torch.tensor([mine[0] & 0xFFFFFFFF, mine[0] >> 32, mine[1] & 0xFFFFFFFF, mine[1] >> 32, occ_rank], dtype=torch.int64, device=dev)
        pieces = torch.zeros(5 * world, dtype=torch.int64, device=dev)
        D.dist.all_gather_into_tensor(pieces, piece)
        pieces = pieces.cpu().tolist()
        if rank == 0 and solo_digest is not None:
            M64 = (1 << 64) - 1
            S = W = base = 0
            for r in range(world):
                s_r = pieces[5 * r] | (pieces[5 * r + 1] << 32)
                w_r = pieces[5 * r + 2] | (pieces[5 * r + 3] << 32)
                S = (S + s_r) & M64
                W = (W + w_r + base * s_r) & M64
                base += pieces[5 * r + 4]
            same_as_n1 = bool(S == solo_digest[0] and W == solo_digest[1])

    # optional collation over NVLink (north_star: "an optional NCCL all-gather only to collate occurrence buffers"):
    # every rank ends up with the whole job's occurrences in pattern order; exercised and timed once, not part of `value`
    collate = None
    if world > 1 and not args.no_collate:
        job.step_dev(); torch.cuda.synchronize()
        sizes = torch.tensor([occ_rank], dtype=torch.int64, device=dev)
        all_sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        D.dist.all_gather_into_tensor(all_sizes, sizes)
        all_sizes = [int(x) for x in all_sizes.cpu().tolist()]
        mx = max(all_sizes)
        if mx * 8 * (world + 1) < (120 << 30):
            send = torch.zeros(mx, dtype=torch.int64, device=dev)
            send[:occ_rank] = job.d_occ[:occ_rank]
            recv = torch.empty(world * mx, dtype=torch.int64, device=dev)
            mine = gpu.digest_dev(job.d_occ.data_ptr(), occ_rank, stream)
            D.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); D.dist.all_gather_into_tensor(recv, send); c1.record()
            torch.cuda.synchronize()
            ok = True
            dg = torch.tensor([mine[0] & 0x7FFFFFFFFFFFFFFF, mine[1] & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
            all_dg = torch.zeros(2 * world, dtype=torch.int64, device=dev)
            D.dist.all_gather_into_tensor(all_dg, dg)
            all_dg = all_dg.cpu().tolist()
            for r in range(world):   # every rank checks every shard it received against the sender's own digest
                got = gpu.digest_dev(recv.data_ptr() + r * mx * 8, all_sizes[r], stream)
                ok = ok and (got[0] & 0x7FFFFFFFFFFFFFFF) == all_dg[2 * r] and (got[1] & 0x7FFFFFFFFFFFFFFF) == all_dg[2 * r + 1]
            okg = D.reduce([1.0 if ok else 0.0], "MIN")[0]
            ms = D.reduce([c0.elapsed_time(c1)], "MAX")[0]
            collate = {"ms": ms, "bytes_received_per_rank": int(8 * sum(all_sizes)), "padded_bytes_per_rank": int(8 * mx * world),
                       "algbw_GBps": 8 * sum(all_sizes) / (ms * 1e-3) / 1e9, "digests_match_on_all_ranks": bool(okg == 1.0),
                       "api": "torch.distributed all_gather_into_tensor (NCCL over NVLink), shards padded to the largest"}
            del send, recv

    # reduce over ranks: max time, sum work
    total_ms_g, count_ms_g, e2e_ms_g = D.reduce([M["total_ms"], count_ms, M.get("e2e_ms", 0.0)], "MAX")
    occ_g, launches_g, h2d_g, d2h_g, has_e2e = D.reduce([float(occ_rank), float(M["launches"]), float(M.get("e2e_h2d", 0)),
                                                           float(M.get("e2e_d2h", 0)), 1.0 if "e2e_ms" in M else 0.0], "SUM")
    rank_ms = D.reduce([M["total_ms"] / args.steps], "SUM")[0] / world
    occ_max = D.reduce([float(occ_rank)], "MAX")[0]
    e2e_sum_ms = D.reduce([M.get("e2e_ms", 0.0)], "SUM")[0]
    eq_occ = None
    if world > 1:   # what equal-count shards would have given each rank (the imbalance the re-cut removes)
        job.plan_dev()
        nocc_all = job.counts_all()
        eq = [float(nocc_all[min(N, r * job.per): min(N, (r + 1) * job.per)].sum()) for r in range(world)]
        eq_occ = max(eq) / (sum(eq) / world) if sum(eq) else 1.0

    if rank == 0:
        peak, peak_src = hbm_peak()
        phs = M["phases"]
        exp_ms = phs["expand_ms"]
        # Roofline of the dominant kernel (the fused expansion kernel; with RIG_VARIANT bit 13 the window pass of the
        # two-kernel form), HBM-bound by the contract's definition. ALGORITHMIC bytes = what must cross HBM for this
        # launch: 8 B per occurrence written + one pass over the tables it reads + the item list (+ one seed-table
        # record per item when the kernel also produces the items).
        two_pass = int(info.seed_jump) > 1 and phs["window_ms"] > 0
        items = occ_rank // max(1, int(info.seed_jump)) + int(chains)  # upper bound: (L-1)/SEG + 1 items per chain of L
        if two_pass and M["expansion_kernels"] == 1:
            # fused producer/consumer kernel: it also writes the items and hops through the seed table (one 64-byte record per item)
            alg_bytes = occ_rank * 8 + int(info.device_bytes) - int(info.seed_bytes) + items * (16 + 16 + 64)
            dom_kernel, dom_ms = "phi_fused_kernel", phs["window_ms"]
            alg_note = ("8 B/occurrence output + one pass over the flattened index without the seed table + per item: 16 B written, "
                        "16 B read, one 64 B seed-table record")
        elif two_pass:
            alg_bytes = occ_rank * 8 + int(info.device_bytes) - int(info.seed_bytes) + items * 16
            dom_kernel, dom_ms = "phi_window_batch_kernel", phs["window_ms"]
            alg_note = "8 B/occurrence output + one pass over the flattened index without the seed table + 16 B per item (slot, count, seed)"
        else:
            alg_bytes = occ_rank * 8 + int(info.device_bytes)
            dom_kernel, dom_ms = "phi_expand_kernel", exp_ms
            alg_note = "8 B/occurrence output + one pass over the flattened index"
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        # the same launch as 32-byte sector requests through the SM->crossbar ports (what ncu shows it is bound by,
        # profiles/r2_window_bound.txt): one store sector and one table-lookup sector per D occurrences, two sectors
        # per item; a copy at the HBM peak issues peak/32 sectors per second through the same ports
        Dj = max(1, int(info.phi_jump))
        sectors = 2 * (occ_rank // Dj) + 2 * items
        request_rate = {"sectors_per_launch": sectors, "achieved_Gsectors_per_s": sectors / (dom_ms * 1e-3) / 1e9,
                        "copy_at_hbm_peak_Gsectors_per_s": peak / 32.0, "frac": sectors / (dom_ms * 1e-3) / 1e9 / (peak / 32.0),
                        "note": "half of these sectors are L2-served table lookups: the HBM fraction above is structurally <= ~0.5 at this request rate"}
        step_ms = M["total_ms"] / args.steps
        # the whole step (search + expansion): output, one pass over the tables, the item list written and read with one seed
        # record per item, the patterns and the per-pattern search results (lo, hi, toehold, run, two offsets)
        step_bytes = (occ_rank * 8 + int(info.device_bytes) - int(info.seed_bytes) + (items * (16 + 16 + 64) if two_pass else 0)
                      + M["patterns_rank"] * (m + 48))
        ell = max(1, int(np.ceil(np.log2(max(2, info.sigma)))))
        e2e_val = occ_g / (e2e_ms_g * 1e-3) if has_e2e == world and e2e_ms_g > 0 else None
        line = {
            "metric": "locate_occurrences_per_s", "value": occ_g * args.steps / (total_ms_g * 1e-3), "unit": "occ/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_g / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": shared_config(args.workload, world, info.n, info.r),
            "detail": {"sigma": int(info.sigma), "occurrences_per_step": int(occ_g), "occurrences_rank0": occ_rank,
                       "patterns_rank0": M["patterns_rank"], "phi_chains_rank0": int(chains), "lf_steps_rank0": int(lf_steps),
                       "index_device_bytes": int(info.device_bytes), "index_load_s": round(load_s, 3),
                       "index_from_flat_file": bool(getattr(gpu, "from_flat", False)),
                       "runs_per_block": int(info.runs_per_block), "phi_jump": int(info.phi_jump),
                       "seed_jump": int(info.seed_jump), "seed_table_bytes": int(info.seed_bytes),
                       "parallelism": ("index replicated x%d; per step: %s, contiguous shards cut at equal occurrence mass, locate"
                                       % (world, "every rank searches the whole batch itself (rig_plan_batch_dev: no collective) and expands its shard (rig_expand_shard_dev)" if job.plan == "replicate" else
                                          "count on equal-count shards, NCCL all-gather of the counts (8 B/pattern)")) if world > 1 else "one GPU",
                       "plan": job.plan if world > 1 else None,
                       "phases_ms_rank0": phs},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
            "e2e": {"value": e2e_val, "unit": "occ/s", "h2d_bytes_per_step": int(h2d_g), "d2h_bytes_per_step": int(d2h_g),
                    "api": "rig_locate_batch (host buffers, pinned)" + ("" if world == 1 else ("; rig_count_batch of the whole batch before it" if job.plan == "replicate"
                                                                     else "; rig_count_batch + gloo all-gather of the counts before it")),
                    "ms_per_step": e2e_ms_g, "steps": e2e_steps},
            "gpu_launches": int(launches_g) + args.steps * world,
            "count": {"metric": "count_patterns_per_s", "value": N * args.steps / (count_ms_g * 1e-3), "unit": "patterns/s",
                      "ms_per_step": count_ms_g / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload, dom_kernel), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms, "algorithmic_bytes": alg_note,
                         "rank": 0, "regime": regime(info.device_bytes, info.seed_bytes),
                         "whole_step": {"ms": step_ms, "algorithmic_bytes": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                                        "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak},
                         "request_rate": request_rate,
                         "survey_touched": {"bytes_per_occurrence": B_PHI, "achieved": occ_rank * B_PHI / (exp_ms * 1e-3) / 1e9,
                                            "note": "SURVEY 8d's touched-bytes figure for the reference's structure / expansion time; L2-served, not an HBM fraction"},
                         "search_kernel": {"name": search_kernel_name(), "launch_ms": phs["search_ms"], "lf_steps": int(lf_steps),
                                           "touched_bytes": int(lf_steps) * 2 * (32 + int(info.lf_record_bytes)),
                                           "note": "per LF step 2 rank queries x (directory sector + block record); L2-served in regime A"}},
        }
        if world > 1:
            ceil = d2h_ceiling(world)
            line["balance"] = {"occurrences_max_over_mean": occ_max / (occ_g / world), "equal_count_shards_would_give": eq_occ,
                               "rank_step_ms_max_over_mean": (total_ms_g / args.steps) / rank_ms if rank_ms else None,
                               "rank_e2e_ms_max_over_mean": (e2e_ms_g / (e2e_sum_ms / world)) if e2e_sum_ms else None}
            line["strong_scaling"] = {"job": WORKLOADS[args.workload][9], "n1": solo, "output_identical_to_n1": same_as_n1,
                                      "speedup_value": (line["value"] / solo["value"]) if solo else None,
                                      "speedup_e2e": (e2e_val / solo["e2e"]) if solo and solo.get("e2e") and e2e_val else None}
            if ceil and e2e_val:
                line["e2e"]["box_d2h_ceiling"] = {"aggregate_GBps": ceil, "occ_per_s_at_ceiling": ceil * 1e9 / 8,
                                                  "fraction_of_ceiling": e2e_val / (ceil * 1e9 / 8),
                                                  "source": "profiles/r2_d2h_probe.json (tools/d2h_probe.py: %d GPUs copying to pinned host memory at once)" % world}
        if collate is not None:
            line["collate"] = collate
        if e2e32_ms is not None:
            line["e2e_u32"] = {"value": occ_rank / (e2e32_ms * 1e-3), "unit": "occ/s", "ms_per_step": e2e32_ms,
                               "d2h_bytes_per_step": int(8 * (3 * N + 1) + 4 * occ_rank), "device_phases_ms": e2e32_dev,
                               "api": "rig_locate_batch32 (32-bit positions, n < 2^32; an addition to the reference's 64-bit surface)"}
        if post is not None:
            line["post"] = post
        if world == 1 and not args.no_cpu:
            ref, how = reference_index(args.workload, host)
            cores = os.cpu_count() or 1
            threads = cores if ref.kind == "reference" else 1
            S = args.cpu_sample or bounded_sample(N, occ_rank / max(1, N), 2.0e8)
            ref_locate_sample(ref, patt, N, m, max(1, S // 10), threads)  # warm-up
            best = None
            for _ in range(3):  # best of 3, all host threads
                occ_c, secs_c, S = ref_locate_sample(ref, patt, N, m, S, threads)
                best = secs_c if best is None else min(best, secs_c)
            line["cpu_baseline"] = {"value": occ_c / best, "unit": "occ/s", "cores": threads, "kind": ref.kind, "index": how,
                                    "sample": "%d of %d patterns (%d occurrences), reference locate_all loop, results dropped as ri-locate does, best of 3: %.2fs" % (S, N, occ_c, best)}
            if ref.kind == "reference":  # the reference as shipped is single-threaded: 1-core figure on 1/8 of the sample
                occ_1, secs_1, S1 = ref_locate_sample(ref, patt, N, m, max(1, S // 8), 1)
                line["cpu_baseline"]["single_core_value"] = occ_1 / secs_1
        print(json.dumps(line), flush=True)
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="default: c2 on one GPU, c5 (strong scaling) on several")
    ap.add_argument("--cpu-sample", type=int, default=0, help="patterns in the cpu_baseline sample (default: ~2e8 occurrences)")
    ap.add_argument("--ref-sample", type=int, default=0, help="patterns per step of --impl reference (default: ~1.5e8 occurrences)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-post", action="store_true", help="skip the -o / -c post-processing timing")
    ap.add_argument("--flat", action="store_true", help="N = 1: keep / reuse the flattened index file in .cache (N > 1 always does)")
    ap.add_argument("--plan", default="auto", choices=["auto", "gather", "replicate"],
                    help="N > 1: how the occurrence counts reach every rank before the shards are cut: all-gather of per-shard "
                         "counts, or every rank counting the whole batch (auto: replicate up to 8 M LF steps per batch)")
    ap.add_argument("--no-solo", action="store_true", help="N > 1: skip rank 0's one-GPU run of the same job")
    ap.add_argument("--no-collate", action="store_true", help="N > 1: skip the NCCL collation of the occurrence buffers")
    ap.add_argument("--runs-per-block", type=int, default=0)
    ap.add_argument("--lf-log2", type=int, default=0)
    ap.add_argument("--phi-log2", type=int, default=0)
    ap.add_argument("--expand-threads", type=int, default=0)
    ap.add_argument("--phi-jump", type=int, default=0)
    ap.add_argument("--seed-jump", type=int, default=0, help="two-pass expansion window SEG: 0 auto, 1 off (single pass), 16..256")
    ap.add_argument("--mode", default="locate", choices=["locate", "count"], help="locate (default) or count (ri-count configs: c4, c4s)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "c2" if max(world, args.gpus) == 1 else "c5"
    if args.workload in ("c4", "c4s"):
        args.mode = "count"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: anything a library prints to fd 1 meanwhile (NCCL's version banner, ...)
    # is sent to stderr, and the real stdout is restored for the final print() calls through sys.stdout.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours_count(args) if args.mode == "count" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
