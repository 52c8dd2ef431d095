#!/usr/bin/env python
"""bench.py -- RLZ-parsed Gbp/s on B200 (BASELINE.json metric) next to the CPU restatement of the reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path (gap strip + 2-bit pack + exact greedy RLZ parse + phrase records) over the
rank's batch of haplotypes.  Workload at every N: BASELINE.json configs[1] per GPU (chr21-sized 48 Mbp synthetic
reference, 50 haplotypes) -- haplotypes shard across ranks, no data-path collective, weak scaling; for N > 1 the
index is built on rank 0 and broadcast with NCCL (reported separately, not part of the parse metric, as defined in
SURVEY.md 8(d)).
  value : whole-job Gbp/s with the MSA rows resident in HBM (CUDA events, max over ranks)
  e2e   : the same through drlz_parse_batch with pinned HOST rows in and HOST phrase records out
  roofline / cpu_baseline : see DESIGN.md "Measurement"
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1; the host-side input generator and the CPU baseline are OpenMP code
_world = int(os.environ.get("WORLD_SIZE", "1"))
_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if "--impl" in sys.argv and "reference" in sys.argv:
    os.environ["OMP_NUM_THREADS"] = str(_cores)
elif _world > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, _cores // _world))

METRIC = "rlz_parse_throughput"
UNIT = "Gbp/s"


def workload(rank: int):
    n = int(os.environ.get("DRLZ_BENCH_N", 48_000_000))
    H = int(os.environ.get("DRLZ_BENCH_H", 50))
    name = "configs[1]: chr21-sized %.0f Mbp synthetic reference, %d haplotypes per GPU (0.1%% SNP, 1e-5 indel), haplotype-sharded" % (n / 1e6, H)
    return {"n": n, "H": H, "p_snp": 1e-3, "p_ins": 1e-5, "p_del": 1e-5, "cfg": 2, "h0": rank * H, "name": name}


def gen_rows(orc, wl, d, pinned_array=None, rows=None):
    """rows h0..h0+H-1 as [H, stride] uint8 (stride multiple of 16, >= M+16); returns (array, M, stride)"""
    seed = orc.seed_for(wl["cfg"])
    M = orc.syn_columns(seed, wl["n"], wl["p_ins"])
    stride = (M + 15) // 16 * 16 + 16
    H = wl["H"] if rows is None else rows
    if pinned_array is None:
        arr = np.zeros((H, stride), dtype=np.uint8)
    else:
        arr = pinned_array[: H * stride].reshape(H, stride)
        arr[:, M:] = 0
    import ctypes as C
    L = orc.lib()
    L.syn_rows(seed, d.ctypes.data_as(C.c_void_p), d.size, wl["h0"], H, wl["p_snp"], wl["p_ins"], wl["p_del"],
               arr.ctypes.data_as(C.c_void_p), stride)
    return arr, M, stride


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples in the upper half of the observed range
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(orc, wl, d, rows_arr, M, sample_rows, threads, sa=None):
    """Oracle (C restatement of DistributedRLZ.scala) on the host cores: one haplotype per task, all threads."""
    t0 = time.time()
    if sa is None:
        sa = orc.sa_build(d)
    t_sa = time.time() - t0
    rows = [rows_arr[h, :M] for h in range(sample_rows)]
    orc.encode_rows_mt(d, sa, rows, threads)                    # warm-up (thread pool, page faults)
    res, m, wt, sec = orc.encode_rows_mt(d, sa, rows, threads)
    return {"pairs": res, "m": m, "seconds": sec, "sa_seconds": t_sa, "bases": int(m.sum()), "would_throw": int(wt.sum()), "sa": sa}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    wl = workload(0)
    threads = _cores
    d = orc.syn_reference(orc.seed_for(wl["cfg"]), wl["n"])
    sample = min(wl["H"], int(os.environ.get("DRLZ_BENCH_REF_ROWS", 16)))
    rows_arr, M, stride = gen_rows(orc, wl, d, rows=sample)
    t0 = time.time()
    sa = orc.sa_build(d)
    t_sa = time.time() - t0
    rows = [rows_arr[h, :M] for h in range(sample)]
    times = []
    bases = 0
    for it in range(args.warmup + args.steps):
        res, m, wt, sec = orc.encode_rows_mt(d, sa, rows, threads)
        bases = int(m.sum())
        if it >= args.warmup:
            times.append(sec)
    ms = 1e3 * float(np.mean(times))
    val = bases / (ms * 1e-3) / 1e9
    sample_txt = "%d of the %d haplotypes of the workload per step (%.0f Mbp), parse only; CPU SA build %.1f s not included" % (
        sample, wl["H"], bases / 1e6, t_sa)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic",
           "config": {"workload": wl["name"], "note": "C restatement of DistributedRLZ.scala (oracle/), not the JVM/Spark job: "
                      "no JVM, Spark or radixSA in this image; one haplotype per task over all host threads (local[N] shape)"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(json.dumps(out))
    return 0


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank to the CPUs next to its GPU (NVML affinity) BEFORE pinned host memory is allocated, so the
    staging buffers live on the NUMA node that feeds the GPU's PCIe root port."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc          # input generator + (rank 0, N=1) the cpu_baseline / parity checker
    from pangenspark_b200 import drlz as D

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    ncpu_bound = bind_to_gpu_numa(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = workload(rank)
    seed = orc.seed_for(wl["cfg"])
    d = orc.syn_reference(seed, wl["n"])
    M = orc.syn_columns(seed, wl["n"], wl["p_ins"])
    stride = (M + 15) // 16 * 16 + 16
    H = wl["H"]
    pin = D.PinnedBuffer(H * stride)
    rows_arr, M, stride = gen_rows(orc, wl, d, pinned_array=pin.array)

    ctx = D.Drlz(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for key in ("kmer", "chunk", "group_bytes"):
        v = os.environ.get("DRLZ_" + key.upper())
        if v:
            ctx.set_option(key, int(v))

    # ---- index: built on rank 0, NCCL broadcast to the other ranks ---------------------------------------
    bcast_ms = 0.0
    if rank == 0:
        ctx.build_index(d)            # warm-up build (allocations, first-touch)
        ctx.build_index(d)
        idx_t = ctx.index_timings()
        info = ctx.index_info()
    if world > 1:
        meta = torch.zeros(2, dtype=torch.int64, device="cuda")
        if rank == 0:
            meta[0], meta[1] = info["n"], info["kinfo"]
        dist.broadcast(meta, 0)
        if rank != 0:
            ctx.index_reserve(int(meta[0]), int(meta[1]))
        bufs = [torch.as_tensor(_Raw(p, nb), device="cuda") for p, nb in ctx.index_buffers()]
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in bufs:
            dist.broadcast(b, 0)
        e1.record(); torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            ctx.index_commit()
            info = ctx.index_info()
            idx_t = {"total_ms": 0.0, "pack_ms": 0.0, "sa_ms": 0.0, "table_ms": 0.0, "lcp_ms": 0.0}

    # ---- rows resident in HBM ------------------------------------------------------------------------------
    dev_rows = torch.empty(H * stride, dtype=torch.uint8, device="cuda")
    dev_rows.copy_(torch.from_numpy(rows_arr.reshape(-1)), non_blocking=False)
    row_off = np.arange(H, dtype=np.uint64) * np.uint64(stride)
    cols = np.full(H, M, dtype=np.uint64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local)
    # ---- value: device-resident ----------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            ptr, npairs, m = ctx.parse_batch_device(dev_rows.data_ptr(), row_off, cols)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kt = {}
        ev0.record(stream)
        for _ in range(args.steps):
            ptr, npairs, m = ctx.parse_batch_device(dev_rows.data_ptr(), row_off, cols)
            for k_, v_ in ctx.parse_timings().items():
                kt[k_] = kt.get(k_, 0.0) + v_
        ev1.record(stream)
        barrier()
        ms_dev = ev0.elapsed_time(ev1) / args.steps
    bases = int(m.sum())
    phrases = int(npairs.sum())
    ms_dev = reduce_max(ms_dev)
    total_bases = reduce_sum(float(bases))
    value = total_bases / (ms_dev * 1e-3) / 1e9

    # ---- e2e: pinned host rows in, host phrase records out -----------------------------------------------------
    out_pin = D.PinnedBuffer(16 * (phrases + 1024))
    out_arr = out_pin.array.view(np.int64)
    host_rows = [rows_arr[h, :M] for h in range(H)]
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            res, m2, _ = ctx.parse_batch(host_rows, pairs_cap=phrases + 1024, out=out_arr)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            res, m2, _ = ctx.parse_batch(host_rows, pairs_cap=phrases + 1024, out=out_arr)
        ev1.record(stream)
        barrier()
        wall_e2e = (time.time() - t0) / args.steps * 1e3
        ms_e2e = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop()
    ms_e2e = reduce_max(max(ms_e2e, 0.0))
    e2e_val = total_bases / (ms_e2e * 1e-3) / 1e9
    assert [r.shape[0] for r in res] == npairs.tolist() and (m2 == m).all()

    # ---- roofline of the dominant kernel (per-launch CUDA-event times measured above) ---------------------------
    peak, peak_src = peaks()
    steps = args.steps
    Mtot, mtot = float(H) * M, float(bases)
    per_launch = {   # name: (avg ms per launch, algorithmic bytes per launch)
        "k_pack": (kt["pack_write_ms"] / steps, Mtot + mtot / 4),
        "k_hint+k_diag_w": (kt["prescan_ms"] / steps, mtot / 4),
        "k_spec": (kt["spec_ms"] / steps, mtot / 4),
        "k_fix": (kt["fix_ms"] / max(1.0, kt["fix_rounds"] - steps) if kt["fix_rounds"] > steps else 0.0, 0.0),
        "k_emit": (kt["emit_ms"] / steps, mtot / 4 + 16.0 * phrases),
    }
    dom = max(per_launch, key=lambda k_: per_launch[k_][0])
    dom_ms, dom_bytes = per_launch[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    alg_step = Mtot + mtot / 2 + 16.0 * phrases            # M + m/4 + m/4 + 16 P
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_ms": dom_ms,
                "step": {"algorithmic_bytes": alg_step, "achieved": alg_step / (ms_dev * 1e-3) / 1e9,
                         "frac": alg_step / (ms_dev * 1e-3) / 1e9 / peak},
                "kernel_ms_per_step": {k_: round(v_ / steps, 4) for k_, v_ in kt.items() if k_.endswith("_ms")}}

    # ---- CPU baseline + parity gate (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    parity_rows = 0
    if rank == 0 and world == 1 and not os.environ.get("DRLZ_BENCH_SKIP_CPU"):
        threads = _cores
        sample = min(H, int(os.environ.get("DRLZ_BENCH_REF_ROWS", 16)))
        sa_gpu = ctx.index_get("sa").astype(np.int64)
        cb = cpu_baseline(orc, wl, d, rows_arr, M, sample, threads)
        if not (cb["sa"] == sa_gpu).all():
            print("PARITY FAILURE: suffix array differs from the oracle", file=sys.stderr)
            return 3
        for h in range(sample):
            if orc.lz_bytes(res[h]) != orc.lz_bytes(cb["pairs"][h]):
                print("PARITY FAILURE: phrase records of row %d differ from the oracle" % h, file=sys.stderr)
                return 3
        parity_rows = sample
        cpu = {"value": cb["bases"] / cb["seconds"] / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of the %d haplotypes (%.0f Mbp), parse only, one haplotype per task; CPU SA build %.1f s excluded; "
                         "C restatement of DistributedRLZ.scala, not the JVM job" % (sample, H, cb["bases"] / 1e6, cb["sa_seconds"]),
               "sa_build_s": cb["sa_seconds"]}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
               "data": "synthetic",
               "config": {"workload": wl["name"], "l2": "inputs (%.2f GB of MSA rows per GPU) exceed the 126 MB L2; no explicit flush" % (H * M / 1e9),
                          "bases_per_step": total_bases, "cpus_bound_per_rank": ncpu_bound, "phrases_per_step_rank0": phrases, "kmer": info["k"],
                          "index": "built once outside the step" + (", NCCL broadcast from rank 0" if world > 1 else "")},
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(H * M), "d2h_bytes_per_step": int(16 * phrases),
                       "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e},
               "gpu_launches": int(kt["launches"]),
               "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
               "index_build": {"total_ms": idx_t["total_ms"], "sa_ms": idx_t["sa_ms"], "table_ms": idx_t["table_ms"],
                               "pack_ms": idx_t["pack_ms"], "rounds": info["sa_rounds"], "n": info["n"],
                               "algorithmic_bytes": 9 * info["n"], "bcast_ms": bcast_ms,
                               "ref_mbp_per_s": info["n"] / max(idx_t["total_ms"], 1e-9) / 1e3},
               "parity_checked_rows": parity_rows, "fix_rounds_per_step": kt["fix_rounds"] / steps}
        _emit(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line: str):
    """The driver parses stdout: exactly one JSON line goes to the real stdout (libraries such as NCCL print
    banners to fd 1, so fd 1 is pointed at stderr for the whole run, see main)."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
