#!/usr/bin/env python
"""bench.py -- RLZ-parsed Gbp/s on B200 (BASELINE.json metric) next to the CPU restatement of the reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path (gap strip + 2-bit pack + exact greedy RLZ parse + phrase records) over the
rank's batch of haplotypes.

Primary line (`value`, `e2e`, `roofline`): BASELINE.json configs[1] per GPU (chr21-sized 48 Mbp synthetic reference,
50 haplotypes) -- haplotypes shard across ranks, no data-path collective, weak scaling; for N > 1 the index is built
on rank 0 and broadcast with NCCL (reported separately, as SURVEY.md 8(d) defines the metric on the parse phase).
  value : whole-job Gbp/s with the MSA rows resident in HBM (CUDA events, max over ranks)
  e2e   : the same through drlz_parse_batch with pinned HOST rows in and HOST phrase records out; next to it the raw
          pinned host->device copy bandwidth of the same bytes on the same ranks at the same time (`h2d_ceiling_gbs`):
          what the box's PCIe / host memory gives with nothing of ours involved
  roofline / cpu_baseline : see DESIGN.md "Measurement"
`strong_config3`: BASELINE.json configs[2] -- chr1-sized 249 Mbp reference, 100 haplotypes in all, sharded over the
N ranks (strong scaling), with a `job` figure that includes the index build and its broadcast (the serial term).
`cli_e2e`: the drop-in itself -- MSA rows as files on tmpfs -> bin/drlz -> .lz + gapless FASTA files, wall clock.
Parity gate at every N: every rank compares its (received) suffix array and phrase records of its own rows with the
CPU oracle; a mismatch anywhere makes every rank exit non-zero.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
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


def workload_cfg3(rank: int, world: int):
    """configs[2]: 249 Mbp reference, 100 haplotypes in all; rank r takes the contiguous slice of the basename-sorted list"""
    n = int(os.environ.get("DRLZ_BENCH_CFG3_N", 249_000_000))
    Htot = int(os.environ.get("DRLZ_BENCH_CFG3_H", 100))
    a, b = rank * Htot // world, (rank + 1) * Htot // world
    name = "configs[2]: chr1-sized %.0f Mbp synthetic reference, %d haplotypes in all, sharded over %d GPU(s)" % (n / 1e6, Htot, world)
    return {"n": n, "H": b - a, "Htot": Htot, "p_snp": 1e-3, "p_ins": 1e-5, "p_del": 1e-5, "cfg": 3, "h0": a, "name": name}


def gen_nrun_workload(orc, n, H, nrun):
    """A reference assembly as it comes: ACGT with runs of N -- a telomere, a centromere-sized run of `nrun`
    bases, a few short ones -- and H haplotypes at configs[1] rates whose N runs are the reference's (a panel called
    against the reference has no variant inside an N region).  Returns (reference bytes, rows [H, stride], M, stride)."""
    import ctypes as C
    seed = orc.seed_for(2)
    d = orc.syn_reference(seed, n)
    runs = [(0, min(n // 100, 500_000)), (n // 3, nrun), (n // 2 + 12345, 50_000), (n - n // 50, 30_000), (n // 5, 700)]
    for a, ln in runs:
        d[a:a + ln] = ord("N")
    M = orc.syn_columns(seed, n, 1e-5)
    stride = (M + 15) // 16 * 16 + 16
    rows = np.zeros((H, stride), dtype=np.uint8)
    orc.lib().syn_rows(seed, d.ctypes.data_as(C.c_void_p), d.size, 1, H, 1e-3, 1e-5, 1e-5, rows.ctypes.data_as(C.c_void_p), stride)
    # the generator treats N as a base and mutates it: every haplotype gets the reference's own columns back over the runs
    ref_row = np.zeros((1, stride), dtype=np.uint8)
    orc.lib().syn_rows(seed, d.ctypes.data_as(C.c_void_p), d.size, 0, 1, 0.0, 1e-5, 0.0, ref_row.ctypes.data_as(C.c_void_p), stride)
    is_n = np.flatnonzero(ref_row[0, :M] == ord("N"))
    breaks = np.flatnonzero(np.diff(is_n) > 16)
    starts = np.concatenate([[is_n[0]], is_n[breaks + 1]])
    ends = np.concatenate([is_n[breaks], [is_n[-1]]]) + 1
    for a, b in zip(starts, ends):
        rows[:, a:b] = ref_row[0, a:b]
    return d, rows, M, stride


def alphabet_4bit(job, steps):
    """The generic-alphabet path (4 bits per symbol) on the same scale as the headline: what a reference with N in it --
    every real chromosome -- gets.  Own context; one row checked against the oracle."""
    orc, torch, D = job.orc, job.torch, job.D
    n = int(os.environ.get("DRLZ_BENCH_N", 48_000_000))
    H = int(os.environ.get("DRLZ_BENCH_4BIT_H", 25))
    d, rows, M, stride = gen_nrun_workload(orc, n, H, 3_000_000)
    ctx = D.Drlz(job.local)
    try:
        ctx.build_index(d)
        info, it = ctx.index_info(), ctx.index_timings()
        dev = torch.from_numpy(rows.reshape(-1)).cuda()
        row_off = np.arange(H, dtype=np.uint64) * np.uint64(stride)
        cols = np.full(H, M, dtype=np.uint64)
        for _ in range(3):
            ptr, npairs, m = ctx.parse_batch_device(dev.data_ptr(), row_off, cols)
        ctx.set_option("keep_timings", 1)
        for _ in range(steps):
            ptr, npairs, m = ctx.parse_batch_device(dev.data_ptr(), row_off, cols)
        kt = ctx.parse_timings()
        ctx.set_option("keep_timings", 0)
        pairs = ctx.copy_pairs_to_host(ptr, int(npairs.sum()))
        sa = orc.sa_build(d)
        ok = bool((ctx.index_get("sa").astype(np.int64) == sa).all())
        want, _ = orc.encode_literal(d, sa, orc.strip_gaps(rows[0, :M]))
        ok = ok and orc.lz_bytes(pairs[: int(npairs[0])]) == orc.lz_bytes(want)
        dev_ms = kt["device_ms"] / steps
        return {"value": float(m.sum()) / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": dev_ms, "bits_per_symbol": info["bits"],
                "sigma": info["sigma"], "kmer": info["k"], "rows": H, "index_ms": it["total_ms"], "sa_rounds": info["sa_rounds"],
                "kernel_ms_per_step": {k_: round(v_ / steps, 4) for k_, v_ in kt.items() if k_.endswith("_ms")},
                "parity_checked_rows": 1, "ok": ok,
                "workload": "%.0f Mbp ACGT reference with runs of N (longest 3 Mbp), %d haplotypes at configs[1] rates; device-resident, "
                            "sum of the kernels' event times per step" % (n / 1e6, H)}
    finally:
        ctx.close()


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


def config_of(wl, M):
    """`config` of the JSON line, the same dict in both arms (the driver compares them): the workload and how the timing
    rule about L2 is met; everything else about the run goes into `config_detail`"""
    return {"workload": wl["name"],
            "l2": "the MSA rows of the workload (%.2f GB per GPU) exceed the 126 MB L2; no explicit flush" % (wl["H"] * M / 1e9)}


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
           "config": config_of(wl, M),
           "note": "C restatement of DistributedRLZ.scala (oracle/), not the JVM/Spark job: no JVM, Spark or radixSA in this image; "
                   "one haplotype per task over all host threads (local[N] shape)",
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(json.dumps(out))
    return 0


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class Job:
    """One rank of the benchmark: torch for device memory, streams and NCCL; the library for everything measured."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from oracle import oracle as orc          # input generator + the cpu_baseline / parity checker
        from pangenspark_b200 import drlz as D
        self.torch, self.dist, self.orc, self.D, self.args = torch, dist, orc, D, args
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = D.Drlz(self.local)
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        for key in ("kmer", "chunk", "group_bytes", "pack_variant", "occ_pack", "tickets", "prefetch"):
            v = os.environ.get("DRLZ_" + key.upper())
            if v:
                self.ctx.set_option(key, int(v))
        self.threads = max(1, _cores // self.world)
        self.shm_tag = "drlz_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.getppid() if self.world > 1 else os.getpid())

    # ---- collectives (scalars) ----------------------------------------------------------------------------------
    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def reduce_max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def reduce_min(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    def reduce_sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    # ---- index: built on rank 0, NCCL broadcast to the other ranks -------------------------------------------------
    def index(self, d, warm=True):
        torch, dist, ctx = self.torch, self.dist, self.ctx
        info, idx_t, bcast_ms, wall_ms = None, None, 0.0, 0.0
        if self.rank == 0:
            # the reference sits in pinned memory (drlz_host_alloc), as the rows do: from pageable memory the driver stages
            # the copy and the build waits 3 ms for 48 MB
            pin = self.D.PinnedBuffer(max(1, d.size))
            pin.array[:d.size] = d
            dp = pin.array[:d.size]
            if warm:
                ctx.build_index(dp)           # warm-up build (allocations, first touch)
            t0 = time.time()
            ctx.build_index(dp)
            wall_ms = (time.time() - t0) * 1e3
            pin.free()
            idx_t = ctx.index_timings()
            info = ctx.index_info()
        if self.world > 1:
            meta = torch.zeros(2, dtype=torch.int64, device="cuda")
            if self.rank == 0:
                meta[0], meta[1] = info["n"], info["kinfo"]
            dist.broadcast(meta, 0)
            if self.rank != 0:
                ctx.index_reserve(int(meta[0]), int(meta[1]))
            bufs = [torch.as_tensor(_Raw(p, nb), device="cuda") for p, nb in ctx.index_buffers()]
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b in bufs:
                dist.broadcast(b, 0)
            e1.record(); torch.cuda.synchronize()
            bcast_ms = self.reduce_max(e0.elapsed_time(e1))
            if self.rank != 0:
                ctx.index_commit()
                info = ctx.index_info()
                idx_t = {"total_ms": 0.0, "pack_ms": 0.0, "sa_ms": 0.0, "table_ms": 0.0, "lcp_ms": 0.0}
        return info, idx_t, bcast_ms, wall_ms

    # ---- one workload: device-resident value, end to end, raw copy ceiling ------------------------------------------
    def measure(self, wl, d, steps, warmup, sampler=None, gapless=True):
        torch, ctx, D, orc = self.torch, self.ctx, self.D, self.orc
        seed = orc.seed_for(wl["cfg"])
        M = orc.syn_columns(seed, wl["n"], wl["p_ins"])
        stride = (M + 15) // 16 * 16 + 16
        H = wl["H"]
        t0 = time.time()
        pin = D.PinnedBuffer(max(1, H) * stride)
        rows_arr, M, stride = gen_rows(orc, wl, d, pinned_array=pin.array)
        gen_s = time.time() - t0
        dev_rows = torch.empty(max(1, H) * stride, dtype=torch.uint8, device="cuda")
        pin_t = torch.from_numpy(rows_arr.reshape(-1))
        dev_rows[: H * stride].copy_(pin_t, non_blocking=False)
        row_off = np.arange(H, dtype=np.uint64) * np.uint64(stride)
        cols = np.full(H, M, dtype=np.uint64)
        stream = self.stream

        # value: rows resident in HBM
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                ptr, npairs, m = ctx.parse_batch_device(dev_rows.data_ptr(), row_off, cols)
            self.barrier()
            if sampler:
                sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.set_option("keep_timings", 1)          # the library sums its per-kernel event times over the timed steps
            rows_ptr = dev_rows.data_ptr()
            ev0.record(stream)
            for _ in range(steps):
                ptr, npairs, m = ctx.parse_batch_device(rows_ptr, row_off, cols)
            ev1.record(stream)
            self.barrier()
            kt = ctx.parse_timings()
            ctx.set_option("keep_timings", 0)
            ms_dev = ev0.elapsed_time(ev1) / steps
        bases = int(m.sum())
        phrases = int(npairs.sum())
        ms_dev = self.reduce_max(ms_dev)
        total_bases = self.reduce_sum(float(bases))
        value = total_bases / (ms_dev * 1e-3) / 1e9

        # e2e: pinned host rows in, host phrase records out
        out_pin = D.PinnedBuffer(16 * (phrases + 1024))
        out_arr = out_pin.array.view(np.int64)
        host_rows = [rows_arr[h, :M] for h in range(H)]
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                res, m2, _ = ctx.parse_batch(host_rows, pairs_cap=phrases + 1024, out=out_arr)
            self.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            ev0.record(stream)
            for _ in range(steps):
                res, m2, _ = ctx.parse_batch(host_rows, pairs_cap=phrases + 1024, out=out_arr)
            ev1.record(stream)
            self.barrier()
            wall_e2e = (time.time() - t0) / steps * 1e3
            ms_e2e = ev0.elapsed_time(ev1) / steps
        ms_e2e = self.reduce_max(max(ms_e2e, 0.0))
        e2e_val = total_bases / (ms_e2e * 1e-3) / 1e9
        assert [r.shape[0] for r in res] == npairs.tolist() and (m2 == m).all()

        # e2e_gapless: the same call with the gapless sequences copied back too (what the CLI asks for: DistributedRLZ.scala
        # :66-80 writes them as FASTA) -- the pack kernel's gapless branch and m more bytes device -> host per step
        e2e_gl = None
        if gapless and H and not os.environ.get("DRLZ_BENCH_SKIP_GAPLESS"):
            gl_pin = D.PinnedBuffer(max(1, H) * stride)
            gl_views = [gl_pin.array[h * stride: h * stride + M] for h in range(H)]
            with torch.cuda.stream(stream):
                for _ in range(max(1, warmup - 1)):
                    res_g, m3, gl = ctx.parse_batch(host_rows, want_gapless=True, pairs_cap=phrases + 1024, out=out_arr, gapless_out=gl_views)
                self.barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
                for _ in range(steps):
                    res_g, m3, gl = ctx.parse_batch(host_rows, want_gapless=True, pairs_cap=phrases + 1024, out=out_arr, gapless_out=gl_views)
                ev1.record(stream)
                self.barrier()
                ms_gl = self.reduce_max(ev0.elapsed_time(ev1) / steps)
            assert (m3 == m).all() and gl[0][: int(m[0])].tobytes() == orc.strip_gaps(rows_arr[0, :M]).tobytes()
            e2e_gl = {"value": total_bases / (ms_gl * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_gl, "h2d_bytes_per_step": int(H * M),
                      "d2h_bytes_per_step": int(16 * phrases + bases), "d2h_gbs": self.reduce_sum(16.0 * phrases + bases) / (ms_gl * 1e-3) / 1e9}
            res, m2, _ = ctx.parse_batch(host_rows, pairs_cap=phrases + 1024, out=out_arr)      # `res` views out_arr: restore the plain result
            # what the box gives both directions at once: the rows host -> device on one stream while as many bytes as the
            # gapless sequences go device -> host on another (nothing of the library involved)
            gl_t = torch.from_numpy(gl_pin.array[: H * stride])
            side = torch.cuda.Stream()
            dsrc = torch.empty(int(bases), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            self.barrier()
            ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ev0.record(stream)
            side.wait_event(ev0)
            for _ in range(steps):
                with torch.cuda.stream(stream):
                    dev_rows[: H * stride].copy_(pin_t, non_blocking=True)
                with torch.cuda.stream(side):
                    gl_t[: int(bases)].copy_(dsrc, non_blocking=True)
            ev2.record(side)
            stream.wait_event(ev2)
            ev1.record(stream)
            torch.cuda.synchronize()
            self.barrier()
            ms_dup = self.reduce_max(ev0.elapsed_time(ev1) / steps)
            e2e_gl["duplex_ceiling"] = {"value": total_bases / (ms_dup * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_dup,
                                        "how": "pinned host -> device copy of the rows and device -> host copy of as many bytes as the gapless sequences, at the same time on two streams, every rank at once"}
            e2e_gl["frac_of_duplex_ceiling"] = ms_dup / ms_gl
            del gl_t, dsrc
            del gl, gl_views
            gl_pin.free()

        # the ceiling of that: nothing but the pinned host->device copy of the same rows, all ranks at once
        with torch.cuda.stream(stream):
            dev_rows[: H * stride].copy_(pin_t, non_blocking=True)
            self.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                dev_rows[: H * stride].copy_(pin_t, non_blocking=True)
            ev1.record(stream)
            self.barrier()
            ms_copy = self.reduce_max(ev0.elapsed_time(ev1) / steps)
        total_h2d = self.reduce_sum(float(H) * M)
        total_copy = self.reduce_sum(float(H) * stride)
        if sampler:
            self.clocks = sampler.stop()
        h2d_gbs = total_h2d / (ms_e2e * 1e-3) / 1e9
        ceil_gbs = total_copy / (ms_copy * 1e-3) / 1e9
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(H * M), "d2h_bytes_per_step": int(16 * phrases),
               "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e, "h2d_gbs": h2d_gbs, "h2d_ceiling_gbs": ceil_gbs,
               "frac_of_h2d_ceiling": h2d_gbs / ceil_gbs if ceil_gbs else None,
               "h2d_ceiling_how": "pinned host -> device copy of the same rows on every rank at the same time, nothing else running (whole job, GB/s)"}
        return {"M": M, "stride": stride, "H": H, "rows_arr": rows_arr, "res": res, "m": m, "npairs": npairs, "kt": kt,
                "ms_dev": ms_dev, "value": value, "total_bases": total_bases, "bases": bases, "phrases": phrases, "e2e": e2e, "e2e_gapless": e2e_gl,
                "gen_s": gen_s, "_keep": (pin, out_pin, dev_rows)}

    # ---- parity gate: every rank, its own rows and its own copy of the index -------------------------------------------
    def parity(self, wl, d, ms, sample_rows, want_cpu_baseline):
        """Rank 0 builds the CPU suffix array and leaves it in /dev/shm; every rank compares the suffix array in its
        own GPU memory (built there on rank 0, received over NCCL elsewhere) with it and its phrase records of
        `sample_rows` of its own rows with the oracle's.  Returns (rows checked in all, cpu_baseline | None, ok)."""
        orc, ctx = self.orc, self.ctx
        path = "/dev/shm/%s_cfg%d_sa.npy" % (self.shm_tag, wl["cfg"])
        t_sa = 0.0
        if self.rank == 0:
            t0 = time.time()
            sa = orc.sa_build(d)
            t_sa = time.time() - t0
            if self.world > 1:
                np.save(path, sa)
        self.barrier()
        if self.rank != 0:
            sa = np.load(path)
        ok = True
        sa_gpu = ctx.index_get("sa")
        if sa_gpu.shape != sa.shape or not (sa_gpu == sa).all():
            print("PARITY FAILURE (rank %d): suffix array differs from the oracle" % self.rank, file=sys.stderr)
            ok = False
        del sa_gpu
        H, M = ms["H"], ms["M"]
        sample = min(H, sample_rows)
        # rows spread over the rank's slice (first, last, and evenly in between)
        pick = sorted(set(int(round(i * (H - 1) / max(1, sample - 1))) for i in range(sample))) if sample else []
        rows = [ms["rows_arr"][h, :M] for h in pick]
        cpu = None
        if ok and rows:
            threads = _cores if self.world == 1 else self.threads
            if want_cpu_baseline:
                orc.encode_rows_mt(d, sa, rows, threads)                    # warm-up (thread pool, page faults)
            cres, cm, wt, sec = orc.encode_rows_mt(d, sa, rows, threads)
            for j, h in enumerate(pick):
                if orc.lz_bytes(ms["res"][h]) != orc.lz_bytes(cres[j]) or int(cm[j]) != int(ms["m"][h]):
                    print("PARITY FAILURE (rank %d): phrase records of row %d differ from the oracle" % (self.rank, wl["h0"] + h), file=sys.stderr)
                    ok = False
            if want_cpu_baseline:
                cpu = {"value": float(cm.sum()) / sec / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": "%d of the %d haplotypes (%.0f Mbp), parse only, one haplotype per task; CPU SA build %.1f s excluded; "
                                 "C restatement of DistributedRLZ.scala, not the JVM job" % (len(pick), H, float(cm.sum()) / 1e6, t_sa),
                       "sa_build_s": t_sa}
        self.barrier()
        if self.rank == 0 and self.world > 1:
            try:
                os.remove(path)
            except OSError:
                pass
        all_ok = self.reduce_min(1.0 if ok else 0.0) > 0.5
        checked = int(self.reduce_sum(float(len(pick) if ok else 0)))
        return checked, cpu, all_ok

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def roofline_of(ms, steps, peak, peak_src):
    """`roofline` of the dominant kernel = the one with the longest average launch, measured by the CUDA events the
    library records around every kernel of the step; `kernels` lists all of them (the two longest are within a few
    per cent of each other: k_pack3 moves 82 % of the step's algorithmic bytes, k_spec waits on dependent random loads)."""
    kt, H, M, bases, phrases = ms["kt"], ms["H"], ms["M"], ms["bases"], ms["phrases"]
    Mtot, mtot = float(H) * M, float(bases)
    per_launch = {   # name: (avg ms per launch, algorithmic bytes per launch)
        "k_pack": (kt["pack_write_ms"] / steps, Mtot + mtot / 4),
        "k_hint+k_diag_w": (kt["prescan_ms"] / steps, mtot / 4),
        "k_spec": (kt["spec_ms"] / steps, mtot / 4),
        "k_fix": (kt["fix_ms"] / max(1.0, kt["fix_rounds"] - steps) if kt["fix_rounds"] > steps else 0.0, 0.0),
        "k_emit": (kt["emit_ms"] / steps, 16.0 * phrases),       # a gather of stored phrases: it writes the records, it does not read the haplotypes
    }
    dom = max(per_launch, key=lambda k_: per_launch[k_][0])
    dom_ms, dom_bytes = per_launch[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_src, tj = None, None, {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(dom)
            traffic_src = tj.get("_source", "profiles/traffic.json (hand-kept from the last ncu --set full capture)")
        except Exception:
            traffic = None
    kernels = {}
    for name, (kms, kb) in per_launch.items():
        gbs = kb / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        kernels[name] = {"ms": round(kms, 4), "algorithmic_bytes": kb, "achieved": round(gbs, 1), "frac": round(gbs / peak, 4),
                         "traffic": tj.get(name)}
    alg_step = Mtot + mtot / 2 + 16.0 * phrases            # M + m/4 + m/4 + 16 P
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
            "kernel_ms": dom_ms, "kernels": kernels,
            "kernels_note": "algorithmic bytes per kernel: k_pack M + m/4, pre-scan m/4 (it streams the packed haplotypes), k_emit 16 P; "
                            "k_spec is listed against the same m/4 (the matching stage of SURVEY 8(d); it reads mismatch lists and the index, "
                            "not the haplotypes), which the step total counts once",
            "step": {"algorithmic_bytes": alg_step, "achieved": alg_step / (ms["ms_dev"] * 1e-3) / 1e9,
                     "frac": alg_step / (ms["ms_dev"] * 1e-3) / 1e9 / peak},
            "kernel_ms_per_step": {k_: round(v_ / steps, 4) for k_, v_ in kt.items() if k_.endswith("_ms")}}


def cli_e2e(job, wl, d, ms):
    """The drop-in itself (rank 0, N GPUs through DRLZ_GPUS): rows as files on tmpfs -> bin/drlz -> .lz + FASTA files."""
    orc = job.orc
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    if not os.path.exists(exe) or not os.path.isdir("/dev/shm"):
        return None
    H, M = ms["H"], ms["M"]
    rows = min(H, int(os.environ.get("DRLZ_BENCH_CLI_ROWS", H)))
    tmp = tempfile.mkdtemp(prefix="drlz_cli_", dir="/dev/shm")
    try:
        os.makedirs(os.path.join(tmp, "msa"))
        d.tofile(os.path.join(tmp, "ref.txt"))
        for h in range(rows):
            ms["rows_arr"][h, :M].tofile(os.path.join(tmp, "msa", "H%05d.21" % (wl["h0"] + h)))
        env = dict(os.environ, DRLZ_GPUS="1", DRLZ_DEVICE=str(job.local), DRLZ_STATS=os.path.join(tmp, "stats.json"))
        cmd = [exe, os.path.join(tmp, "radix"), os.path.join(tmp, "ref.txt"), os.path.join(tmp, "msa", "*.21"), "file:///",
               os.path.join(tmp, "out"), os.path.join(tmp, "gapless")]
        bases = float(ms["m"][:rows].sum())

        def clear():
            shutil.rmtree(os.path.join(tmp, "out"), ignore_errors=True)
            shutil.rmtree(os.path.join(tmp, "gapless"), ignore_errors=True)

        # first the same files through the same pipeline with the device left out (DRLZ_IO_PROBE: reads into the pinned
        # slots, writes outputs of the same sizes): what the file system of the box gives this process.  It also is the
        # warm-up of the leg: the first pass over fresh tmpfs pages of a VM is slower than every later one.
        io = None
        for _ in range(2):
            clear()
            penv = dict(env, DRLZ_IO_PROBE="1", DRLZ_STATS=os.path.join(tmp, "probe.json"))
            q = subprocess.run(cmd, env=penv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            if q.returncode == 0:
                ps = json.load(open(os.path.join(tmp, "probe.json")))
                io = {"value": bases / ps["encode_s"] / 1e9, "unit": UNIT, "encode_s": ps["encode_s"], "read_wait_s": ps.get("read_wait_s"),
                      "write_wait_s": ps.get("write_wait_s")}
        clear()
        t0 = time.time()
        p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        wall = time.time() - t0
        if p.returncode != 0:
            return {"error": (p.stderr or p.stdout)[-300:]}
        stats = {}
        try:
            stats = json.load(open(os.path.join(tmp, "stats.json")))
        except Exception:
            pass
        # sampled .lz files against the records of the library call above (themselves gated against the oracle)
        ok = True
        sha = []
        for h in sorted(set([0, rows // 2, rows - 1])):
            b = open(os.path.join(tmp, "out", "H%05d.21.lz" % (wl["h0"] + h)), "rb").read()
            sha.append(hashlib.sha256(b).hexdigest()[:16])
            ok = ok and b == orc.lz_bytes(ms["res"][h])
        fa = open(os.path.join(tmp, "gapless", "part-00000"), "rb").read()
        ok = ok and fa == b">H%05d.21\n" % wl["h0"] + orc.strip_gaps(ms["rows_arr"][0, :M]).tobytes() + b"\n"
        parse_s = stats.get("encode_s", wall)
        # the phrases alone (DRLZ_GAPLESS=0: callers that keep the gapless sequences elsewhere): no FASTA to write
        lz_only, lz_runs = None, []
        for _ in range(3):          # the GPU side of this run is 25 copies of 96 MB: a stall anywhere on the shared box shows (0.07 .. 0.7 s seen)
            clear()
            q = subprocess.run(cmd, env=dict(env, DRLZ_GAPLESS="0", DRLZ_STATS=os.path.join(tmp, "lzonly.json")), stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True, timeout=600)
            if q.returncode != 0:
                continue
            ls = json.load(open(os.path.join(tmp, "lzonly.json")))
            b = open(os.path.join(tmp, "out", "H%05d.21.lz" % (wl["h0"] + rows - 1)), "rb").read()
            ok = ok and b == orc.lz_bytes(ms["res"][rows - 1])
            lz_runs.append(round(bases / ls["encode_s"] / 1e9, 3))
            if lz_only is None or ls["encode_s"] < lz_only["encode_s"]:
                lz_only = {"value": bases / ls["encode_s"] / 1e9, "unit": UNIT, "encode_s": ls["encode_s"], "read_wait_s": ls.get("read_wait_s"),
                           "gpu_wait_s": ls.get("gpu_wait_s"), "write_wait_s": ls.get("write_wait_s")}
        if lz_only:
            lz_only["runs"] = lz_runs
            lz_only["what"] = "best of the runs listed"
        return {"value": bases / parse_s / 1e9, "unit": UNIT, "wall_s": wall, "encode_s": parse_s, "rows": rows, "files_ok": bool(ok),
                "lz_sha256_16": sha, "input_gbs": rows * M / parse_s / 1e9, "stats": stats,
                "io_ceiling": io, "frac_of_io_ceiling": (bases / parse_s / 1e9 / io["value"]) if io else None, "lz_only": lz_only,
                "what": "rows as files on tmpfs -> bin/drlz (one GPU) -> .lz and gapless FASTA part files on tmpfs; `value` = bases / "
                        "seconds of the encode phase (read + copy + parse + write, index build excluded as for `e2e`); wall_s = whole process; "
                        "io_ceiling = the same files through the same pipeline with the device left out (DRLZ_IO_PROBE=1; run first, twice, "
                        "also the warm-up of the tmpfs pages); lz_only = the same run with DRLZ_GAPLESS=0 (no FASTA written)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args):
    job = Job(args)
    orc, torch = job.orc, job.torch
    rank, world = job.rank, job.world
    peak, peak_src = peaks()
    steps = args.steps

    # ================= primary: configs[1] per GPU =====================================================================
    wl = workload(rank)
    d = orc.syn_reference(orc.seed_for(wl["cfg"]), wl["n"])
    info, idx_t, bcast_ms, idx_wall = job.index(d)
    ms = job.measure(wl, d, steps, args.warmup, sampler=ClockSampler(job.local))
    clocks = job.clocks
    roofline = roofline_of(ms, steps, peak, peak_src)
    skip_cpu = bool(os.environ.get("DRLZ_BENCH_SKIP_CPU"))
    parity_rows, cpu, ok = 0, None, True
    if not skip_cpu:
        sample = int(os.environ.get("DRLZ_BENCH_REF_ROWS", 16 if world == 1 else 2))
        parity_rows, cpu, ok = job.parity(wl, d, ms, sample, want_cpu_baseline=(world == 1))
        if not ok:
            job.close()
            return 3
    cli = None
    if rank == 0 and not os.environ.get("DRLZ_BENCH_SKIP_CLI"):
        try:
            cli = cli_e2e(job, wl, d, ms)
        except Exception as ex:       # the CLI leg never takes the kernel numbers down with it
            cli = {"error": repr(ex)[:300]}
    a4 = None
    if rank == 0 and world == 1 and not os.environ.get("DRLZ_BENCH_SKIP_4BIT"):
        try:
            a4 = alphabet_4bit(job, max(3, steps // 2))
        except Exception as ex:
            a4 = {"error": repr(ex)[:300]}
    job.barrier()
    primary_kt = ms["kt"]
    H, M = ms["H"], ms["M"]
    out = {"metric": METRIC, "value": ms["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
           "ms_per_step": ms["ms_dev"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
           "data": "synthetic",
           "config": config_of(wl, M),
           "config_detail": {"bases_per_step": ms["total_bases"], "phrases_per_step_rank0": ms["phrases"], "kmer": info["k"],
                             "index": "built once outside the step" + (", NCCL broadcast from rank 0" if world > 1 else "")},
           "e2e": ms["e2e"], "e2e_gapless": ms["e2e_gapless"], "alphabet_4bit": a4, "gpu_launches": int(primary_kt["launches"]),
           "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
           "index_build": {"total_ms": idx_t["total_ms"], "wall_ms": idx_wall, "sa_ms": idx_t["sa_ms"], "table_ms": idx_t["table_ms"],
                           "pack_ms": idx_t["pack_ms"], "lcp_ms": idx_t["lcp_ms"], "rounds": info["sa_rounds"], "n": info["n"],
                           "algorithmic_bytes": 9 * info["n"], "bcast_ms": bcast_ms,
                           "achieved_gbs": 9 * info["n"] / max(idx_t["total_ms"], 1e-9) / 1e6,
                           "frac": 9 * info["n"] / max(idx_t["total_ms"], 1e-9) / 1e6 / peak,
                           "includes": "host->device copy of the reference (pinned memory), alphabet scan, pack, suffix array, k-mer table, LCP, uniqueness array + block maxima; the inverse suffix array only when doubling rounds are needed (rounds > 0)",
                           "ref_mbp_per_s": info["n"] / max(idx_t["total_ms"], 1e-9) / 1e3},
           "parity_checked_rows": parity_rows, "fix_rounds_per_step": primary_kt["fix_rounds"] / steps,
           "cli_e2e": cli}
    del ms, d
    import gc
    gc.collect()

    # ================= configs[2]: chr1-sized reference, 100 haplotypes in all, strong scaling ======================
    if not os.environ.get("DRLZ_BENCH_SKIP_CFG3"):
        w3 = workload_cfg3(rank, world)
        d3 = orc.syn_reference(orc.seed_for(3), w3["n"])
        info3, idx3, bcast3, wall3 = job.index(d3, warm=False)
        s3 = max(1, min(steps, int(os.environ.get("DRLZ_BENCH_CFG3_STEPS", 3))))
        m3 = job.measure(w3, d3, s3, 3, gapless=False)
        chk3, _, ok3 = (0, None, True) if skip_cpu else job.parity(w3, d3, m3, 2, want_cpu_baseline=False)
        if not ok3:
            job.close()
            return 3
        job_ms = job.reduce_max(wall3) + bcast3 + m3["e2e"]["ms_per_step"]      # first (and only) build of a real job: allocations included
        out["strong_config3"] = {
            "workload": w3["name"], "scaling": "strong", "rows_per_rank_max": int(job.reduce_max(float(w3["H"]))),
            "value": m3["value"], "unit": UNIT, "ms_per_step": m3["ms_dev"], "steps": s3, "warmup": 3,
            "e2e": m3["e2e"], "bases_per_step": m3["total_bases"],
            "index_build_ms": job.reduce_max(idx3["total_ms"]), "index_first_call_wall_ms": job.reduce_max(wall3), "bcast_ms": bcast3,
            "job": {"value": m3["total_bases"] / (job_ms * 1e-3) / 1e9, "unit": UNIT, "ms": job_ms,
                    "what": "index build on rank 0 + NCCL broadcast + one end-to-end pass over all rows (the serial terms included)"},
            "roofline_step_frac": roofline_of(m3, s3, peak, peak_src)["step"]["frac"],
            "kernel_ms_per_step": {k_: round(v_ / s3, 4) for k_, v_ in m3["kt"].items() if k_.endswith("_ms")},
            "parity_checked_rows": chk3, "kmer": info3["k"]}

    if rank == 0:
        _emit(json.dumps(out))
    job.close()
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
