#!/usr/bin/env python
"""Headline benchmark: point-vs-center feature evaluations per second (BASELINE.json metric) on the
C2 workload (100 k synthetic 1.5 kb 16S-like sequences, --id 0.97 --kmer 4), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is 1000 Trainer::get_close-shaped scans (1 center x all n points each: every live feature +
the GLM decision + argmax / mark reduction) issued the way accumulate() issues them: a DEPENDENT
chain, one launch per scan, each behind the one before it (programmatic dependent launch).
  value  : evals/s of that chain with the histograms resident in HBM.  The batch is stored R times
           (R*29 MB > 2x the 126 MB L2) and consecutive scans rotate through the replicas, so every scan
           streams its rows from HBM ("inputs larger than L2").
  independent_scans: the same scans as independent work (S of them share one launch) -- a micro-benchmark
           of the kernel; Phase A never issues that shape.
  phase_a_in_product: Phase A of bin/meshclust on the full C2 input (one persistent kernel: range, scan,
           exchange, mean, next center on the device), evals/s and fraction of the HBM roofline.
  e2e    : the same step through host buffers on every rank: everything the step's scans read (the R replicas,
           larger than L2) goes up from pinned host memory first (mc_load_histograms), the dependent scans run,
           summaries and marks come back to the host, all inside the timed region (wall clock, max over ranks).  e2e.one_upload_per_10_scans is
           the round-1 form (mc_scan_host: S scans per upload, chunked upload overlapped with the scans).
  --gpus N (torchrun, one process per GPU): weak scaling: every rank holds all N*n points and evaluates
           its n of every scan; the summaries cross GPUs through NVLink peer inboxes (CUDA IPC) on a second
           stream; value = evals of all ranks / max-over-ranks time.  extra.c4_shape_scan is the 1 M-point
           shape strong-scaled.  DESIGN.md section 5.
  roofline: the scan kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json);
           algorithmic bytes per eval = 4^k + 33.
  cpu_baseline: the compiled unmodified reference (oracle/_ref/libmcref.so, Feature::compute +
           GLM exactly as Trainer::get_close runs them, OpenMP over points like the reference) on a
           bounded sample of the same workload, all host threads.
--impl reference prints the reference CPU arm alone in the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "feature_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "c2"
K = 4
S_CENTERS = 10            # scans per enqueue call (and per launch of the independent-scan micro-benchmark)
CALLS_PER_STEP = 100      # enqueue calls per step: a step is 1000 scans
FALLBACK_HBM_GBS = 6650.0


_JSON_FD = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_workload(seed_shift: int = 0):
    from meshclust_b200 import synth
    c = synth.CONFIGS[WORKLOAD]
    t0 = time.time()
    letters, offs, tmpl = synth.generate(c.n, c.templates, c.length, c.mu, c.seed + 1000 * seed_shift, c.related)
    log(f"[bench] generated {WORKLOAD}: n={c.n} bases={offs[-1]} in {time.time() - t0:.1f}s")
    return c, letters, offs, tmpl


def fit_model(ctx, n, tmpl, rng):
    """Bounds from sampled pairs + least-squares GLM on template labels (a stand-in for
    Trainer::train, whose alignment-labelled sampling is host control, not the measured path)."""
    m = 3000
    a = rng.integers(0, n, m).astype(np.int32)
    ntem = int(tmpl.max()) + 1
    b = np.where(rng.random(m) < 0.5, (a + ntem * rng.integers(1, 50, m)) % n, rng.integers(0, n, m)).astype(np.int32)
    raw, _ = ctx.pair_features(a, b)
    mins, maxs = raw.min(0), np.maximum(raw.max(0), np.finfo(float).tiny)
    ctx.set_model(mins, maxs, np.array([0.0, 1.0, 1.0, 1.0, 1.0]), 4)
    _, _, _, feats = ctx.pair_classify(a, b)
    X = np.concatenate([np.ones((m, 1)), feats], 1)
    y = np.where(tmpl[a] == tmpl[b], 1.0, -1.0)
    w = np.linalg.lstsq(X, y, rcond=None)[0]
    acc = float((np.sign(X @ w) == y).mean())
    return mins, maxs, w, acc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        # NVML in-process (a sample every 5 ms, the first one at once); the nvidia-smi loop of the recipe as fall-back
        # (its start-up alone is ~0.1 s: a 0.14 s timed region saw 1 to 13 samples)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.nvml = (pynvml, h)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.samples, self.stop_flag = [], False
            # the timed loop is a tight sequence of short ctypes calls: with the default 5 ms switch interval the
            # sampler thread gets the interpreter a few times per 100 ms only
            # (one process per GPU under torchrun: the loop there spends longer inside each call and the sampler was
            # never starved -- 19-21 samples -- so the interval is left alone)
            self.switch_interval = sys.getswitchinterval()
            if int(os.environ.get("WORLD_SIZE", "1")) == 1:
                sys.setswitchinterval(0.0005)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _nvml_index(self):
        # NVML numbers the physical GPUs; CUDA_VISIBLE_DEVICES (indices) maps the local rank onto them
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def _poll(self):
        pynvml, h = self.nvml
        while not self.stop_flag:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                mx = self.max_mhz
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if getattr(self, "nvml", None):
            pynvml, _ = self.nvml
            self.stop_flag = True
            self.t.join(timeout=1)
            sys.setswitchinterval(self.switch_interval)
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            reasons = sorted(k for k, b in bits.items() if any(rs & b for _, _, rs in self.samples))
            sm = [x[0] for x in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(x[1] for x in self.samples)) if sm else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


# ------------------------------------------------------------------------------------------------
# reference CPU arm (also the cpu_baseline of our arm)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(hist, lens, mins, maxs, w, centers, target_s: float, steps: int | None = None, warmup: int = 0):
    """evals/s of the compiled reference's get_close (Feature::compute + GLM + argmax/mark reduction,
    OpenMP over points) on a bounded sample; the points exist as DivergencePoint objects beforehand,
    like in the reference.  Returns (value, cores, kind, sample description, ms_per_step)."""
    import _oracle
    n = hist.shape[0]
    ns = min(n, 20000)
    sub, sublen = np.ascontiguousarray(hist[:ns]), np.ascontiguousarray(lens[:ns])
    centers = [int(c) % ns for c in centers]
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)   # the reference prints "Adding combo ..." on every Feature set-up
    try:
        if _oracle.have_ref():
            lib, kind = _oracle.ref(), "reference"
            lib.lib.ref_set_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
            cores = int(lib.lib.ref_max_threads())
            ps = lib.pointset(sub, sublen)
            lib.pointset_set_model(ps, mins, maxs, 4)
            flags = np.zeros(ns, np.uint8)
            one_pass = lambda c: lib.pointset_scan(ps, c, w, flags)
        else:   # the reference could not be compiled where this snapshot was built: oracle port
            lib, kind = _oracle.oracle(), "port"
            cores = os.cpu_count() or 1
            ps = None
            one_pass = lambda c: lib.scan(sub, sublen, sub[c], int(sublen[c]), mins, maxs, w, 4)
        one_pass(centers[0])
        t0 = time.perf_counter()
        one_pass(centers[0])
        one = time.perf_counter() - t0
        if steps is None:
            calls = max(2, int(target_s / max(one, 1e-4)))
            t0 = time.perf_counter()
            for i in range(calls):
                one_pass(centers[i % len(centers)])
            el = time.perf_counter() - t0
            value = calls * ns / el
            ms_step = el / calls * 1e3
            sample = f"{calls} get_close passes over the first {ns} points of {WORKLOAD} ({el:.1f} s)"
        else:
            per_step = max(1, min(S_CENTERS, int(target_s / max(one, 1e-4) / max(steps, 1))))
            for i in range(warmup):
                one_pass(centers[0])
            t0 = time.perf_counter()
            for s in range(steps):
                for j in range(per_step):
                    one_pass(centers[(s * per_step + j) % len(centers)])
            el = time.perf_counter() - t0
            value = steps * per_step * ns / el
            ms_step = el / steps * 1e3
            sample = f"each step = {per_step} get_close passes over the first {ns} points of {WORKLOAD}"
        if ps is not None:
            lib.pointset_destroy(ps)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return value, cores, kind, sample, ms_step


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)     # a step is 1000 scans (~5.5 ms): 20 steps = 0.11 s of timed region
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C4-shape probe and the CLI leg")
    ap.add_argument("--scaling-only", action="store_true", help="development: skip the e2e and CPU-baseline legs too")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    # the contract is ONE JSON line on stdout: libraries that print banners to fd 1 (NCCL's version line)
    # are sent to stderr, the JSON line goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from meshclust_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: meshclust_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg, letters, offs, tmpl = make_workload(seed_shift=rank)
    n = cfg.n
    nbins = 4 ** K
    rng = np.random.default_rng(7 + rank)
    ctx = api.Context(local_rank)

    # ---- stage 1 once (also reported): letters -> histograms on the GPU
    t0 = time.perf_counter()
    hist, mx = ctx.kmer_histograms_host(letters, offs, K, 1)
    t_hist = time.perf_counter() - t0
    lens = np.diff(offs).astype(np.uint64)
    assert mx <= 255
    mins, maxs, w, acc = fit_model(ctx, n, tmpl, rng)
    log(f"[bench] rank {rank}: histograms {t_hist * 1e3:.0f} ms (host->host), model acc {acc:.3f}")

    # ---- N > 1: every rank holds the whole histogram matrix (all-gathered once over NCCL, the
    # "centers are broadcast" of SURVEY 8(e) paid once instead of per scan); scan work and alive
    # flags are sharded block-interleaved (blocks of ~256 KB of consecutive rows go round-robin to the
    # ranks): of the N = world*n points of a scan every rank evaluates n
    exchange = "none"
    N = n * world
    if world > 1:
        t_loc = torch.from_numpy(hist).cuda()
        t_all = torch.empty((N, nbins), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(t_all, t_loc)
        l_loc = torch.from_numpy(lens.astype(np.int64)).cuda()
        l_all = torch.empty(N, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(l_all, l_loc)
        hist_all = t_all.cpu().numpy()
        lens_all = l_all.cpu().numpy().astype(np.uint64)
        del t_loc, t_all, l_loc, l_all
        model = [mins, maxs, w]
        dist.broadcast_object_list(model, src=0)   # one classifier for the job
        mins, maxs, w = model
    else:
        hist_all, lens_all = hist, lens

    # ---- resident data: R replicas so that consecutive launches never re-read L2-resident rows
    row_bytes = nbins + 24
    R = int(np.ceil(2.2 * 126e6 / (n * row_bytes)))
    big = np.ascontiguousarray(np.tile(hist_all, (R, 1)))
    biglens = np.tile(lens_all, R)
    ctx.load_histograms(big, biglens, K)
    ctx.set_model(mins, maxs, w, 4)
    del big
    S = S_CENTERS
    centers_global = np.random.default_rng(1234).integers(0, N, 64)   # the same centers on every rank
    centers_local = centers_global if world == 1 else rng.integers(0, n, 64)

    # arguments of the enqueue calls, built once outside the timed region: call c scans S centers, scan s of it
    # against replica (c*S + s) % R; the sequence repeats after 64*R calls
    NCALLARGS = 64 * R
    call_args = []
    for c in range(NCALLARGS):
        reps = [(c * S + s) % R for s in range(S)]
        cr = np.array([r * N + centers_global[(c * S + s) % 64] for s, r in enumerate(reps)], np.int64)
        lo = np.array([r * N for r in reps], np.int64)
        call_args.append((cr, lo, np.ascontiguousarray(lo + N - 1)))

    # N > 1: the scans run back to back on the scan stream; a second stream folds each scan's CTA partials,
    # stores the record into all ranks' inboxes over NVLink peer memory (CUDA IPC between the processes) and
    # combines the world records per scan on the device, behind the scans.
    # Fallback when peer memory cannot be opened: device fold + NCCL all-gather.
    if world > 1:
        try:
            handle = ctx.comm_init(rank, world)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            ctx.comm_connect(handles)
            ok = torch.ones(1, device="cuda")
        except Exception as e:   # noqa: BLE001
            log(f"[bench] rank {rank}: peer inboxes unavailable ({e}); falling back to NCCL all-gather")
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        exchange = "peer_inbox" if float(ok.item()) > 0 else "nccl_allgather"
        if exchange == "nccl_allgather":
            from meshclust_b200 import sharding
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            records = torch.zeros((S, 4), dtype=torch.int64, device="cuda")
    last_results = [None]
    inflight = []   # calls whose summaries have not been collected yet (at most 2 + the one being enqueued)
    call_no = [0]

    def one_call(mode):
        """S scans: mode = MC_SCAN_CHAIN (one launch per scan, each behind the previous one: what Phase A issues)
        or MC_SCAN_KEEP (independent scans, one launch carries all S)."""
        c = call_no[0]
        call_no[0] += 1
        cr, lo, hi = call_args[c % NCALLARGS]
        if world == 1:
            ctx.scan_enqueue_many(cr, lo, hi, mode, 0)
        elif exchange == "nccl_allgather":
            ctx.scan_enqueue_many(cr, lo + rank * n, lo + (rank + 1) * n - 1, mode, 0)
            ctx.scan_fold_dev(0, S, records.data_ptr())
            last_results[0] = sharding.combine_scan_records(records, 0)
        else:
            # software pipeline, one C-ABI call per S scans: enqueue them (+ fold / send / combine on the exchange
            # stream) and collect the summaries of the call before the previous one
            inflight.append(c)
            if len(inflight) > 2:
                old = inflight.pop(0)
                last_results[0] = ctx.scan_sharded_burst(cr, lo, hi, mode, (c % 3) * 16, (old % 3) * 16, S)
            else:
                ctx.scan_sharded_burst(cr, lo, hi, mode, (c % 3) * 16, 0, 0)

    def drain():
        if world > 1 and exchange == "peer_inbox":
            e = np.zeros(0, np.int64)
            while inflight:
                old = inflight.pop(0)
                last_results[0] = ctx.scan_sharded_burst(e, e, e, 0, 0, (old % 3) * 16, S)

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def timed(ncalls_warm, ncalls, mode, sampler=None):
        """device time (CUDA events on the library's stream) and wall time of `ncalls` enqueue calls, max over ranks"""
        for _ in range(ncalls_warm):
            one_call(mode)
        drain()
        ctx.sync()
        l0 = ctx.launches
        if sampler:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        ev0.record(stream)
        for _ in range(ncalls):
            one_call(mode)
        drain()
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - tw) * 1e3
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        dev_ms = ev0.elapsed_time(ev1)
        # device time when there is no exchange; wall (barrier + sync bracketed) when there is one
        total_ms = dev_ms if world == 1 else wall_ms
        if world > 1:
            tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = float(tt.item())
        return total_ms, dev_ms, ctx.launches - l0, clocks

    def check_results():
        if world == 1:
            res = ctx.scan_collect(0, S)
            assert all(r[0] == n for r in res), "scan did not evaluate every point"
        else:
            res = last_results[0]
            assert all(r[0] == N for r in res), f"sharded scan did not evaluate every point: {res[:2]}"
        return res

    # ================= headline: the dependent chain =================
    total_ms, dev_ms, gpu_launches, clocks = timed(args.warmup * CALLS_PER_STEP, args.steps * CALLS_PER_STEP, api.MC_SCAN_CHAIN,
                                                   ClockSampler(local_rank))
    results = check_results()
    if world > 1 and exchange == "peer_inbox":
        # self-check of the exchange: this rank also holds all rows, so one un-sharded scan over the
        # whole replica must give exactly the exchanged summary
        cr, lo, hi = call_args[(call_no[0] - 1) % NCALLARGS]
        ctx.scan_enqueue(int(cr[0]), int(lo[0]), int(hi[0]), False, 100)
        whole = ctx.scan_collect(100, 1)[0]
        assert whole == results[0], f"sharded summary {results[0]} != single-GPU summary {whole}"
    ms_per_step = total_ms / args.steps
    scans_per_step = CALLS_PER_STEP * S
    evals_per_step = scans_per_step * n * world
    value = evals_per_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (scan): one scan per launch in the chain.  launch_us = device time of
    # the timed region (CUDA events on the stream the kernel is launched on) / launches; bytes = n x (4^k + 33)
    peak, peak_src = measured_peak()
    bytes_per_scan = n * (nbins + 33)
    launch_us = dev_ms * 1e3 / (args.steps * scans_per_step)
    achieved = bytes_per_scan / (launch_us * 1e-6) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))[f"c2_{nbins}"]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": f"scan_tma_kernel<1,{nbins}>", "launch_us": round(launch_us, 3), "scans_per_launch": 1,
                "algorithmic_bytes_per_launch": bytes_per_scan,
                "note": "dependent chain: one launch per scan, chained by programmatic dependent launch; launch_us = CUDA-event time of "
                        "the timed region / launches in it; traffic = dram bytes of one launch from the committed ncu capture (profiles/)"}

    # ================= the same scans as independent work: S of them share a launch =================
    ind_calls = 200
    ind_ms, ind_dev_ms, ind_launches, _ = timed(20, ind_calls, api.MC_SCAN_KEEP)
    check_results()
    ind_us_scan = ind_dev_ms * 1e3 / (ind_calls * S)
    independent = {"value": ind_calls * S * n * world / (ind_ms * 1e-3), "unit": UNIT, "us_per_scan_on_device": round(ind_us_scan, 3),
                   "scans_per_launch": S, "launches": int(ind_launches),
                   "achieved_GBs": round(bytes_per_scan / (ind_us_scan * 1e-6) / 1e9, 1),
                   "frac_of_peak": round(bytes_per_scan / (ind_us_scan * 1e-6) / 1e9 / peak, 4),
                   "note": "micro-benchmark: scans that remove nothing are independent and share a launch (blockIdx.y = scan); "
                           "Phase A never issues this shape"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic",
           "config": {"workload": f"{WORKLOAD}: 100k synthetic 1.5 kb 16S-like sequences, --id 0.97 --kmer 4",
                      "step": f"{scans_per_step} dependent get_close scans (1 center x all points each), one launch per scan",
                      "points_per_gpu": n, "bins": nbins, "scans_per_step": scans_per_step, "evals_per_step": evals_per_step,
                      "l2": f"inputs larger than L2: {R} replicas of the batch ({R * n * row_bytes / 1e6:.0f} MB), consecutive scans rotate through them",
                      "model": "4 features, bounds from 3000 sampled pairs, least-squares GLM", "parallelism": f"points sharded x{world}", "exchange": exchange},
           "clocks": clocks, "gpu_launches": int(gpu_launches), "roofline": roofline, "independent_scans": independent}

    # ================= e2e: the host-buffer C-ABI call on every rank's own points =================
    if not args.scaling_only:
        hp = torch.empty((n, nbins), dtype=torch.uint8, pin_memory=True)
        hp.numpy()[:] = hist
        lp = torch.empty(n, dtype=torch.int64, pin_memory=True)
        lp.numpy()[:] = lens.astype(np.int64)
        marks = torch.empty((S, n), dtype=torch.uint8, pin_memory=True)
        ctx2 = api.Context(local_rank)
        ctx2.set_model(mins, maxs, w, 4)
        cr = centers_local[:S].astype(np.int64)
        hnp, lnp, mnp = hp.numpy(), lp.numpy().view(np.uint64), marks.numpy()
        for _ in range(3):
            ctx2.scan_host(hnp, lnp, K, cr, mnp)
        reps = 30
        summ = torch.zeros((S, 4), dtype=torch.float64)
        summ_dev = torch.zeros((S, 4), dtype=torch.float64, device="cuda")
        gath_dev = torch.zeros((world * S, 4), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            res = ctx2.scan_host(hnp, lnp, K, cr, mnp)
            if world > 1:
                # the job's summaries: counts add up, arg-max over the ranks (what Trainer::get_close reduces)
                for i, r in enumerate(res):
                    summ[i, 0], summ[i, 1], summ[i, 2], summ[i, 3] = r[0], r[1], r[3], (r[2] + rank * n if r[2] >= 0 else -1)
                summ_dev.copy_(summ)
                dist.all_gather_into_tensor(gath_dev, summ_dev)
                gath = gath_dev.cpu()   # every rank has the job's S summaries: evals / positives add up, arg-max over the ranks
        e2e_s = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        # bytes mc_scan_host moves per call: histograms + lengths + copies of the S center rows up; S mark
        # arrays + the CTA partial records of S scans x 4 chunks (160 x 32 B each) down
        e2e_small = {"value": world * S * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(world * (n * nbins + n * 8 + S * (nbins + 8))),
                     "d2h_bytes_per_step": int(world * (S * n + 4 * S * 160 * 32)), "ms_per_step": e2e_s * 1e3,
                     "step": f"{S} get_close scans of every rank's {n} points from pinned host memory (one upload per {S} scans)",
                     "call": "mc_scan_host on every rank (pinned host histograms -> chunked upload overlapped with S scans per chunk -> marks + summaries on the host); "
                             "max over ranks of the time per call"}
        # ---- the bench's own step through host buffers.  Exactly the scans of the headline -- the same dependent chain,
        # rotating through the same R replicas so that every scan streams its rows from HBM -- but everything they read
        # goes up from pinned host memory inside the timed region first (mc_load_histograms: R x n rows + lengths), and
        # the summaries of the last call + the marks of a final mc_scan come back.  At N > 1 every rank does this
        # with its own n points and the job's summaries are all-gathered, as above.
        bigp = torch.empty((R * n, nbins), dtype=torch.uint8, pin_memory=True)
        bigp.numpy()[:] = np.tile(hist, (R, 1))
        biglp = torch.empty(R * n, dtype=torch.int64, pin_memory=True)
        biglp.numpy()[:] = np.tile(lens.astype(np.int64), R)
        bnp, blnp = bigp.numpy(), biglp.numpy().view(np.uint64)
        chain_args = []
        for c in range(CALLS_PER_STEP):
            reps_ = [(c * S + s_) % R for s_ in range(S)]
            chain_args.append((np.array([r_ * n + centers_local[(c * S + s_) % 64] for s_, r_ in enumerate(reps_)], np.int64),
                               np.array([r_ * n for r_ in reps_], np.int64), np.array([r_ * n + n - 1 for r_ in reps_], np.int64)))

        def chain_step():
            ctx2.load_histograms(bnp, blnp, K)
            for cr_, lo_, hi_ in chain_args:
                ctx2.scan_enqueue_many(cr_, lo_, hi_, api.MC_SCAN_CHAIN, 0)
            res = ctx2.scan_collect(0, S)
            last, mk = ctx2.scan(int(chain_args[0][0][0]), 0, n - 1)
            return res + [last.as_tuple()], mk

        for _ in range(2):
            chain_step()
        reps2 = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps2):
            res, mk = chain_step()
            if world > 1:
                for i, r in enumerate(res[:S]):
                    summ[i, 0], summ[i, 1], summ[i, 2], summ[i, 3] = r[0], r[1], r[3], (r[2] + rank * n if r[2] >= 0 else -1)
                summ_dev.copy_(summ)
                dist.all_gather_into_tensor(gath_dev, summ_dev)
                gath = gath_dev.cpu()
        chain_s = (time.perf_counter() - t0) / reps2
        if world > 1:
            tt = torch.tensor([chain_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            chain_s = float(tt.item())
        assert all(r[0] == n for r in res[:S]), "e2e chain: a scan did not evaluate every point"
        nscans = CALLS_PER_STEP * S + 1
        out["e2e"] = {"value": world * nscans * n / chain_s, "unit": UNIT,
                      "h2d_bytes_per_step": int(world * (R * n * nbins + R * n * 8 + CALLS_PER_STEP * S * 24 + 24)),
                      "d2h_bytes_per_step": int(world * (n + (S + 1) * 160 * 32)), "ms_per_step": chain_s * 1e3,
                      "step": f"the headline's step with its input uploaded inside the timed region: {R} x {n} rows per rank ({R * n * nbins / 1e6:.0f} MB, "
                              f"larger than L2) from pinned host memory, then the {nscans} dependent get_close scans rotating through them, "
                              "summaries + marks read back",
                      "call": "mc_load_histograms + mc_scan_enqueue_many (MC_SCAN_CHAIN) x %d + mc_scan_collect + mc_scan, wall clock, max over ranks" % CALLS_PER_STEP,
                      "one_upload_per_10_scans": e2e_small}
        del bigp, biglp
        if rank == 0:
            # parity spot-check of what was just timed (oracle as the checker only)
            import _oracle
            s_o, f0_o, fl_o = _oracle.oracle().scan(hist[:4000], lens[:4000], hist[cr[0]], int(lens[cr[0]]), mins, maxs, w, 4)
            near = np.abs(s_o) < 1e-9
            assert np.array_equal(mnp[0, :4000][~near], fl_o[~near]), "bench: scan marks differ from the oracle"
        ctx2.close()

    if rank == 0 and not args.scaling_only:
        # ---- CPU baseline beside it: compiled reference on a bounded sample, all host threads
        v, cores, kind, sample, _ = cpu_reference_rate(hist, lens, mins, maxs, w, centers_local[:8], target_s=12.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        out["stage1_histograms_ms_host_to_host"] = t_hist * 1e3

    # ---- extra: the HBM-bound regime, the C4 shape (1 M points x 1024 bins, 1 GB > L2), STRONG-scaled: the same
    # 1 M rows at every N, every rank scans its share of each scan
    if not args.no_extra and not args.scaling_only:
        try:
            c4 = c4_shape_probe(api, local_rank, torch, dist, peak, rank, world, exchange)
        except Exception as e:   # never lose the headline line to the probe
            c4 = {"error": str(e)[:200]}
        if rank == 0:
            out["extra"] = {"c4_shape_scan": c4}

    if rank == 0:
        # ---- second half of the metric: sequences clustered per second, FASTA in -> CLSTR out, through
        # the drop-in CLI (bin/meshclust) on the full C2 input; the reference CLI beside it on a bounded
        # sample (its training sorts are O(150 n log n 4^k) on the CPU).  At N > 1 the same input goes through
        # `bin/meshclust --gpus N` (one process driving N GPUs); this rank's own CUDA context stays alive next to it
        if not args.no_extra and not args.scaling_only:
            try:
                out["seqs_clustered"] = cli_leg(letters, offs, tmpl, cfg, world)
                pa = out["seqs_clustered"].pop("phase_a", None)
                if pa:
                    # Phase A as the product runs it (one persistent kernel: range, scan, exchange, mean, next center)
                    pa["achieved_GBs"] = round(pa["evals"] * (nbins + 33) / pa["device_s"] / 1e9, 1)
                    pa["frac_of_peak"] = round(pa["achieved_GBs"] / peak, 4)
                    pa["note"] = ("bin/meshclust, accumulate() on the device: evals = alive points actually evaluated; the kernel compacts "
                                  "the rows whenever an eighth has left, so it streams at most 8/7 of them; algorithmic bytes = evals x "
                                  "(4^k + 33).  C2's 25.6 MB matrix lives in L2: a step is two exchange latencies around a 3.6 us scan; "
                                  "the HBM-sized inputs run at 0.73 (C4) and 0.61 (C5) of the peak, profiles/r02_cli_stages.md")
                    out["phase_a_in_product"] = pa
            except Exception as e:
                out["seqs_clustered"] = {"error": str(e)[:200]}
        emit(out)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cli_leg(letters, offs, tmpl, cfg, gpus=1):
    import re
    import tempfile
    from meshclust_b200 import build, synth
    import _oracle
    cli = build.build_cli()
    res = {}
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "c2.fa")
        synth.write_fasta(fa, letters, offs, synth.headers_for(cfg.n, tmpl))
        t0 = time.perf_counter()
        env = dict(os.environ)
        if gpus > 1:
            env.pop("OMP_NUM_THREADS", None)   # torchrun exports OMP_NUM_THREADS=1 for its ranks; the CLI is its own program
        r = subprocess.run([cli, fa, "--id", str(cfg.identity), "--kmer", str(cfg.kmer), "--gpus", str(gpus), "--output", os.path.join(d, "o.clstr")],
                           capture_output=True, text=True, timeout=600, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-200:])
        # the CUDA context is created on a helper thread while the FASTA is read; only the part the
        # main thread had to wait for is start-up cost on the critical path
        m = re.search(r"gpu context ([0-9.]+)s on a helper thread, waited ([0-9.]+)s", r.stdout)
        ctx_s = float(m.group(2)) if m else 0.0
        ncl = open(os.path.join(d, "o.clstr")).read().count(">Cluster")
        pa = re.search(r"Accumulation: (\d+) clusters, (\d+) scans, (\d+) evals, on the device in ([0-9.]+)s \(([0-9.]+) us per step\)", r.stdout)
        near = re.search(r"Pairs within 1e-9 of the decision threshold: (\d+)", r.stdout)
        res = {"workload": f"c2 full (100k x 1.5 kb), bin/meshclust --id 0.97 --kmer 4 --gpus {gpus}", "wall_s": round(wall, 3),
               "cuda_context_wait_s": round(ctx_s, 3), "value": cfg.n / wall, "value_excluding_cuda_context_wait": cfg.n / max(wall - ctx_s, 1e-9),
               "stages": [ln.strip() for ln in r.stdout.splitlines() if "[" in ln and "s]" in ln][:16],
               "unit": "seqs/s", "clusters": ncl, "pairs_within_1e-9_of_threshold": int(near.group(1)) if near else None}
        if pa:
            res["phase_a"] = {"clusters": int(pa.group(1)), "scans": int(pa.group(2)), "evals": int(pa.group(3)), "device_s": float(pa.group(4)),
                              "us_per_step": float(pa.group(5)), "evals_per_s": int(pa.group(3)) / max(float(pa.group(4)), 1e-9)}
        if os.path.exists(_oracle.REF_BIN) and gpus == 1:
            ns = 4000
            fs = os.path.join(d, "c2_sample.fa")
            synth.write_fasta(fs, letters[: offs[ns]], offs[: ns + 1], synth.headers_for(ns, tmpl))
            t0 = time.perf_counter()
            rr = subprocess.run([_oracle.REF_BIN, fs, "--id", str(cfg.identity), "--kmer", str(cfg.kmer), "--output", os.path.join(d, "r.clstr")],
                                capture_output=True, text=True, timeout=900)
            rwall = time.perf_counter() - t0
            res["cpu_reference"] = {"sample": f"first {ns} sequences of c2, reference CLI, all host threads ({os.cpu_count()})",
                                    "wall_s": round(rwall, 2), "value": ns / rwall, "unit": "seqs/s", "rc": rr.returncode}
    return res


def c4_shape_probe(api, device, torch, dist, peak, rank, world, exchange):
    """BASELINE configs[3] shape: 1 M points x 1024 bins (1.06 GB per scan, > L2), the same rows at every N:
    every rank holds them all and scans its share of each scan (block-interleaved), summaries through the
    peer inboxes.  Both as a dependent chain (one launch per scan) and as independent scans."""
    n, k = 1_000_000, 5
    nb = 4 ** k
    if world > 1 and exchange != "peer_inbox":
        return {"skipped": "no peer memory"}
    rng = np.random.default_rng(1)
    base = rng.integers(1, 6, (1000, nb), dtype=np.uint8)
    hist = base[rng.integers(0, 1000, n)]
    lens = np.full(n, 1000, np.uint64)
    ctx = api.Context(device)
    ctx.load_histograms(hist, lens, k)
    del hist
    ctx.set_model(np.array([0, 0.5, 0, -1, 100.0]), np.array([100, 1, 4000, 1, 4000.0]), np.array([-1.0, 2, 1, 0.5, 0.5]), 4)
    if world > 1:
        handle = ctx.comm_init(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ctx.comm_connect(handles)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", device))
    S4 = 8
    cr = np.array([11, 500_000, 999_999, 123_456, 654_321, 42, 777_777, 31_337], np.int64)
    lo = np.zeros(S4, np.int64)
    hi = np.full(S4, n - 1, np.int64)
    inflight, last, cno = [], [None], [0]
    e = np.zeros(0, np.int64)

    def call(mode):
        c = cno[0]
        cno[0] += 1
        if world == 1:
            ctx.scan_enqueue_many(cr, lo, hi, mode, 0)
            return
        inflight.append(c)
        if len(inflight) > 2:
            old = inflight.pop(0)
            last[0] = ctx.scan_sharded_burst(cr, lo, hi, mode, (c % 3) * 16, (old % 3) * 16, S4)
        else:
            ctx.scan_sharded_burst(cr, lo, hi, mode, (c % 3) * 16, 0, 0)

    def drain():
        while inflight:
            old = inflight.pop(0)
            last[0] = ctx.scan_sharded_burst(e, e, e, 0, 0, (old % 3) * 16, S4)

    def run(mode, warm, calls):
        for _ in range(warm):
            call(mode)
        drain()
        ctx.sync()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        e0.record(stream)
        for _ in range(calls):
            call(mode)
        drain()
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) if world == 1 else (time.perf_counter() - tw) * 1e3
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            assert all(r[0] == n for r in last[0]), "c4 probe: sharded scan did not evaluate every point"
        us = ms * 1e3 / (calls * S4)
        by = n * (nb + 33)
        ach = by / (us * 1e-6) / 1e9
        return {"us_per_scan": round(us, 2), "evals_per_s": n / (us * 1e-6), "achieved_GBs": round(ach, 1),
                "frac_of_n_gpu_peak": round(ach / (peak * world), 4)}

    res = {"points": n, "bins": nb, "scaling": "strong: 1 M rows in total at every N",
           "dependent_chain": run(api.MC_SCAN_CHAIN, 2, 8), "independent_scans": run(api.MC_SCAN_KEEP, 2, 8),
           "note": "1.06 GB of rows per scan (> L2): every scan streams from HBM; timing = CUDA events (N = 1) / wall, max over ranks (N > 1)"}
    ctx.close()
    return res


def reference_arm(args):
    """The reference's own CPU implementation of the path (compiled, unmodified; OpenMP over points)
    on a bounded sample of the same workload."""
    import _oracle
    cfg, letters, offs, tmpl = make_workload()
    n = cfg.n
    ns = 20000
    sub_offs = offs[: ns + 1]
    sub_letters = letters[: sub_offs[-1]]
    o = _oracle.oracle()   # histograms + model for the sample come from the CPU oracle: no GPU on this arm
    rc, hist, mx = o.hist_batch(sub_letters, sub_offs, K, 1)
    lens = np.diff(sub_offs).astype(np.uint64)
    rng = np.random.default_rng(7)
    m = 3000
    a = rng.integers(0, ns, m)
    ntem = int(tmpl.max()) + 1
    b = np.where(rng.random(m) < 0.5, (a + ntem * rng.integers(1, 10, m)) % ns, rng.integers(0, ns, m))
    raw = np.array([o.features(hist[i], hist[j], int(lens[i]), int(lens[j]))[0] for i, j in zip(a, b)])
    mins, maxs = raw.min(0), np.maximum(raw.max(0), np.finfo(float).tiny)
    c = (raw - mins) / (maxs - mins)
    c[:, [0, 2, 3]] = 1 - c[:, [0, 2, 3]]
    feats = np.stack([c[:, 0] * c[:, 1], (c[:, 0] * c[:, 2]) ** 2, c[:, 3], (c[:, 0] * c[:, 4]) ** 2], 1)
    X = np.concatenate([np.ones((m, 1)), feats], 1)
    y = np.where(tmpl[a] == tmpl[b], 1.0, -1.0)
    w = np.linalg.lstsq(X, y, rcond=None)[0]
    centers = rng.integers(0, ns, 64)
    v, cores, kind, sample, ms_step = cpu_reference_rate(hist, lens, mins, maxs, w, centers, target_s=60.0,
                                                        steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic",
           "config": {"workload": f"{WORKLOAD}: 100k synthetic 1.5 kb 16S-like sequences, --id 0.97 --kmer 4 (bounded sample)",
                      "bins": 4 ** K, "parallelism": f"OpenMP x{cores} host threads"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
