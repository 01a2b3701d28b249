#!/usr/bin/env python
"""bench.py -- SR4000 frame-pairs/s (match + RANSAC) on synthetic 176x144-frame-shaped pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

Workload (BASELINE.json configs[2], the one the metric "frame-pairs/s (match+RANSAC)" is quoted
on): ONE synthetic sequence of 4096 frame pairs -- 512 SIFT descriptors (128-d, class double
holding float32 values) + 512 3-D points per frame, 300 re-observed features per step, 30 %
outliers, 5-point RANSAC with 2000 seeded sample sets per pair, the reference's adaptive stop --
"sharded by pair across 1/2/4/8 B200": rank r takes the contiguous block of pairs
dist.split_range(4096, r, N) and the 240-byte result records of all pairs are all-gathered to every
rank INSIDE the timed region ("scaling": "strong").  A step = one pass over the whole sequence.
config.weak_scaling carries the secondary figure with 4096 pairs PER GPU.

  value      device-resident inputs, pre3_pairs_dev, CUDA events on the launching stream
  e2e        pre3_pairs with pinned HOST buffers: H2D of descriptors / points and D2H of the
             results inside the timed region
  roofline   the dominant kernel of the step, timed live with CUDA events (pre3_timing_*)
  cpu_baseline  the reference CPU path (reference siftmatch.c from oracle/_ref when present +
             the C restatement of RANSAC_CALC_VER2), one thread, bounded sample
--impl reference times that CPU path with every host core (one process per core over pairs).
"""
from __future__ import annotations

import argparse
import importlib
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

K_FEAT, N_CORR, OUTLIER, H_HYP, K_MIN, MAX_IT = 512, 300, 0.30, 2000, 5, 2000
SEED = 3000  # 1000*cfg + index
FLOPS_PER_EVAL = 27.0        # R*y 15, +t 3, -x 3, squares 5, sqrt 1
WORKLOAD = "cfg3: ONE sequence of {P}+1 consecutive SR4000 frames = {P} frame pairs, sharded by pair over the GPUs, 512 " \
           "descriptors (double) per frame, 300 re-observed features per step, 30% outliers, k=5, 2000 sample sets, " \
           "adaptive stop"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------
def _cpu_pair_worker(args):
    """One pair through the CPU path: reference siftmatch.c (oracle/_ref) when built, else the C
    port; then gather + the C restatement of RANSAC_CALC_VER2 (oracle/pre3_oracle.c)."""
    seed, pair_id = args
    from oracle import oracle as orc
    from oracle import refmex
    fp = _cpu_pair_data(seed)
    t0 = time.perf_counter()
    if refmex.available():
        m = refmex.siftmatch(fp.desc1, fp.desc2, nout=1)[0]
        pairs = (m.T - 1).astype(np.int32)
    else:
        pairs, _ = orc.siftmatch(fp.desc1, fp.desc2, 1.5)
    Ya, Yb = fp.xyz1[pairs[:, 0]], fp.xyz2[pairs[:, 1]]
    samples = orc.sample_sets(7, pair_id, H_HYP, len(pairs), K_MIN)
    t1 = time.perf_counter()
    r = orc.ransac(Ya, Yb, samples, method=0, max_iteration=MAX_IT, adaptive=True)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, r.best_fit


_PAIR_CACHE = {}


def _cpu_pair_data(seed):
    """Synthetic pair `seed`, generated once per process (outside the timed region)."""
    if seed not in _PAIR_CACHE:
        synth = importlib.import_module("3pre_b200.synth")
        _PAIR_CACHE[seed] = synth.make_frame_pair(seed, K1=K_FEAT, K2=K_FEAT, n_corr=N_CORR, outlier_ratio=OUTLIER)
    return _PAIR_CACHE[seed]


def _cpu_pool_init(n):
    for i in range(n):
        _cpu_pair_data(SEED + i)
    from oracle import oracle as orc
    orc.lib()


def cpu_kind():
    from oracle import refmex
    return "reference" if refmex.available() else "port"


def cpu_baseline_single(n_pairs=24):
    ts = [_cpu_pair_worker((SEED + i, i)) for i in range(n_pairs)]
    total = sum(t[0] for t in ts)
    return {
        "value": n_pairs / total, "unit": "frame-pairs/s", "cores": 1, "kind": cpu_kind(),
        "sample": f"{n_pairs} pairs of the workload's shape (512x512 descriptors, ~300 matches), one thread: "
                  f"reference siftmatch.c (oracle/_ref) for matching "
                  f"[{sum(t[1] for t in ts) / total:.0%} of the time] + C restatement of RANSAC_CALC_VER2 "
                  "(MATLAB/Octave absent)" if cpu_kind() == "reference" else
                  f"{n_pairs} pairs of the workload, one thread, C restatement (oracle/pre3_oracle.c)",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = max(cores, 8) * 4
    # the reference's compiled siftmatch.c (oracle/_ref) and the C restatement are loaded HERE, in the parent, and
    # exercised once, so that the process the driver observes maps the libraries the forked workers run
    _cpu_pool_init(1)
    from oracle import refmex
    if refmex.available():
        refmex.lib()
    _cpu_pair_worker((SEED, 0))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_cpu_pool_init, initargs=(per_step,)) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_pair_worker, [(SEED + i, i) for i in range(per_step)], chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = per_step / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "SR4000 frame-pairs/s (match+RANSAC)", "value": value,
        "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD.format(P=4096) + f" [bounded sample: {per_step} pairs per step, all host cores]"},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": cpu_kind(),
                         "sample": f"{per_step} pairs per step over {cores} processes"},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML through pynvml in a thread (a sample every ~2 ms:
    the timed region of the default run is tens of milliseconds), nvidia-smi -lms 100 as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(reasons(self.handle))
                self.rows.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in ln.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (the timed region); when the region is too short for
        three samples, every sample since start() (warm-up + timed region, same load)."""
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            rows = [r for r in self.rows if t0 is None or t0 <= r[0] <= t1]
            window = "timed region"
            if len(rows) < 3:
                rows, window = list(self.rows), "warm-up + timed region"
            sm = [r[1] for r in rows]
            bits = 0
            for r in rows:
                bits |= r[2]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(v for k, v in self.BITS.items() if bits & k), "samples": len(sm), "window": window,
                    "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 6 and (t0 is None or t0 - 0.02 <= t <= t1 + 0.12)]
        window = "timed region"
        if len(rows) < 3:
            rows, window = [r for _, r in self.rows if len(r) >= 6], "warm-up + timed region"
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window, "source": "nvidia-smi"}



# ------------------------------------------------------------------------------------------
# the other BASELINE.json configs, measured briefly next to the headline workload
# ------------------------------------------------------------------------------------------
def _time_steps(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _kernel_times(ctx, fn, reps=2):
    ctx.timing_enable(True)
    for _ in range(reps):
        fn()
    kt = ctx.timing_read()
    ctx.timing_enable(False)
    return {k: {"ms_per_step": v[0] / reps, "launches_per_step": v[1] // reps} for k, v in kt.items()}


def bench_cfg2(ctx, pre3, synth, dev, rank, P=256, K=2048, steps=5, warmup=2):
    """configs[1]: SIFT descriptor matching 2k x 2k 128-d, batch of 256 pairs, ratio test, one GPU."""
    import torch
    slabs = [synth.make_batch_torch(32, 2000 + 100000 * rank + s0, dev, K1=K, K2=K, n_corr=K // 2)
             for s0 in range(0, P, 32)]
    d1 = torch.cat([s["desc1"] for s in slabs]).contiguous()
    d2 = torch.cat([s["desc2"] for s in slabs]).contiguous()
    del slabs
    pairs = torch.zeros(P, K, 2, dtype=torch.int32, device=dev)
    n_out = torch.zeros(P, dtype=torch.int32, device=dev)

    def step():
        ctx.siftmatch_batch_dev(d1, d2, pairs, None, n_out, 1.5)

    ms = _time_steps(step, steps, warmup)
    kt = _kernel_times(ctx, step)
    flops = 2.0 * K * K * 128 * P
    pk = peaks()
    gemm = kt.get("match_tc", {}).get("ms_per_step", 0.0)
    out = {"workload": f"cfg2: siftmatch {K}x{K}x128 (double), batch of {P} pairs, ratio 1.5",
           "pairs_per_s": P / (ms * 1e-3), "ms_per_step": ms,
           "tflops_whole_step": flops / (ms * 1e-3) / 1e12,
           "accepted_matches_per_pair": float(n_out.float().mean().item()), "kernels": kt}
    if gemm > 0:
        a = flops / (gemm * 1e-3) / 1e12
        out["roofline"] = {"kernel": "match_tc", "bound": "tensor", "achieved": a, "peak": pk["bf16_tflops"],
                           "unit": "TFLOP/s", "frac": a / pk["bf16_tflops"], "traffic": None,
                           "note": f"algorithmic 2*K1*K2*128 per pair; peak {pk['source']}"}
    # the uint8 variant of the same config (sift_demo2.m:93-94: uint8(512 * descr)): tcgen05 kind::i8, exact, no rescore
    u1 = torch.clamp(torch.floor(512.0 * d1 + 0.5), 0, 255).to(torch.uint8)
    u2 = torch.clamp(torch.floor(512.0 * d2 + 0.5), 0, 255).to(torch.uint8)
    del d1, d2

    def step_u8():
        ctx.siftmatch_batch_dev(u1, u2, pairs, None, n_out, 1.5)

    ms8 = _time_steps(step_u8, steps, warmup)
    kt8 = _kernel_times(ctx, step_u8)
    g8 = kt8.get("match_tc", {}).get("ms_per_step", 0.0)
    out["uint8"] = {"pairs_per_s": P / (ms8 * 1e-3), "ms_per_step": ms8,
                    "accepted_matches_per_pair": float(n_out.float().mean().item()), "kernels": kt8}
    if g8 > 0:
        a8 = flops / (g8 * 1e-3) / 1e12
        out["uint8"]["roofline"] = {"kernel": "k_i8_gemm_pair", "bound": "tensor", "achieved": a8,
                                    "peak": 2.0 * pk["bf16_tflops"], "unit": "TOP/s", "frac": a8 / (2.0 * pk["bf16_tflops"]),
                                    "traffic": None,
                                    "note": "algorithmic 2*K1*K2*128 integer ops per pair; peak = 2 x the measured bf16 rate "
                                            "(the int8 tensor rate is not in MEASURED_PEAKS.json; nominal int8 = 2 x bf16)"}
    del u1, u2
    torch.cuda.empty_cache()
    return out


EKF_FLOPS_PER_EVAL = 600.0   # K-form state update 468 + sin/cos 72 + projection/distortion 60 (mul and add counted apart)


def bench_cfg4(ctx, pre3, dev, rank, Fr=256, n_id=200, H=1000, steps=3, warmup=1):
    """configs[3]: 1-point-RANSAC EKF hypothesis support, 200 inverse-depth features (n = 1213)."""
    import torch
    se = importlib.import_module("3pre_b200.synth_ekf")
    parts = [se.make_ekf_frames(64, 4000 + 100000 * rank + f0, device=dev, n_id=n_id, outlier_ratio=0.2)
             for f0 in range(0, Fr, 64)]
    b = {k: (torch.cat([p[k] for p in parts]).contiguous() if hasattr(parts[0][k], "shape") else parts[0][k])
         for k in parts[0]}
    del parts
    b["cam"] = dict(se.CAM)
    li = torch.zeros(Fr, b["F"], dtype=torch.uint8, device=dev)
    res = torch.zeros(Fr, 32, dtype=torch.uint8, device=dev)
    out = {"workload": f"cfg4: ransac_hypotheses on {Fr} resident EKF frames (of the 10k-frame config), {n_id} "
                       f"inverse-depth features, n = {b['n']}, {H} seeded match triples per frame, thr = 1 px, 20% outliers"}
    fp64_peak = ctx.measure_fp64_peak()
    for name, adaptive in (("adaptive", True), ("fixed_H", False)):
        o = pre3.make_ekf_opts(H=H, adaptive=adaptive, seed=11)

        def step():
            ctx.ransac_hypotheses_batch_dev(b, o, li, res)

        ms = _time_steps(step, steps, warmup)
        kt = _kernel_times(ctx, step, reps=1)
        r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.EKF_RESULT_DTYPE)
        ends = ctx.ekf_eval_schedule(o)
        done = ends[np.minimum(np.searchsorted(ends, np.maximum(r["n_evaluated"], 1) - 1, side="right"), len(ends) - 1)]
        evals = float((done.astype(np.float64) * n_id).sum())
        sc = kt.get("ekf_score", {}).get("ms_per_step", 0.0)
        gn = kt.get("ekf_gain", {}).get("ms_per_step", 0.0)
        o_ = {"frames_per_s": Fr / (ms * 1e-3), "ms_per_step": ms, "hyp_x_feature_evals_per_s": evals / (ms * 1e-3),
              "hypotheses_evaluated_per_frame": float(done.mean()),
              "hypotheses_needed_by_reference_loop_per_frame": float(r["n_evaluated"].mean()),
              "mean_support": float(r["max_support"].mean()), "kernels": kt}
        if sc > 0:
            a = EKF_FLOPS_PER_EVAL * evals / (sc * 1e-3) / 1e12
            o_["roofline_score"] = {"kernel": "ekf_score", "bound": "fp64", "achieved": a, "peak": fp64_peak,
                                    "unit": "TFLOP/s", "frac": a / fp64_peak if fp64_peak else None, "traffic": None,
                                    "note": "600 fp64 ops per hypothesis x feature eval; peak = DFMA-chain microbenchmark "
                                            "of this run (counts 2 per FMA; the path may not fuse, so 0.5 is its ceiling)"}
        if gn > 0:
            byt = Fr * (b["n"] * b["n"] * 8.0 + b["F"] * 2 * b["n"] * 8.0)
            a = byt / (gn * 1e-3) / 1e9
            pk = peaks()
            o_["roofline_gain"] = {"kernel": "ekf_gain", "bound": "hbm", "achieved": a, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": a / pk["hbm_gbs"], "traffic": None,
                                   "note": "reads P (n*n*8 B) once, writes G (F*2*n*8 B) per frame"}
        out[name] = o_
    del b
    torch.cuda.empty_cache()
    return out


def bench_cfg5(ctx, pre3, synth, dev, rank, world, N=20000, H=1000000, steps=3, warmup=1):
    """configs[4]: one pair, 20k correspondences x 1M hypotheses, 60% outliers, hypotheses split over
    the ranks, ONE NCCL max-reduce of (inlier count, hypothesis id)."""
    import torch
    import torch.distributed as dist
    pd = importlib.import_module("3pre_b200.dist")
    c = synth.make_correspondences(5000, N=N, outlier_ratio=0.6)
    Ya, Yb = torch.from_numpy(c.Ya).to(dev), torch.from_numpy(c.Yb).to(dev)
    if world > 1:  # replicate the correspondences (480 KB) from rank 0
        dist.broadcast(Ya, 0)
        dist.broadcast(Yb, 0)
    opts = pre3.make_opts(method=0, k=5, max_iteration=H + 1, adaptive=False, H=H, seed=5)
    out = {"workload": f"cfg5: one pair, {N} correspondences x {H} hypotheses (k=5, fixed H, 60% outliers), "
                       f"hypothesis blocks over {world} GPU(s)", "scaling": "strong"}
    for mode in ("first", "reference"):
        def step():
            return pd.ransac_hypothesis_split(ctx, Ya, Yb, opts, mode=mode, want_mask=False)

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            rec, _ = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[mode] = {"ms_per_solve": ms, "hyp_x_match_evals_per_s": float(N) * H / (ms * 1e-3),
                     "best_fit": int(rec["best_fit"]), "best_sample": int(rec["best_sample"]),
                     "collective": "all_reduce(MAX) of one u64 key" if mode == "first"
                     else "all_gather of 16 B per rank (count, id, ErrorSum)"}
    # the scoring kernel alone, on this rank's block (fp32 CUDA-core roofline at large N)
    h0, h1 = pd.split_range(H, rank, world)
    key = torch.zeros(1, dtype=torch.int64, device=dev)
    thr = torch.zeros(1, dtype=torch.float64, device=dev)
    ctx.distance_threshold_dev(Yb, thr)
    ctx.sync()
    kt = _kernel_times(ctx, lambda: ctx.ransac_block_dev(Ya, Yb, opts, h0, h1 - h0, float(thr.item()), key), reps=2)
    ev = kt.get("eval", {}).get("ms_per_step", 0.0)
    if ev > 0:
        fp32_peak = ctx.measure_fp32_peak()
        fp32_peak_3reg = ctx.measure_fp32_peak_3reg()
        a = FLOPS_PER_EVAL * float(N) * (h1 - h0) / (ev * 1e-3) / 1e12
        out["roofline"] = {"kernel": "eval (k_eval<5,0>)", "bound": "fp32", "achieved": a, "peak": fp32_peak,
                           "unit": "TFLOP/s", "frac": a / fp32_peak, "traffic": None,
                           "peak_register_operands": fp32_peak_3reg, "frac_of_register_operand_peak": a / fp32_peak_3reg,
                           "evals_per_s": float(N) * (h1 - h0) / (ev * 1e-3),
                           "note": "27 FLOP per hypothesis x match eval; the fp64 5-point fits run in the same kernel; "
                                   "peak = FFMA-chain microbenchmark of this run"}
    out["kernels"] = kt
    return out


def bench_dr_ye(ctx, pre3, synth, dev, rank, P=1024, steps=5, warmup=2):
    """the code_from_dr_ye variant (vodometry_dr_ye.m) on a cfg3-shaped sequence: match +
    700 four-match hypotheses per pair, every one scored (no early stop: the reference's for-range is fixed),
    first-max selection, refit, residual statistics."""
    import torch
    L = importlib.import_module("3pre_b200._lib")
    sq = synth.make_sequence_torch(P + 1, 8100 + 100000 * rank, dev, K=K_FEAT, n_corr=N_CORR, outlier_ratio=OUTLIER)
    opts = pre3.make_opts(method=L.METHOD_DR_YE, k=4, max_iteration=700, adaptive=False, H=700, seed=7)
    res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
    masks = torch.zeros(P, K_FEAT, dtype=torch.uint8, device=dev)

    def step():
        ctx.sequence_dev(sq["desc"], sq["xyz"], opts, res, None, masks, pair_id0=0)

    ms = _time_steps(step, steps, warmup)
    kt = _kernel_times(ctx, step)
    rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
    evals = float((rec["n_consumed"].astype(np.int64) * rec["n_matches"]).sum())
    out = {"workload": f"code_from_dr_ye variant: sequence of {P}+1 frames, {K_FEAT} descriptors per frame, 700 "
                       "four-match hypotheses per pair, all scored (vodometry_dr_ye.m:162-183)",
           "pairs_per_s": P / (ms * 1e-3), "ms_per_step": ms, "kernels": kt,
           "pairs_solved": int((rec["status"] == 0).sum()), "mean_support": float(rec["best_fit"].mean()),
           "hyp_x_match_evals_per_s": evals / (ms * 1e-3)}
    ev = kt.get("eval", {}).get("ms_per_step", 0.0)
    if ev > 0:
        a = 27.0 * evals / (ev * 1e-3) / 1e12
        out["roofline"] = {"kernel": "eval", "bound": "fp32", "achieved": a, "unit": "TFLOP/s",
                           "note": "27 FLOP per hypothesis x match eval (the fp64 4-point fits run in the same kernel)"}
    if rank == 0:
        from oracle import oracle as orc
        fps = [synth.make_frame_pair(8200 + i, K1=K_FEAT, K2=K_FEAT, n_corr=N_CORR, outlier_ratio=OUTLIER) for i in range(3)]
        t0 = time.perf_counter()
        for i, f in enumerate(fps):
            pairs, _ = orc.siftmatch(f.desc1, f.desc2, 1.5)
            orc.vodometry_dr_ye(f.xyz1[pairs[:, 0]], f.xyz2[pairs[:, 1]], match=pairs, seed=7, pair=i)
        dt = time.perf_counter() - t0
        out["cpu"] = {"pairs_per_s": 3 / dt, "cores": 1, "kind": "port",
                      "sample": "3 pairs: C restatement of siftmatch + vodometry_dr_ye (oracle/), one thread"}
    del sq
    torch.cuda.empty_cache()
    return out


def bench_frames(ctx, pre3, dev, F=2048, K=512, steps=5, warmup=2):
    """SR4000 frame batches -> per-feature 3-D points (read_xyz_sr4000.m + the loop of
    SIFT_extract_save.m:75-88).  Two forms: the full filtered maps (what read_xyz_sr4000 returns) and the fused
    per-feature lookup + compaction the matching path actually needs."""
    import torch
    g = torch.Generator(device=dev).manual_seed(77)
    sr = torch.empty(F, 176, 720, dtype=torch.float64, device=dev)
    sr[:, :, 0:144] = 2.5 + torch.rand(F, 176, 144, generator=g, device=dev, dtype=torch.float64)
    sr[:, :, 144:432] = torch.rand(F, 176, 288, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    sr[:, :, 432:720] = torch.rand(F, 176, 288, generator=g, device=dev, dtype=torch.float64) * 60000
    fr = torch.zeros(F, K, 4, dtype=torch.float64, device=dev)
    fr[:, :, 0] = torch.rand(F, K, generator=g, device=dev, dtype=torch.float64) * 175
    fr[:, :, 1] = torch.rand(F, K, generator=g, device=dev, dtype=torch.float64) * 143
    desc = torch.rand(F, K, 128, generator=g, device=dev, dtype=torch.float64)
    o = pre3.make_frame_opts()
    x, y, z = (torch.empty(F, 176, 144, dtype=torch.float64, device=dev) for _ in range(3))
    mc = torch.empty(F, dtype=torch.float64, device=dev)
    xyz = torch.empty(F, K, 3, dtype=torch.float64, device=dev)
    nk = torch.empty(F, dtype=torch.int32, device=dev)
    dout = torch.empty_like(desc)
    pk = peaks()
    out = {"workload": f"{F} SR4000 frames (720 x 176 doubles each), {K} features per frame"}
    ms = _time_steps(lambda: ctx.read_xyz_sr4000_batch_dev(sr, o, x, y, z, mc), steps, warmup)
    b = F * (7 * 144 * 176 * 8.0)   # read z, x, y + confidence, write z, x, y
    out["maps"] = {"ms_per_step": ms, "frames_per_s": F / (ms * 1e-3),
                   "roofline": {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                "frac": b / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                                "note": "algorithmic bytes: 3 maps in + 3 maps out + confidence map = 1.42 MB per frame"}}
    ms = _time_steps(lambda: ctx.features_xyz_batch_dev(sr, o, fr, xyz=xyz, n_keep=nk, desc_in=desc, desc_out=dout),
                     steps, warmup)
    kept = float(nk.float().sum().item())
    b = F * (144 * 176 * 8.0 + K * (27 * 8 + 8 + 32 + 24 + 1024)) + kept * 1024
    out["fused_features"] = {"ms_per_step": ms, "frames_per_s": F / (ms * 1e-3), "mean_kept": float(nk.float().mean().item()),
                             "roofline": {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                          "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                                          "note": "algorithmic bytes per frame: confidence map 203 KB + per feature 27 "
                                                  "stencil taps, frame entry, point, 1 KB of compacted descriptor "
                                                  "written, and 1 KB read per surviving feature"}}
    del sr, fr, desc, x, y, z, dout
    torch.cuda.empty_cache()
    return out


def bench_ekf_update(ctx, pre3, dev, Fr=64, n_id=200, steps=3, warmup=1):
    """SURVEY.md 8f rank 3 (first part): update.m on the low-innovation inliers of cfg4-shaped frames (n = 1213,
    ~160 of 200 features flagged -> m ~ 320 stacked rows): dense fp64 linear algebra, 2 n^2 m FLOP for the
    covariance update alone."""
    import torch
    se = importlib.import_module("3pre_b200.synth_ekf")
    b = se.make_ekf_frames(Fr, 4000, device=dev, n_id=n_id, outlier_ratio=0.2)
    sel = (~b["outlier"]).to(torch.uint8)
    xo, Po = torch.empty_like(b["x"]), torch.empty_like(b["P"])
    mo = torch.zeros(Fr, dtype=torch.int32, device=dev)

    def step():
        ctx.ekf_update_batch_dev(b, sel, xo, Po, m_out=mo)

    ms = _time_steps(step, steps, warmup)
    n = b["n"]
    m = mo.double()
    # update.m as written: K = (P H') inv(S): 2 n m^2; K S: 2 n m^2; (K S) K': 2 n^2 m; inv 2 m^3; P H', H (P H'): 38 (n + m) m
    flops_ref = float((4 * n * m * m + 2 * n * n * m + 2 * m ** 3 + 38 * (n + m) * m).sum().item())
    # executed here: K S = P H' is not recomputed, and only the lower triangle of the symmetric K S K' is formed
    flops_exec = float((2 * n * m * m + n * n * m + 2 * m ** 3 + 38 * (n + m) * m).sum().item())
    out = {"workload": f"update.m on {Fr} frames, n = {n}, mean stacked rows m = {float(m.mean().item()):.0f}",
           "frames_per_s": Fr / (ms * 1e-3), "ms_per_step": ms,
           "fp64_tflops_executed": flops_exec / (ms * 1e-3) / 1e12,
           "fp64_tflops_of_reference_formula": flops_ref / (ms * 1e-3) / 1e12,
           "note": "fp64 DFMA peak measured by csrc/peak.cu: ~34 TFLOP/s",
           "kernels": _kernel_times(ctx, step, reps=1)}
    try:  # CPU: the dense numpy / LAPACK restatement of update.m (multi-threaded BLAS), one frame
        from oracle import ref_numpy_ekf as rne
        fr = se.frame(b, 0)
        flags = sel[0].cpu().numpy()
        rne.ekf_update_inliers(fr, flags)
        t0 = time.perf_counter()
        rne.ekf_update_inliers(fr, flags)
        dt = time.perf_counter() - t0
        out["cpu"] = {"frames_per_s": 1.0 / dt, "cores": os.cpu_count(), "kind": "port",
                      "sample": "1 frame, numpy / LAPACK restatement (oracle/ref_numpy_ekf.py: update), BLAS threads"}
    except Exception as e:
        out["cpu"] = {"error": f"{type(e).__name__}: {e}"}
    del b, xo, Po
    torch.cuda.empty_cache()
    return out


def bench_cfg1(ctx, pre3, synth, dev):
    """configs[0]: one synthetic SR4000 frame pair (~300 matches, 30% outliers), 2000 hypotheses: a LATENCY
    number on the GPU (one pair cannot fill the machine); k=5 (RANSAC_CALC_VER2.m:85) and k=3 (BASELINE wording)."""
    import torch
    fp = synth.make_frame_pair(1000, K1=512, K2=512, n_corr=300, outlier_ratio=0.30)
    h = {k: np.ascontiguousarray(getattr(fp, k))[None] for k in ("desc1", "desc2", "xyz1", "xyz2")}
    d = {k: torch.from_numpy(v).to(dev) for k, v in h.items()}
    res = torch.zeros(1, 240, dtype=torch.uint8, device=dev)
    out = {"workload": "cfg1: one SR4000 frame pair, 512x512 descriptors -> ~300 matches, 30% outliers, 2000 sample "
                       "sets, adaptive stop (latency; match + RANSAC + refit)"}
    for k in (5, 3):
        opts = pre3.make_opts(method=0, k=k, max_iteration=2000, adaptive=True, H=2000, seed=7)
        ms_dev = _time_steps(lambda: ctx.pairs_dev(d["desc1"], d["desc2"], d["xyz1"], d["xyz2"], opts, res), 20, 3)
        t0 = time.perf_counter()
        for _ in range(20):
            r, _, _ = ctx.pairs(h["desc1"], h["desc2"], h["xyz1"], h["xyz2"], opts)
        ms_host = (time.perf_counter() - t0) / 20 * 1e3
        # fixed H (adaptive stop off): all 2000 sample sets are evaluated -> hypothesis x match evals/s of ONE pair
        fopts = pre3.make_opts(method=0, k=k, max_iteration=2001, adaptive=False, H=2000, seed=7)
        ms_fix = _time_steps(lambda: ctx.pairs_dev(d["desc1"], d["desc2"], d["xyz1"], d["xyz2"], fopts, res), 20, 3)
        out[f"k{k}"] = {"ms_per_pair_device_resident": ms_dev, "ms_per_pair_host_buffers": ms_host,
                        "best_fit": int(r["best_fit"][0]), "n_matches": int(r["n_matches"][0]),
                        "sample_sets_consumed": int(r["n_consumed"][0]),
                        "fixed_H": {"ms_per_pair_device_resident": ms_fix,
                                    "hyp_x_match_evals_per_s": 2000.0 * int(r["n_matches"][0]) / (ms_fix * 1e-3)}}
    return out


def cpu_other_baselines(synth):
    """The CPU path (one thread) on bounded samples of cfg2 / cfg4 / cfg5, for the report."""
    from oracle import oracle as orc
    from oracle import refmex
    out = {}
    fp = synth.make_frame_pair(2000, K1=2048, K2=2048, n_corr=1024)
    t0 = time.perf_counter()
    if refmex.available():
        refmex.siftmatch(fp.desc1, fp.desc2, nout=1)
    else:
        orc.siftmatch(fp.desc1, fp.desc2, 1.5)
    dt = time.perf_counter() - t0
    out["cfg2"] = {"pairs_per_s": 1.0 / dt, "cores": 1, "kind": cpu_kind(), "sample": "1 pair 2048x2048x128 (double)"}
    se = importlib.import_module("3pre_b200.synth_ekf")
    b = se.make_ekf_frames(2, 4000, n_id=200, outlier_ratio=0.2)
    t0 = time.perf_counter()
    ne = 0
    for f in range(2):
        ne += orc.ransac_hypotheses(se.frame(b, f), None, H=1000, seed=11, frame_id=f)["n_evaluated"]
    dt = time.perf_counter() - t0
    out["cfg4"] = {"frames_per_s": 2.0 / dt, "hyp_x_feature_evals_per_s": ne * 200 / dt, "cores": 1, "kind": "port",
                   "sample": f"2 frames, adaptive stop ({ne} hypotheses evaluated), C restatement (oracle/pre3_oracle_ekf.c)"}
    c = synth.make_correspondences(5000, N=20000, outlier_ratio=0.6)
    smp = orc.sample_sets(5, 0, 200, 20000, 5)
    t0 = time.perf_counter()
    orc.ransac(c.Ya, c.Yb, smp, method=0, max_iteration=201, adaptive=False)
    dt = time.perf_counter() - t0
    out["cfg5"] = {"hyp_x_match_evals_per_s": 200 * 20000 / dt, "cores": 1, "kind": "port",
                   "ms_per_solve_extrapolated": dt / 200 * 1e6 * 1e3,
                   "sample": "200 of the 1M hypotheses, extrapolated linearly (the literal reference stores a support "
                             "set per hypothesis, infeasible at 1M)"}
    return out


def other_workloads(ctx, pre3, synth, dev, rank, world):
    out = {}
    jobs = [("cfg5", lambda: bench_cfg5(ctx, pre3, synth, dev, rank, world))]
    if rank == 0:  # single-GPU configs: measured on rank 0 only (the other ranks wait at the next barrier)
        jobs = [("cfg1", lambda: bench_cfg1(ctx, pre3, synth, dev)),
                ("cfg2", lambda: bench_cfg2(ctx, pre3, synth, dev, rank)),
                ("cfg4", lambda: bench_cfg4(ctx, pre3, dev, rank)),
                ("dr_ye", lambda: bench_dr_ye(ctx, pre3, synth, dev, rank)),
                ("frames", lambda: bench_frames(ctx, pre3, dev)),
                ("ekf_update", lambda: bench_ekf_update(ctx, pre3, dev)),
                ("cpu", lambda: cpu_other_baselines(synth))] + jobs
    # cfg5 involves every rank: run it first everywhere so that no rank waits inside a collective
    jobs.sort(key=lambda j: j[0] != "cfg5")
    for name, fn in jobs:
        try:
            out[name] = fn()
        except Exception as e:  # a secondary workload must not take the headline line down
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def _summ(o, keys):
    """Flat numeric summary of a secondary workload (what the driver's record keeps inside `config`)."""
    out = {}
    for k in keys:
        v = o
        for part in k.split("."):
            v = v.get(part) if isinstance(v, dict) else None
        if v is not None:
            out[k] = v
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pre3 = importlib.import_module("3pre_b200")
    synth = importlib.import_module("3pre_b200.synth")
    pd = importlib.import_module("3pre_b200.dist")

    P = args.pairs                       # pairs of the WHOLE sequence (strong scaling: split over the ranks)
    p0, p1 = pd.split_range(P, rank, world)
    Pl = p1 - p0                         # this rank's block of pairs = frames p0 .. p1
    # one side stream for the library AND torch / NCCL (the legacy default stream cannot be captured into a graph)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    ctx = pre3.Context(local)
    ctx.use_torch_stream()
    ctx.set_graphs(True)   # the step repeats one call signature: its launch sequence is captured and replayed
    # the same synthetic sequence on every rank (same seed, generated on the device); a rank keeps its block
    sq = synth.make_sequence_torch(P + 1, SEED, dev, K=K_FEAT, n_corr=N_CORR, outlier_ratio=OUTLIER)
    full = {"desc": sq["desc"], "xyz": sq["xyz"]} if world > 1 else None
    data = {"desc": sq["desc"][p0:p1 + 1].contiguous(), "xyz": sq["xyz"][p0:p1 + 1].contiguous()}
    del sq
    torch.cuda.empty_cache()
    opts = pre3.make_opts(method=0, k=K_MIN, max_iteration=MAX_IT, adaptive=True, H=H_HYP, seed=7)
    res = torch.zeros(Pl, 240, dtype=torch.uint8, device=dev)
    matches = torch.zeros(Pl, K_FEAT, 2, dtype=torch.int32, device=dev)
    masks = torch.zeros(Pl, K_FEAT, dtype=torch.uint8, device=dev)
    even = Pl * world == P
    res_all = torch.zeros(P, 240, dtype=torch.uint8, device=dev) if world > 1 else res
    gathered = [res_all]

    def step():
        ctx.sequence_dev(data["desc"], data["xyz"], opts, res, matches, masks, pair_id0=p0)
        if world > 1:  # the records of every pair on every rank: 240 B per pair, one all-gather
            if even:
                dist.all_gather_into_tensor(res_all, res)
            else:
                gathered[0] = pd.gather_records(res, P)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    # no cyclic-GC pause inside the timed region: at 8 GPUs it is K x 0.26 ms long and is the MAX over 8 host processes
    import gc
    gc.collect()
    gc.disable()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    tw0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    gc.enable()
    clocks = sampler.stop(tw0, time.perf_counter()) if rank == 0 else None
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    ms_step = float(t.item()) / args.steps
    value = P / (ms_step * 1e-3)

    # sanity of the timed work: every pair of the WHOLE sequence solved (records gathered in the timed region)
    r = np.frombuffer(gathered[0].cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
    ok_pairs = int(((r["status"] == 0) & (r["best_fit"] > 150)).sum())
    rl = r[p0:p1]
    ends = ctx.eval_schedule(opts, P=p1 - p0)  # schedule of the batched path: a pair that consumed n sets had ends[first i: n < ends[i]] evaluated
    wave = np.minimum(np.searchsorted(ends, rl["n_consumed"], side="right"), len(ends) - 1)
    hyps_done = ends[wave].astype(np.float64)
    evals_done = float((rl["n_matches"].astype(np.float64) * hyps_done).sum())          # evaluated on this GPU
    evals_needed = float((rl["n_matches"].astype(np.float64) * rl["n_consumed"]).sum())  # the reference's loop

    # ---- per-kernel timing (separate steps, events around every launch; this rank's block) -----------
    ctx.timing_enable(True)
    for _ in range(2):
        ctx.sequence_dev(data["desc"], data["xyz"], opts, res, matches, masks, pair_id0=p0)
    kt = ctx.timing_read()
    ctx.timing_enable(False)
    per_kernel = {k: {"ms_per_step": v[0] / 2, "launches_per_step": v[1] // 2} for k, v in kt.items()}
    tot = sum(v["ms_per_step"] for v in per_kernel.values()) or 1.0
    for v in per_kernel.values():
        v["share"] = v["ms_per_step"] / tot
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    pk = peaks()
    fp32_peak = ctx.measure_fp32_peak()
    fp32_peak_3reg = ctx.measure_fp32_peak_3reg()
    desc_bytes = float((Pl + 1) * K_FEAT * 128 * data["desc"].element_size())  # every frame is converted once
    match_flops = 2.0 * K_FEAT * K_FEAT * 128 * Pl
    rooflines = {}
    for name, v in per_kernel.items():
        sec = v["ms_per_step"] * 1e-3 / max(v["launches_per_step"], 1)
        L = max(v["launches_per_step"], 1)
        if name == "eval":
            a = FLOPS_PER_EVAL * evals_done / L / sec / 1e12
            rooflines[name] = {"bound": "fp32", "achieved": a, "peak": fp32_peak, "unit": "TFLOP/s",
                               "frac": a / fp32_peak, "traffic": None,
                               "frac_on_required_evals": FLOPS_PER_EVAL * evals_needed / L / sec / 1e12 / fp32_peak,
                               "note": "27 FLOP per hypothesis x match eval over the evaluations executed (the fp64 "
                                       "minimal fits run in the same kernel and are not counted); peak = FFMA-chain "
                                       "microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure); "
                                       "frac_on_required_evals counts only the evals the reference loop needs",
                               "evals_per_s": evals_done / L / sec,
                               "peak_register_operands": fp32_peak_3reg, "frac_of_register_operand_peak": a / fp32_peak_3reg}
        elif name in ("match_tc",):
            a = match_flops / L / sec / 1e12
            rooflines[name] = {"bound": "tensor", "achieved": a, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": a / pk["bf16_tflops"], "traffic": None, "note": f"of {pk['source']}"}
        elif name in ("match_exact",):
            a = match_flops / L / sec / 1e12
            rooflines[name] = {"bound": "fp64", "achieved": a, "peak": None, "unit": "TFLOP/s", "frac": None,
                               "traffic": None, "note": "exact fp64 brute force (3 flop per 2 algorithmic)"}
        elif name == "match_fused":
            # conversion + proposal GEMM in one kernel (k_tc_seq_fused): the roofline that bounds it at this shape is HBM
            # (reading the fp64 descriptors: 0.33 ms at the measured copy rate) -- the tensor floor of its
            # 2*K1*K2*128 flop per pair is 0.16 ms at the measured bf16 rate
            a = desc_bytes / L / sec / 1e9  # every descriptor set read once (class double: 8 B per value); nothing written back
            tf = match_flops / L / sec / 1e12
            rooflines[name] = {"bound": "hbm", "achieved": a, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": a / pk["hbm_gbs"], "traffic": None,
                               "tensor_tflops": tf, "tensor_frac": tf / pk["bf16_tflops"],
                               "note": f"algorithmic bytes = the descriptor sets of the sequence read once; of {pk['source']}; "
                                       "the same launch also does the 2*K1*K2*128 flop per pair (tensor_tflops)"}
        elif name == "convert":
            a = desc_bytes * 1.25 / L / sec / 1e9  # read f64 descriptors + write the f16 operand image
            rooflines[name] = {"bound": "hbm", "achieved": a, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": a / pk["hbm_gbs"], "traffic": None, "note": f"of {pk['source']}"}
    # traffic: DRAM bytes per launch measured by ncu (profiles/r01_traffic.json: bytes per pair at the same
    # per-pair shape), scaled to the pairs one launch processes here
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tr = json.load(open(tpath))["per_pair_bytes"]
        for name in rooflines:
            if name in tr:
                units = (Pl + 1) if name in ("convert", "match_fused") else Pl  # bytes per descriptor set (frame)
                rooflines[name]["traffic"] = tr[name]["bytes"] * units
                rooflines[name]["traffic_note"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch " \
                                                  f"({os.path.basename(tpath)}, per pair) x pairs per launch"
    roofline = dict(rooflines.get(dom, {"bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                        "frac": None, "traffic": None}))
    roofline["kernel"] = dom
    roofline["per_kernel"] = {k: {kk: v.get(kk) for kk in ("bound", "achieved", "peak", "unit", "frac")}
                              for k, v in rooflines.items()}

    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "value": value, "ms_per_step": ms_step, "kernels": per_kernel,
                  "roofline": roofline, "gpu_launches": int(launches)})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- secondary: weak scaling (the whole 4096-pair sequence on EVERY GPU, no gather) ------------------
    weak = None
    if world > 1:
        wres = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        wm = torch.zeros(P, K_FEAT, 2, dtype=torch.int32, device=dev)
        wk = torch.zeros(P, K_FEAT, dtype=torch.uint8, device=dev)

        def wstep():
            ctx.sequence_dev(full["desc"], full["xyz"], opts, wres, wm, wk, pair_id0=0)

        for _ in range(2):
            wstep()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(5):
            wstep()
        w1.record()
        torch.cuda.synchronize()
        tw = torch.tensor([w0.elapsed_time(w1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"pairs_per_gpu": P, "ms_per_step": float(tw.item()), "value": world * P / (float(tw.item()) * 1e-3),
                "unit": "frame-pairs/s", "note": "every GPU runs the whole sequence (independent replicas, no collective)"}
        del wres, wm, wk, full
        torch.cuda.empty_cache()

    # ---- e2e: host buffers through pre3_sequence, this rank's block; records gathered to every rank ------
    Pe = min(Pl, max(1, args.e2e_pairs // world))
    host = {k: torch.empty(data[k][:Pe + 1].shape, dtype=data[k].dtype, pin_memory=True) for k in data}
    for k in host:
        host[k].copy_(data[k][:Pe + 1])
    torch.cuda.synchronize()
    hn = {k: v.numpy() for k, v in host.items()}
    out = (np.zeros(Pe, pre3.RESULT_DTYPE), np.zeros((Pe, K_FEAT, 2), np.int32), np.zeros((Pe, K_FEAT), np.uint8))
    rec_pin = torch.empty(Pe, 240, dtype=torch.uint8, pin_memory=True)
    rec_dev = torch.empty(Pe, 240, dtype=torch.uint8, device=dev)
    rec_all = torch.empty(Pe * world, 240, dtype=torch.uint8, device=dev)
    ectx = pre3.Context(local)

    def e2e_step():
        o = ectx.sequence(hn["desc"], hn["xyz"], opts, pair_id0=p0, out=out)
        if world > 1:  # every rank ends up with the records of all pairs on the HOST
            rec_pin.numpy()[:] = np.frombuffer(out[0].tobytes(), np.uint8).reshape(Pe, 240)
            rec_dev.copy_(rec_pin, non_blocking=True)
            dist.all_gather_into_tensor(rec_all, rec_dev)
            return rec_all.cpu()
        return o

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    tb0 = ectx.transfer_bytes()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * Pe * e2e_steps / float(te.item())
    tb1 = ectx.transfer_bytes()
    hb = torch.tensor([(tb1[0] - tb0[0]) // e2e_steps, (tb1[1] - tb0[1]) // e2e_steps,
                       sum(int(hn[k].nbytes) for k in hn)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)   # bytes of the whole job per step
    h2d, d2h, host_in = (int(x) for x in hb.tolist())
    assert np.array_equal(out[0]["best_fit"], rl["best_fit"][:Pe]), "e2e path disagrees with the device path"
    ectx.close()
    e2e_gbs = (h2d + d2h) / (float(te.item()) / e2e_steps) / 1e9

    others = None
    if not args.no_other:
        del data, host, hn, res, matches, masks
        torch.cuda.empty_cache()
        others = other_workloads(ctx, pre3, synth, dev, rank, world)
        barrier()
    if rank == 0:
        cpu = cpu_baseline_single()
        other_cfg = {}
        if others:
            o = others
            other_cfg = {
                "cfg1_latency": _summ(o.get("cfg1", {}), ["k5.ms_per_pair_device_resident", "k5.ms_per_pair_host_buffers",
                                                          "k3.ms_per_pair_device_resident"]),
                "cfg2_matching_256x2048x2048": _summ(o.get("cfg2", {}), ["pairs_per_s", "ms_per_step", "roofline.achieved",
                                                                         "roofline.frac", "uint8.pairs_per_s",
                                                                         "uint8.ms_per_step", "uint8.roofline.achieved"]),
                "cfg4_ekf_200_features": _summ(o.get("cfg4", {}), ["adaptive.frames_per_s", "fixed_H.frames_per_s",
                                                                   "fixed_H.hyp_x_feature_evals_per_s",
                                                                   "fixed_H.roofline_score.frac", "fixed_H.roofline_gain.frac"]),
                "cfg5_20k_x_1M_split": _summ(o.get("cfg5", {}), ["first.ms_per_solve", "reference.ms_per_solve",
                                                                  "first.hyp_x_match_evals_per_s", "roofline.frac",
                                                                  "roofline.evals_per_s", "error"]),
                "dr_ye": _summ(o.get("dr_ye", {}), ["pairs_per_s"]),
                "frames": _summ(o.get("frames", {}), ["maps.roofline.frac", "fused_features.roofline.frac"]),
                "ekf_update": _summ(o.get("ekf_update", {}), ["frames_per_s", "fp64_tflops_executed"]),
            }
            for nm, key in (("cfg2", "cfg2_match_tc"), ("cfg5", "cfg5_eval")):
                rf = (o.get(nm) or {}).get("roofline")
                if rf:
                    roofline["per_kernel"][key] = {kk: rf.get(kk) for kk in ("bound", "achieved", "peak", "unit", "frac")}
            rf4 = ((o.get("cfg4") or {}).get("fixed_H") or {}).get("roofline_score")
            if rf4:
                roofline["per_kernel"]["cfg4_ekf_score"] = {kk: rf4.get(kk) for kk in ("bound", "achieved", "peak", "unit", "frac")}
        line = {
            "metric": "SR4000 frame-pairs/s (match+RANSAC)", "value": value, "unit": "frame-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(P=P), "pairs_total": P, "pairs_per_gpu": Pl,
                       "sharding": "by frame pair: rank r takes pairs dist.split_range(P, r, N); the 240-byte records of "
                                   "all pairs are all-gathered to every rank inside the timed region",
                       "api": "pre3_sequence_dev with pre3_set_graphs(1) (value) / pre3_sequence (e2e): pair p = (frame p, "
                              "frame p+1)",
                       "l2": f"inputs ({desc_bytes / 1e9:.2f} GB of descriptors per GPU) "
                             + ("larger than L2" if desc_bytes > 126e6 else "smaller than L2: at this GPU count the step "
                                "re-reads them from L2"),
                       "match_engine": "tcgen05 CTA-pair proposal (conversion fused into the GEMM kernel) + exact rescore" if "match_fused" in per_kernel
                       else "tcgen05 CTA-pair proposal + exact rescore" if "match_tc" in per_kernel
                       else "exact fp64 brute force",
                       "weak_scaling": weak, "other_configs": other_cfg},
            "e2e": {"value": e2e_val, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pairs_per_step": Pe * world, "steps": e2e_steps, "host_input_bytes_per_step": host_in,
                    "pcie_gb_per_s_whole_job": e2e_gbs,
                    "bound": "host side: one PCIe link per GPU (~55 GB/s) at N = 1; with several ranks on one host the "
                             "shared host DRAM / root complexes (~180 GB/s aggregate measured on this pool) cap the job",
                    "note": "pre3_sequence on pinned host buffers (class double), every rank its block of the sequence, "
                            "records all-gathered and read back on every rank. With one rank per host, descriptors whose "
                            "values survive (double)(float)x == x are narrowed by a host thread pool and cross PCIe as "
                            "float (h2d_bytes_per_step < host_input_bytes_per_step); with several ranks per host they "
                            "cross as doubles (every GPU has its own link, the host memory is shared)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "rooflines": rooflines,
            "kernels": per_kernel,
            "evals": {"hyp_x_match_evals_per_s": world * evals_done / (ms_step * 1e-3),
                      "executed_per_step_this_rank": evals_done, "required_by_reference_loop_per_step_this_rank": evals_needed},
            "cpu_baseline": cpu,
            "check": {"pairs_solved": ok_pairs, "pairs": P},
            "other_workloads": others,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL's version
    banner, warnings) was redirected to stderr in main()."""
    txt = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(txt)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, txt.encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of native libraries (NCCL banner) must not precede the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4096, help="frame pairs of the whole sequence (split over the GPUs)")
    ap.add_argument("--e2e-pairs", type=int, default=4096, help="frame pairs per e2e step, whole job (pinned host memory)")
    ap.add_argument("--no-other", action="store_true", help="skip the short cfg2 / cfg4 / cfg5 measurements")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: skips the e2e and cpu_baseline legs (their keys are null)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
