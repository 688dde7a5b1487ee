#!/usr/bin/env python
"""Headline benchmark: PCG audio-seconds preprocessed + augmented per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], named in `config.workload`): 1024 synthetic Training-A-shaped recordings per
GPU, two channels (PCG + ECG), 30 s at 2 kHz -> resample to 4125 Hz, Schmidt despike (PCG), 25-450 Hz / 2-40 Hz
band (fs-normalised, as the reference does), abs-max normalise, 4 s windows with 0.25 s overlap; then the torchaug
chain (augment_pcg_batch, default AugmentConfig: noise, wandering volume, parametric EQ, noise, each behind its per-row
mask) on the 7168 PCG windows that come out.  A step = one pass of the hot path over that batch: two launches of
this repo's kernels (fused preprocess+segment, fused augmentation chain).

  value      whole-job audio-s/s with inputs already resident in HBM (two kernel launches per step),
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e        same metric through HostPipeline: pinned HOST buffers in, HOST buffers out, copies inside the
             timed region.
  roofline   the dominant kernel (fused preprocess+segment, ~70 % of a step): its algorithmic bytes (input read once +
             windows written once) / its own duration (CUDA events around each of its launches inside the timed
             region), against the measured HBM peak in MEASURED_PEAKS.json.  `stages` holds both kernels' times.
  other_paths   (N=1, outside the timed region) the fused preprocess kernel on configs[0]-shaped rows (480 000 samples at
             16 kHz, the 64 recordings replicated 16x), the fused augmentation chain (configs[2]) with and without
             `collapse`, the tensor-core log-mel (configs[3]), and `reference_gpu`: the reference's own tensor path
             (torchaudio resample / lfilter = iir_cu_kernel, Python despike loop) on this GPU for a sub-batch.
  cpu_baseline  the NumPy oracle (the reference's CPU algorithm, restated in oracle/numpy_path.py) timed on the
             box's host cores on a bounded sample of the same workload (rank 0, N=1 only): all cores, plus the
             single-core figure with its per-stage split.

--impl reference times that CPU implementation alone (all host threads) and prints the same line shape.
Multi-GPU: recordings are sharded by index, no collective on the data path ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RECORDINGS = 1024           # per GPU
T_IN = 60000                # 30 s at 2 kHz
FS_IN, FS_OUT = 2000, 4125
SECONDS = T_IN / FS_IN
WINDOW_S = 4.0
KINDS = ("pcg", "ecg")
METRIC = "PCG audio-seconds preprocessed+augmented/sec"
UNIT = "audio-s/s"
WORKLOAD = ("configs[1]: Training-A-shaped PCG+ECG two-channel preprocessing, 1024 synthetic 30 s recordings per GPU, "
            "2 kHz -> 4125 Hz, despike (PCG), PCG 25-450 Hz / ECG 2-40 Hz band, abs-max norm, 4 s / 0.25 s-overlap windows, "
            "then augment_pcg_batch (default AugmentConfig) on the 7168 PCG windows per GPU")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "fused_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


# ----------------------------------------------------------------------------------------- CPU arm
def _cpu_draws(rng, rows, t):
    """Random draws of one augment_pcg_batch call in the oracle's injected form (default AugmentConfig probabilities)."""
    import numpy as np
    import torch
    u = lambda *s: torch.from_numpy(rng.random(s))
    bands = []
    for _ in range(5):
        lo = float(rng.uniform(2, 0.95 * 500))
        bands.append((lo, float(rng.uniform(lo + 0.05 * 498, 500))))
    d = {"std1": float(rng.choice([1e-4, 1e-3, 1e-2])), "scale1": u(rows, 1) * 0.1, "noise1": torch.from_numpy(rng.standard_normal((rows, t))),
         "mask1": (u(rows, 1) < 0.3 / 4).double(), "amp": 0.01 + u(rows, 2) * 0.24, "phase": u(rows, 2),
         "freq": torch.stack([0.05 + u(rows) * 0.45, 0.001 + u(rows) * 0.049], dim=1), "mask2": (u(rows, 1) < 0.75).double(),
         "bands": bands, "mask3": (u(rows, 1) < 0.25).double(), "std2": float(rng.choice([1e-4, 1e-3, 1e-2])),
         "scale2": u(rows, 1) * 0.1, "noise2": torch.from_numpy(rng.standard_normal((rows, t))), "mask4": (u(rows, 1) < 0.3 / 4).double()}
    return d


def _cpu_one(args):
    """One recording through the reference's CPU algorithms (oracle ports): NumPy preprocessing of both channels +
    windows, then the torchaug chain on the PCG windows (float64, as the parity tests run it)."""
    import numpy as np
    import torch
    from oracle import numpy_path as onp
    from oracle import torch_path as otp
    pcg, ecg, seed = args
    spec = onp.WindowSpec(WINDOW_S)
    p = onp.preprocess_pcg(pcg, FS_IN, FS_OUT)
    e = onp.preprocess_ecg(ecg, FS_IN, FS_OUT)
    w = onp.segment(np.stack([p, e], axis=1), FS_OUT, spec)                  # [N, win, 2]
    wp = torch.from_numpy(np.ascontiguousarray(w[:, :, 0]))
    aug = otp.augment_pcg_batch(wp, FS_OUT, _cpu_draws(np.random.default_rng(seed), wp.shape[0], wp.shape[1]))
    return w.shape[0] + int(aug.shape[0])


def cpu_throughput(n_recordings: int, cores: int, seed: int = 1234, steps: int = 1, warmup: int = 0):
    """audio-s/s of the oracle port over `n_recordings` synthetic recordings per step using `cores` worker processes
    (one pool for all steps).  Returns (audio-s/s over the timed steps, seconds of the timed steps)."""
    import multiprocessing as mp
    from wav2vec_heart_sounds_b200.synth import synth_pair
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    x = synth_pair(n_recordings, T_IN, FS_IN, seed=seed).numpy()
    work = [(x[i, 0], x[i, 1], seed + i) for i in range(n_recordings)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, work[:cores])                     # warm the workers (imports, FFT plans)
        for _ in range(warmup):
            pool.map(_cpu_one, work, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_one, work, chunksize=1)
        dt = time.perf_counter() - t0
    return n_recordings * SECONDS * steps / dt, dt


def cpu_single_core(n_recordings: int = 6, seed: int = 4321):
    """One process, one thread: the configs[1] preprocessing of `n_recordings` recordings stage by stage."""
    import numpy as np
    from oracle import numpy_path as onp
    from wav2vec_heart_sounds_b200.synth import synth_pair
    x = synth_pair(n_recordings, T_IN, FS_IN, seed=seed).numpy().astype(np.float64)
    spec = onp.WindowSpec(WINDOW_S)
    st = {"resample": 0.0, "despike": 0.0, "band": 0.0, "normalise": 0.0, "segment": 0.0}
    t_all = time.perf_counter()
    for i in range(n_recordings):
        cols = []
        for ch, band in ((0, onp.PCG_BAND), (1, onp.ECG_BAND)):
            t0 = time.perf_counter(); v = onp.resample(onp.fill_nans(x[i, ch]), FS_IN, FS_OUT); st["resample"] += time.perf_counter() - t0
            if ch == 0:
                t0 = time.perf_counter(); v = onp.remove_spikes(v, FS_OUT); st["despike"] += time.perf_counter() - t0
            t0 = time.perf_counter(); v = onp.bandpass_cascade(v, FS_OUT, *band); st["band"] += time.perf_counter() - t0
            t0 = time.perf_counter(); v = onp.abs_max_normalise(v); st["normalise"] += time.perf_counter() - t0
            cols.append(v)
        t0 = time.perf_counter(); onp.segment(np.stack(cols, axis=1), FS_OUT, spec); st["segment"] += time.perf_counter() - t0
    dt = time.perf_counter() - t_all
    return {"value": n_recordings * SECONDS / dt, "unit": UNIT, "cores": 1, "recordings": n_recordings,
            "what": "NumPy/SciPy float64 preprocessing + windows of both channels (no augmentation), OMP_NUM_THREADS=1",
            "ms_per_recording": {k: 1e3 * v / n_recordings for k, v in st.items()}}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # bounded sample: K + W steps must fit in ~4 minutes of CPU wall clock at ~0.35 core-seconds per recording; whole rounds of
    # the worker pool so that no core idles at the end of a step
    per_step = int(240.0 / (args.steps + args.warmup) * cores / 0.35)
    per_step = max(cores, min(per_step, RECORDINGS))
    per_step = min(RECORDINGS, -(-per_step // cores) * cores)
    torch.set_num_threads(1)
    value, t_all = cpu_throughput(per_step, cores, steps=args.steps, warmup=args.warmup)
    sample = (f"{per_step} recordings per step (2 channels x 30 s; NumPy/SciPy float64 preprocessing + windows, then the "
              f"torchaug chain on the PCG windows in float64 on CPU), {cores} worker processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample, "recordings_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import wav2vec_heart_sounds_b200 as pkg
    from wav2vec_heart_sounds_b200.pipeline import HostPipeline
    from wav2vec_heart_sounds_b200.synth import synth_pair
    spec = pkg.WindowSpec(WINDOW_S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # shard `rank` of the job: its own 1024 recordings, generated on the device (no host or peer traffic)
    from wav2vec_heart_sounds_b200 import AugmentConfig, torchaug
    import numpy as np
    cfg = AugmentConfig()
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    x = synth_pair(RECORDINGS, T_IN, FS_IN, seed=1234 + rank, device=dev)
    # channel-major windows [2, B, N, win]: the PCG windows are one contiguous [B * N, win] batch for the augmentation
    out = pkg.preprocess_segment(x, FS_IN, FS_OUT, spec, kinds=KINDS, fused=True, channel_major=True)   # also a warm-up
    n_win, win = out.shape[2], out.shape[3]
    pcg_windows = out[0].view(RECORDINGS * n_win, win)
    aug_out = torch.empty_like(pcg_windows)
    algo_bytes = x.numel() * 4 + out.numel() * 4
    aug_bytes = 2 * pcg_windows.numel() * 4
    ev = []

    def step(timed=False):
        if timed:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        pkg.preprocess_segment(x, FS_IN, FS_OUT, spec, kinds=KINDS, fused=True, channel_major=True, out=out)
        if timed:
            e1.record()
        torchaug.augment_pcg_batch(pcg_windows, FS_OUT, cfg, noise="philox", fused=True, out=aug_out)
        if timed:
            e2.record()
            ev.append((e0, e1, e2))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        start.record()
        for _ in range(args.steps):
            step(timed=True)
        stop.record()
        barrier()
    ms = start.elapsed_time(stop)
    pre_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps         # the fused preprocess kernel alone
    aug_ms = sum(b.elapsed_time(c) for _, b, c in ev) / args.steps         # the augmentation chain (its small RNG launches included)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * RECORDINGS * SECONDS * args.steps / (ms * 1e-3)
    step_ms = ms / args.steps
    kernel_ms = pre_ms

    # ---- end to end: host buffers in, host buffers out
    hp = HostPipeline(RECORDINGS, 2, T_IN, FS_IN, FS_OUT, spec, kinds=KINDS, chunk=64, device=dev, augment=cfg)
    x_host = x.cpu().pin_memory()
    out_host = hp.empty_output()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        hp(x_host, out_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hp(x_host, out_host)
        torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * RECORDINGS * SECONDS * e2e_steps / float(t.item())
    same = bool(torch.equal(out_host[1, :8], out[1, :8].cpu()))               # the ECG windows (the PCG ones carry fresh random draws)
    # the ceiling of that leg on this box: the same bytes as bare pinned copies (one cudaMemcpyAsync per chunk and
    # direction, both directions at once, every rank at the same time), no kernel in between
    copy_s = bare_copy_seconds(x_host, out_host, dev, barrier, chunk=64)
    t = torch.tensor([copy_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    copy_value = world * RECORDINGS * SECONDS / float(t.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 samples, f64 filter/normalise state", "data": "synthetic",
                "config": {"workload": WORKLOAD, "recordings_per_gpu": RECORDINGS, "channels": list(KINDS),
                           "layout": f"[{RECORDINGS}, 2, {T_IN}] -> [2, {RECORDINGS}, {n_win}, {win}] -> PCG [{RECORDINGS * n_win}, {win}] augmented",
                           "mode": "torch", "augment": "default AugmentConfig, in-kernel Philox noise, fresh draws every step",
                           "cache": "inputs (492 MB), windows (946 MB) and augmented windows (473 MB) per step exceed the 126 MB L2",
                           "sharding": f"{world} x index-sharded, no collective"},
                "gpu_launches": 3 * args.steps,          # fused preprocess, per-row draws, augmentation chain
                "stages": {"preprocess_segment_ms": pre_ms, "augment_chain_ms": aug_ms,
                           "preprocess_only_audio_s_per_s": RECORDINGS * SECONDS / (pre_ms * 1e-3),
                           "augment_GB/s": aug_bytes / (aug_ms * 1e-3) / 1e9},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes,
                        "d2h_bytes_per_step": hp.d2h_bytes, "steps": e2e_steps, "matches_device_run": same,
                        "bare_copy_ceiling": copy_value, "frac_of_copy_ceiling": e2e_value / copy_value},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": recorded_traffic(), "algorithmic_bytes_per_launch": algo_bytes,
                             "kernel": "fused_stream_kernel<33,16,30,4>", "share_of_step": pre_ms / step_ms,
                             "peak_source": peak_src},
                "clocks": clocks.summary()}
        if world == 1 and not args.no_extras:
            line["other_paths"] = other_paths(dev, peak)
        if world == 1 and not args.no_cpu:
            # a fresh interpreter: the CPU leg forks worker processes that run torch CPU ops, which must not inherit
            # this process's CUDA context and thread pools
            import subprocess
            cores = host_cores()
            n = min(RECORDINGS, 32 * cores)                   # ~10 s of wall clock on the host cores
            try:
                proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-leg", str(n)], capture_output=True,
                                      text=True, timeout=240, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
                leg = json.loads(proc.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = {"value": leg["value"], "unit": UNIT, "cores": cores, "kind": "port", "single_core": leg.get("single_core"),
                                        "sample": f"{n} of the {RECORDINGS} recordings, NumPy/SciPy float64 oracle "
                                                  f"(oracle/numpy_path.py) + float64 torchaug chain (oracle/torch_path.py), "
                                                  f"{cores} worker processes, {leg['seconds']:.1f} s"}
            except Exception as exc:                          # never lose the GPU line over the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"CPU leg failed: {type(exc).__name__}: {exc}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bare_copy_seconds(x_host, out_host, dev, barrier, chunk=64, reps=3):
    """Seconds per step of the e2e leg's PCIe traffic alone: its H2D and D2H bytes as plain pinned cudaMemcpyAsync
    calls in the same chunking, the two directions on two streams at once, nothing else on the device."""
    n = x_host.shape[0]
    d_in = torch.empty((chunk,) + tuple(x_host.shape[1:]), device=dev)
    d_out = torch.empty((out_host.shape[0], chunk) + tuple(out_host.shape[2:]), device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        for lo in range(0, n, chunk):
            hi = min(lo + chunk, n)
            with torch.cuda.stream(s_in):
                d_in[: hi - lo].copy_(x_host[lo:hi], non_blocking=True)
            with torch.cuda.stream(s_out):
                for c in range(out_host.shape[0]):
                    out_host[c, lo:hi].copy_(d_out[c, : hi - lo], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)       # (first pass is the warm-up)
    return best


def other_paths(dev, peak):
    """Outside the timed region, N=1 only: kernel times of the two other hot-path families at BASELINE.json's
    configs[2] and configs[3] shapes (CUDA events, best of 5, inputs resident), for the record beside the headline."""
    import numpy as np
    from wav2vec_heart_sounds_b200 import torchaug as ta, AugmentConfig, MelConfig, log_mel, _lib, design

    def best_ms(fn, reps=5):
        fn(); fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    res = {}
    import wav2vec_heart_sounds_b200 as pkg
    from wav2vec_heart_sounds_b200.synth import synth_pcg, synth_pair
    # configs[0]-shaped rows through the fused kernel: 64 recordings of 30 s at 2 kHz -> 16 kHz (480 000 samples per row,
    # 26 tiles streamed by one CTA), despike, 25-450 Hz band, abs-max norm, 4 s windows; replicated 16x so that the
    # launch fills the machine
    x0 = synth_pcg(64, T_IN, float(FS_IN), seed=11, device=dev).repeat(16, 1)
    spec4 = pkg.WindowSpec(4.0)
    out0 = pkg.preprocess_segment(x0, FS_IN, 16000, spec4, fused=True)
    ms = best_ms(lambda: pkg.preprocess_segment(x0, FS_IN, 16000, spec4, fused=True, out=out0))
    nbytes = 4 * (x0.numel() + out0.numel())
    res["config0"] = {"workload": "configs[0] x16: 1024 single-channel PCG rows, 30 s @2 kHz -> 16 kHz (480000 samples), despike, "
                                  "25-450 Hz band, abs-max norm, 4 s / 0.25 s-overlap windows, one fused launch",
                      "ms": ms, "audio_s_per_s": x0.shape[0] * SECONDS / (ms * 1e-3), "algorithmic_bytes": nbytes,
                      "GB/s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                      "target_audio_s_per_s_at_0.6": 0.6 * peak * 1e9 / 67733.0}
    del x0, out0
    # the reference's own tensor path on this GPU (torchproc arithmetic: torchaudio resample + lfilter -> iir_cu_kernel,
    # the Python despike loop with its host syncs), float32, on a sub-batch of the headline workload; ours on the same rows
    try:
        from oracle import torch_path as otp
        from oracle import numpy_path as onp
        nsub = 32
        xs = synth_pair(nsub, T_IN, FS_IN, seed=1234, device=dev)

        def ref_gpu():
            p = otp.preprocess_pcg(xs[:, 0], FS_IN, FS_OUT)
            e = otp.preprocess_ecg(xs[:, 1], FS_IN, FS_OUT)
            return otp.segment(p, FS_OUT, onp.WindowSpec(WINDOW_S)), otp.segment(e, FS_OUT, onp.WindowSpec(WINDOW_S))
        ref_gpu()                                             # warm-up (cuDNN / kernel selection); its despike loop takes seconds
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref_gpu()
        torch.cuda.synchronize()
        ref_ms = 1e3 * (time.perf_counter() - t0)
        outs = pkg.preprocess_segment(xs, FS_IN, FS_OUT, spec4, kinds=KINDS, fused=True)
        our_ms = best_ms(lambda: pkg.preprocess_segment(xs, FS_IN, FS_OUT, spec4, kinds=KINDS, fused=True, out=outs))
        res["reference_gpu"] = {"workload": f"{nsub} recordings x (PCG, ECG) of the headline workload, preprocessing + windows, float32, "
                                            "the reference's tensor-path arithmetic (torchaudio resample/lfilter kernels, Python "
                                            "despike loop) on this GPU vs one fused launch of this repo",
                                "reference_ms": ref_ms, "ours_ms": our_ms, "ours_over_theirs": ref_ms / our_ms,
                                "reference_audio_s_per_s": nsub * SECONDS / (ref_ms * 1e-3)}
        del xs, outs
    except Exception as exc:                                  # never lose the line over a side measurement
        res["reference_gpu"] = {"error": f"{type(exc).__name__}: {exc}"}
    # configs[2]: full torchaug chain on 4096 windows x 64000 samples @ 16 kHz, in-kernel Philox noise, default masks
    b, t, fs = 4096, 64000, 16000
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(b, t, device=dev, generator=g)
    torch.manual_seed(7); np.random.seed(7)
    cfg = AugmentConfig()
    rowp1, _ = ta._draw_noise(x, None, None, "philox"); m1 = ta._mask(b, cfg.prob_noise / 4, dev).reshape(b)
    rowp2 = ta._draw_sines(x, 0.24); m2 = ta._mask(b, cfg.prob_wandering_volume, dev).reshape(b)
    bands = ta._draw_bands(2, 500, 5); m3 = ta._mask(b, cfg.prob_banding, dev).reshape(b)
    rowp4, _ = ta._draw_noise(x, None, None, "philox"); m4 = ta._mask(b, cfg.prob_noise / 4, dev).reshape(b)
    sos = np.ascontiguousarray(design.eq_band_sos(fs, bands), dtype=np.float64)
    y = torch.empty_like(x)
    work = _lib.aug_workspace(x)

    def chain():
        _lib.check(_lib.lib().mpcg_aug_chain_f32(x.data_ptr(), y.data_ptr(), b, t, float(fs), rowp1.data_ptr(), None,
                                                 m1.data_ptr(), 1, 2, rowp2.data_ptr(), m2.data_ptr(), sos.ctypes.data,
                                                 sos.shape[0], m3.data_ptr(), rowp4.data_ptr(), None, m4.data_ptr(), 3, 4,
                                                 1, work.data_ptr(), work.numel(), torch.cuda.current_stream().cuda_stream),
                   "aug chain")
    ms = best_ms(chain)
    nbytes = 2 * b * t * 4
    res["augment_chain"] = {"workload": "configs[2]: augment_pcg_batch, 4096 windows x 64000 samples @16 kHz, one fused "
                                        "launch (mpcg_aug_chain_f32), Philox noise, default masks",
                            "ms": ms, "windows_per_s": b / (ms * 1e-3), "audio_s_per_s": b * 4.0 / (ms * 1e-3),
                            "algorithmic_bytes": nbytes, "GB/s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
    ms_api = best_ms(lambda: ta.augment_pcg_batch(x, fs, cfg, noise="philox"))
    res["augment_chain"]["ms_through_python_api"] = ms_api
    # collapse=False re-normalises masked-off rows too, as the reference does (torchaug.py:240-243); same result within 1e-5
    ms_nc = best_ms(lambda: ta.augment_pcg_batch(x, fs, cfg, noise="philox", collapse=False))
    res["augment_chain"]["ms_through_python_api_collapse_off"] = ms_nc
    # configs[2] in full: the composed pipeline of augment/pipelines.py:43-61 on the same 4096 windows -- min-max, HPSS
    # recombination (p = 0.75; transform sizes fixed at a middle setting of the reference's ranges so that the figure is
    # reproducible), noise, micro-stretch, wandering volume, noise, EQ, abs-max -- per-row masks, default probabilities
    try:
        import random as _random
        from wav2vec_heart_sounds_b200 import pipelines as pl
        _random.seed(7); np.random.seed(7)
        hp = dict(n_fft1=1024, hop1=64, n_fft2=1024, hop2=64, margin1=(1.5, 1.5), margin2=(2.5, 2.5), kernel1=(17, 17),
                  kernel2=(17, 17), w1=[1.0, 2.0, 3.0, 4.0], w2=[4.0, 3.0, 2.0, 1.0], w_mix=0.03)
        sub = x[:1024]

        def full():
            return pl.augment_pcg(sub, fs, cfg, draws={"hpss": hp})
        full(); torch.cuda.synchronize()
        t0 = time.perf_counter(); full(); torch.cuda.synchronize()
        ms_full = 1e3 * (time.perf_counter() - t0)
        res["augment_full_chain"] = {"workload": "configs[2] full chain incl. HPSS: pipelines.augment_pcg on 1024 of the 4096 windows x 64000 "
                                                 "@16 kHz (HPSS n_fft 1024 / hop 64 / medians 17, p = 0.75; wall clock incl. host orchestration)",
                                     "ms_per_1024_windows": ms_full, "windows_per_s": 1024 / (ms_full * 1e-3),
                                     "ms_for_4096_windows_extrapolated": 4 * ms_full}
    except Exception as exc:
        res["augment_full_chain"] = {"error": f"{type(exc).__name__}: {exc}"}
    del x, y
    # configs[3]: log-mel conditioning of 8192 windows x 64000 samples @ 16 kHz on the tensor-core tier
    xm = torch.randn(8192, 64000, device=dev, generator=g)
    tr = MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build(fast=True)
    ms = best_ms(lambda: log_mel(xm, tr))
    nbytes = 8192 * 64000 * 4 + 8192 * 80 * 251 * 4
    flops = 3 * 2.0 * (8192 * 254) * 256 * 80                     # three split-fp16 MMAs over [hop rows x hop] . [hop x 80]
    res["log_mel"] = {"workload": "configs[3] (mel part): log_mel of 8192 windows x 64000 @16 kHz, n_fft 1024, hop 256, "
                                  "80 mels, " + tr.backend,
                      "ms": ms, "windows_per_s": 8192 / (ms * 1e-3), "algorithmic_bytes": nbytes,
                      "GB/s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                      "tensor_TFLOP/s": flops / ms / 1e9}
    # the default tier of MelConfig.build(): float64 on the fp64 tensor path (DMMA), within 1e-5 on any input
    tr64 = MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build()
    ms64 = best_ms(lambda: log_mel(xm, tr64), reps=3)
    fma64 = 8192.0 * 256 * 1024 * 2 * tr64.kpad                  # frames padded to whole 32-frame tiles x samples x (cos, sin) x bins
    res["log_mel_default"] = {"workload": "same input, MelConfig.build() default: " + tr64.backend,
                              "ms": ms64, "windows_per_s": 8192 / (ms64 * 1e-3), "fp64_TFLOP/s": 2 * fma64 / ms64 / 1e9,
                              "GB/s": nbytes / ms64 / 1e6, "frac_of_hbm_peak": nbytes / ms64 / 1e6 / peak}
    del xm
    # row H: one HPSS split (STFT, two medians, masks + inverse transform) at the composed pipeline's transform size
    try:
        from wav2vec_heart_sounds_b200 import hpss as _hpss
        xh = torch.randn(256, 64000, device=dev, generator=g)
        ms_h = best_ms(lambda: _hpss.hpss_split(xh, 1024, 64, (1.5, 2.0), (17, 17)), reps=3)
        res["hpss_split"] = {"workload": "hpss_split of 256 windows x 64000 @16 kHz, n_fft 1024 / hop 64 / medians (17, 17)",
                             "ms": ms_h, "windows_per_s": 256 / (ms_h * 1e-3)}
        del xh
    except Exception as exc:
        res["hpss_split"] = {"error": f"{type(exc).__name__}: {exc}"}
    # generator conditioning: normalise, fade, fit, default log-mel, crop of 4096 x 32000 @4 kHz
    try:
        xs = torch.randn(4096, 32000, device=dev, generator=g)
        melc = MelConfig(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build()
        ms_c = best_ms(lambda: pkg.condition_generator_batch(xs, xs, 4000, melc, 96, 256), reps=3)
        res["condition_generator_batch"] = {"workload": "condition_generator_batch, 4096 x 32000 @4 kHz, DiffWave preset (127 weighted bins), "
                                                        + melc.backend, "ms": ms_c, "items_per_s": 4096 / (ms_c * 1e-3)}
        del xs
    except Exception as exc:
        res["condition_generator_batch"] = {"error": f"{type(exc).__name__}: {exc}"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the augmentation / mel side measurements")
    ap.add_argument("--cpu-leg", type=int, default=0, help=argparse.SUPPRESS)      # internal: time the CPU port on N recordings
    args = ap.parse_args()
    if args.cpu_leg > 0:
        torch.set_num_threads(1)
        v, dt = cpu_throughput(args.cpu_leg, host_cores())
        print(json.dumps({"value": v, "seconds": dt, "single_core": cpu_single_core()}))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
