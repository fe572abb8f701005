#!/usr/bin/env python
"""bench.py — IQ MSamples/s through the fused unpack+FFT+power+detect path (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md §8d "Config 2"): ONE 20 MS/s band, N = 16384-point FFT, r = 1,
T = 4096 frames per step = 67.1 M complex samples = 134 MB of int8 IQ (> the 126 MB L2, so every step streams its
input from HBM; no L2 flush is needed). A step is one b2s_band_push of T frames: K1 (unpack, window, FFT, dB),
K2 (noise, Averager, boxcar, threshold, spectrogram) and the host-side signal bookkeeping of the detections. The band
runs in its asynchronous result mode (B2S_FLAG_ASYNC): the bookkeeping of step k overlaps the kernels of step k+1, as the
reference's mailbox does; the timed region ends with b2s_band_sync, i.e. after ALL work of all K steps.

  value : steady-state throughput with the IQ already resident in HBM (B2S_FLAG_IQ_ON_DEVICE), CUDA-event timed.
  e2e   : the same call with the IQ in pinned HOST memory: the host->device copy of every step's input and the
          device->host read of its results are inside the timed region.
  roofline : dominant kernel (K1 k_spectrum3): algorithmic bytes 6 B/sample (2 B int8 IQ read + 4 B fp32 dB row written,
          SURVEY.md §8d) x T x N per launch / that kernel's mean launch time (CUDA events inside the library, on the
          launching stream) vs. the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port (fp32, oracle/liboracle.so — FFTW itself is not available here) on a bounded
          sample of the same workload, all host cores.
N > 1 GPUs (torchrun): one process and one 20 MS/s band per GPU, no data-path collective ("weak" scaling); the time
is the max over ranks (device-timed), value = N x samples / that.

`--impl reference` times the reference's CPU path instead (the oracle port; the reference's own GNU Radio/FFTW chain
cannot be built in this image — see DESIGN.md), with all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_FFT = 16384
SAMPLE_RATE = 20_000_000
FRAMES = 4096
LEARN = 100
ALG_BYTES_PER_SAMPLE = 6.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_spectrum3<16> launch of this workload, from the committed
# `ncu --set full` capture profiles/r01_k1_v3_ncu_summary.txt (134.5 MB read + 210.2 MB written; part of the 268 MB of
# rows is still dirty in the 126 MB L2 when the kernel ends)
K1_DRAM_TRAFFIC_BYTES = 344.7e6
K1_TRAFFIC_SOURCE = "profiles/r01_k1_v3_ncu_summary.txt"
K2_ALG_BYTES_PER_SAMPLE = 4.0  # k_detect reads every fp32 dB row once (its outputs are sparse)
METRIC = "IQ MSamples/s through FFT+power+detect"


def bench_tones(synth, n_fft, frames, learn):
    """Four keyed NFM-like carriers (SURVEY.md §8d generator): starts, stops and time-outs all occur inside a step."""
    span = frames - learn
    a = learn
    return [
        synth.Tone(0.31 * n_fft / 2 + 0.1, amplitude=40.0, fm_dev_bins=5.0, on_frames=[(a + int(0.05 * span), a + int(0.60 * span))]),
        synth.Tone(-0.62 * n_fft / 2 + 0.1, amplitude=40.0, fm_dev_bins=5.0, on_frames=[(a + int(0.20 * span), a + int(0.35 * span)), (a + int(0.55 * span), a + int(0.90 * span))]),
        synth.Tone(0.055 * n_fft / 2 + 0.1, amplitude=40.0, fm_dev_bins=5.0, on_frames=[(a + int(0.10 * span), a + int(0.45 * span))], phase=1.0),
        synth.Tone(-0.17 * n_fft / 2 + 0.1, amplitude=40.0, fm_dev_bins=5.0, on_frames=[(a + int(0.40 * span), a + int(0.95 * span))], phase=2.0),
    ]


def bands_for_rank(n_bands: int, rank: int, world: int):
    """SURVEY.md §8e partitioning: band b lives on GPU b mod G; frames of a band never leave their GPU."""
    return [b for b in range(n_bands) if b % world == rank]


def reduce_step_time(ms_local: float, device=None) -> float:
    """Max over ranks of the device-timed region (the slowest rank defines the step); identity for one rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local)
    t = torch.tensor([ms_local], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_msps(samples_per_rank_step: int, steps: int, world: int, ms_max: float) -> float:
    """Whole-job throughput: all ranks' samples over the slowest rank's time."""
    return world * samples_per_rank_step * steps / (ms_max / 1000.0) / 1e6


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of the reference path (oracle port, fp32), all host threads."""
    if rank != 0:
        return
    import numpy as np
    import __graft_entry__ as ge
    import oracle_lib as ol

    b2s, synth = ge.load_b2s(), ge.load_synth()
    cores = os.cpu_count() or 1
    cfg = b2s.make_config(N_FFT, SAMPLE_RATE, learn_frames=LEARN)
    period = synth.frame_period_ms(N_FFT, SAMPLE_RATE)
    # bounded sample: `cores` independent segments of the step's frames, one chain per thread
    per_thread = 512
    frames = per_thread * cores
    tones = bench_tones(synth, N_FFT, per_thread, LEARN)
    seg = synth.make_iq_int8(N_FFT, per_thread, tones, seed=synth.seed_for(2), quiet_frames=LEARN)
    iq = np.tile(seg, cores)
    L = ol.oracle()
    times = []
    for i in range(args.warmup + args.steps):
        dt = L.orc_bench_run(C.byref(cfg), iq.ctypes.data_as(C.c_void_p), frames, period, cores)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    samples = frames * N_FFT * args.steps
    value = samples / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: single 20 MS/s band, 16384-pt FFT, r=1 (CPU arm: bounded sample)", "fft_size": N_FFT, "sample_rate_hz": SAMPLE_RATE,
                   "frames_per_step": frames, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": cores, "kind": "port",
                         "sample": f"{frames} frames ({cores} threads x {per_thread} frames, one chain per thread) of the configs[1] workload per step; restated CPU path, fp32 FFT (FFTW/GNU Radio unavailable)"},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2s", choices=["b2s", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per step (default = the BASELINE workload)")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2s" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    b2s, synth = ge.load_b2s(), ge.load_synth()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    T = args.frames
    period = synth.frame_period_ms(N_FFT, SAMPLE_RATE)
    tones = bench_tones(synth, N_FFT, T, LEARN)
    iq_dev = synth.make_iq_int8_torch(N_FFT, T, tones, seed=synth.seed_for(2, rank), quiet_frames=LEARN, device=dev)
    torch.cuda.synchronize()

    eng = b2s.Engine(local_rank)
    stream = torch.cuda.current_stream(dev)

    def make_band(flags):
        cfg = b2s.make_config(N_FFT, SAMPLE_RATE, center_hz=150_000_000 + 1_000_000 * rank, learn_frames=LEARN, max_frames_per_push=T, flags=flags)
        band = b2s.Band(eng, cfg)
        band.set_stream(stream.cuda_stream)
        band.set_profiling(True)
        return band

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(band, ptr, steps, warmup, label):
        res = b2s.Result()
        t_ms = 0
        for i in range(warmup):
            band.push_raw(ptr, T, int(t_ms), period, res)
            t_ms += T * period
        band.sync(res)
        band.get_profile(reset=True)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last = 0
        for i in range(steps):
            band.push_raw(ptr, T, int(t_ms), period, res)
            t_ms += T * period
        band.sync(res)  # async result mode: every push's kernels AND bookkeeping are complete before the clock stops
        last = res.n_transmissions
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        prof = band.get_profile(reset=True)
        return reduce_step_time(ms, dev), prof, clocks, last, res.n_detect_entries

    # ---- device-resident run (value + roofline) ----
    band = make_band(b2s.FLAG_IQ_ON_DEVICE | b2s.FLAG_ASYNC)
    ms, prof, clocks, n_tx, n_ent = timed(band, iq_dev.data_ptr(), args.steps, args.warmup, "device")
    samples_step = T * N_FFT
    value = aggregate_msps(samples_step, args.steps, world, ms)
    k1_ms = prof.spectral_ms / max(prof.spectral_launches, 1)
    peak, peak_src = measured_peaks()
    achieved = ALG_BYTES_PER_SAMPLE * samples_step / (k1_ms / 1000.0) / 1e9
    # kernels of this library per step: k_spectrum3, k_detect, k_entries_prefix, k_entries_sort (+ k_window_query when the tracker asks)
    launches = int(prof.spectral_launches + 3 * prof.detect_launches + prof.window_launches)
    k2_ms = prof.detect_ms / max(prof.detect_launches, 1)
    # K2's per-CTA balance: three extra untimed pushes with profiling level 2 (kept out of the timed region: it costs one small
    # device->host copy per push)
    band.set_profiling(2)
    res_b = b2s.Result()
    for i in range(3):
        band.push_raw(iq_dev.data_ptr(), T, int((args.warmup + args.steps + i) * T * period), period, res_b)
    band.sync(res_b)
    prof_b = band.get_profile(reset=True)
    band.close()

    # ---- end-to-end run: pinned host IQ, H2D inside the timed region ----
    e2e = None
    if not args.skip_e2e:
        host = torch.empty(iq_dev.numel(), dtype=torch.int8, pin_memory=True)
        host.copy_(iq_dev)
        torch.cuda.synchronize()
        band_h = make_band(b2s.FLAG_ASYNC)
        steps_e = max(3, min(args.steps, 10))
        ms_e, prof_e, _, _, _ = timed(band_h, host.data_ptr(), steps_e, 3, "e2e")
        e2e = {
            "value": aggregate_msps(samples_step, steps_e, world, ms_e), "unit": "MS/s",
            "h2d_bytes_per_step": int(prof_e.h2d_bytes // max(prof_e.pushes, 1)), "d2h_bytes_per_step": int(prof_e.d2h_bytes // max(prof_e.pushes, 1)),
            "steps": steps_e, "ms_per_step": ms_e / steps_e,
        }
        band_h.close()
        del host

    # ---- cpu baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle_lib as ol

        cores = os.cpu_count() or 1
        per_thread = 512
        frames_c = per_thread * cores
        seg = iq_dev[: per_thread * N_FFT * 2].cpu().numpy()
        iq_c = np.tile(seg, cores)
        cfg_c = b2s.make_config(N_FFT, SAMPLE_RATE, learn_frames=LEARN)
        L = ol.oracle()
        L.orc_bench_run(C.byref(cfg_c), iq_c.ctypes.data_as(C.c_void_p), frames_c, period, cores)  # warm-up
        reps, tot = 0, 0.0
        while tot < 10.0 and reps < 20:
            tot += L.orc_bench_run(C.byref(cfg_c), iq_c.ctypes.data_as(C.c_void_p), frames_c, period, cores)
            reps += 1
        cpu = {"value": frames_c * N_FFT * reps / tot / 1e6, "unit": "MS/s", "cores": cores, "kind": "port",
               "sample": f"{reps} x {frames_c} frames ({cores} threads x {per_thread} frames) of the step's IQ; restated CPU path with fp32 FFT (FFTW unavailable)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: single 20 MS/s band, 16384-pt FFT, fused unpack+FFT+power+detect", "fft_size": N_FFT, "sample_rate_hz": SAMPLE_RATE,
                       "frames_per_step": T, "bands_per_gpu": 1, "input": "int8 IQ (CS8)", "l2": "134 MB input per step > 126 MB L2 (no flush needed)",
                       "parallelism": f"bands sharded, {world} GPU(s), no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": "k_spectrum3<16> (N=16384)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": K1_DRAM_TRAFFIC_BYTES if T == FRAMES else None, "traffic_source": K1_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * samples_step, "kernel_ms": k1_ms,
                         "other_kernels_ms": {"k_detect+list_ordering": k2_ms, "k_window_query_total": prof.window_ms / args.steps,
                                              "host_tracker": prof.tracker_host_ms / args.steps},
                         "k_detect": {"achieved": K2_ALG_BYTES_PER_SAMPLE * samples_step / (k2_ms / 1000.0) / 1e9, "unit": "GB/s",
                                      "frac": K2_ALG_BYTES_PER_SAMPLE * samples_step / (k2_ms / 1000.0) / 1e9 / peak,
                                      "cta_median_ms": prof_b.detect_cta_median_ms / max(prof_b.detect_launches, 1),
                                      "cta_max_ms": prof_b.detect_cta_max_ms / max(prof_b.detect_launches, 1)}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "detections": {"transmissions_after_last_step": n_tx, "detect_entries_last_step": n_ent},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
