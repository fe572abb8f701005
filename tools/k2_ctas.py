#!/usr/bin/env python
"""Per-CTA start / run times of k_detect on the benchmark scene (B2S_K2_DUMP_CTAS, profiling level 2). Measurement helper.
Usage: python tools/k2_ctas.py [quiet]   ("quiet": the same push without carriers)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2S_K2_DUMP_CTAS"] = "1"
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402


def main():
    import torch

    b2s, synth = ge.load_b2s(), ge.load_synth()
    n, fs, T = 16384, 20_000_000, 4096
    dev = torch.device("cuda", 0)
    eng = b2s.Engine(0)
    tones = [] if "quiet" in sys.argv else bench.bench_tones(synth, n, T, bench.LEARN)
    iq = synth.make_iq_int8_torch(n, T, tones, seed=synth.seed_for(2, 0), quiet_frames=bench.LEARN, device=dev)
    cfg = b2s.make_config(n, fs, learn_frames=bench.LEARN, max_frames_per_push=T, flags=b2s.FLAG_IQ_ON_DEVICE | b2s.FLAG_ASYNC)
    band = b2s.Band(eng, cfg)
    band.set_profiling(2)
    period = synth.frame_period_ms(n, fs)
    t = 0.0
    for i in range(4):
        band.push_raw(iq.data_ptr(), T, int(t), period)
        t += T * period
    band.sync()
    p = band.get_profile(reset=True)
    print("k2 per launch ms", p.detect_ms / p.detect_launches, "cta median", p.detect_cta_median_ms / p.detect_launches, "max", p.detect_cta_max_ms / p.detect_launches)
    band.close()


if __name__ == "__main__":
    main()
