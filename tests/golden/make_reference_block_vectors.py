#!/usr/bin/env python
"""Generates tests/golden/reference_blocks_n512.npz from the REFERENCE'S OWN compiled blocks.

Run in the build container (needs /root/reference and `make -C oracle ref`):  python tests/golden/make_reference_block_vectors.py
The PSD rows of a small keyed-carrier scene (N = 512, fs = 200 kS/s so that the spectrogram decimates by 2, 25 ms frame clock,
220 frames) are pushed through the reference's NoiseLearner -> Transmission and Spectrogram -> DataController objects
(oracle/ref_blocks_shim.cpp, injected clock). Stored: the PSD rows (input), and what the reference objects produced — the
NoiseLearner rows, the per-frame FrequencyFlush lists and the published spectrogram payloads. The GPU box and any machine
without /root/reference can then pin the oracle (tests/test_oracle_golden_blocks.py) with no reference code present.
The PSD rows are stored rather than the IQ so that the fixture does not depend on a libm (log10f) bit pattern."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib as ol  # noqa: E402

b2s, synth = ge.load_b2s(), ge.load_synth()

N, FS, FRAMES, PERIOD_MS, T0 = 512, 200_000, 220, 25.0, 1_700_000_000_000
LEARN = b2s.lib().b2s_learn_frames_from_ms(2000, PERIOD_MS)  # NOISE_LEARNING_TIME (config.h:24) is compiled into the reference objects
BW = 16 * FS // N


def config():
    return b2s.make_config(N, FS, learn_frames=LEARN, recording_bandwidth_hz=BW, min_time_ms=200, timeout_ms=300)


def main():
    assert ol.have_ref_blocks(), "build oracle/_ref/libref.so first: make -C oracle ref"
    cfg = config()
    tones = synth.standard_scene(N, FRAMES, LEARN)
    iq = synth.make_iq_int8(N, FRAMES, tones, seed=synth.seed_for(3), quiet_frames=LEARN)
    psd = ol.OracleChain(cfg).push(iq, FRAMES, T0, PERIOD_MS, dense=("psd_db",)).psd_db
    ref = ol.RefBlocksChain(cfg, T0, BW, with_spectrogram=True)
    q = np.empty_like(psd)
    tx_count = np.zeros(FRAMES, dtype=np.int32)
    tx = np.zeros((FRAMES, 8, 2), dtype=np.int32)
    for k in range(FRAMES):
        now = T0 + int(np.floor(k * PERIOD_MS + 0.5))
        q[k], lst = ref.push_row(psd[k], now)
        tx_count[k] = len(lst)
        for i, (f, fl) in enumerate(lst[:8]):
            tx[k, i] = (f, fl)
    payloads = [p for t, p in ref.published() if t == "sdr/dev/spectrogram"]
    assert tx_count.max() <= 8 and tx_count.sum() > 100 and len(payloads) >= 4
    out = os.path.join(ROOT, "tests", "golden", "reference_blocks_n512.npz")
    np.savez_compressed(out, psd=psd, noise_sub=q, tx_count=tx_count, tx=tx, payloads=np.array([np.frombuffer(p, dtype=np.uint8) for p in payloads]),
                        meta=np.array([N, FS, FRAMES, LEARN, BW, T0], dtype=np.int64), period_ms=np.array([PERIOD_MS]))
    print("wrote", out, os.path.getsize(out), "bytes;", int(tx_count.sum()), "transmission records,", len(payloads), "spectrogram payloads")


if __name__ == "__main__":
    main()
