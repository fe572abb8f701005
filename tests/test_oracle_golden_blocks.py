"""Pins the oracle with a COMMITTED fixture produced by the reference's own compiled blocks (tests/golden/
make_reference_block_vectors.py, run where /root/reference exists): the PSD rows of a keyed-carrier scene in, and out what
the reference's NoiseLearner / Transmission / Spectrogram / DataController objects made of them. Needs no reference code, so
it also runs where oracle/_ref/libref.so is absent."""
import os
import struct

import numpy as np

import oracle_lib as ol
from conftest import ROOT, load_b2s

b2s = load_b2s()
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_blocks_n512.npz"), allow_pickle=False)


def test_oracle_reproduces_what_the_reference_blocks_emitted():
    n, fs, frames, learn, bw, t0 = (int(x) for x in G["meta"])
    period = float(G["period_ms"][0])
    cfg = b2s.make_config(n, fs, learn_frames=learn, recording_bandwidth_hz=bw, min_time_ms=200, timeout_ms=300)
    assert learn == b2s.lib().b2s_learn_frames_from_ms(2000, period)  # the reference's NOISE_LEARNING_TIME in frames of this clock
    assert cfg.spectrogram_out_size == 256 and n // cfg.spectrogram_out_size == 2  # decimating spectrogram (spectrogram.cpp:50-58)
    orc = ol.OracleChain(cfg)
    r = orc.push(G["psd"], frames, t0, period, dense=("noise_sub_db",), psd_rows=True)
    # NoiseLearner::work: every row bit for bit (the learning rows are -100, noise_learner.cpp:45-51)
    assert np.array_equal(r.noise_sub_db.view(np.uint32), G["noise_sub"].view(np.uint32))
    # Transmission::work: the list handed to TransmissionNotification::notify, frame by frame
    for k in range(frames):
        want = [(int(G["tx"][k, i, 0]), int(G["tx"][k, i, 1])) for i in range(int(G["tx_count"][k]))]
        assert [(f, fl) for f, fl, _, _ in r.frame_tx[k]] == want, f"frame {k}"
    assert int(G["tx_count"].sum()) > 100 and int(G["tx"][:, :, 1].sum()) > 10  # starts, flushes, stops are all in there
    # Spectrogram::work + DataController::pushSpectrogram: payloads byte for byte (first row: header only — the reference
    # leaves Container::m_counter uninitialised until the first send, spectrogram.cpp:9)
    times, centers, rows = orc.get_spectrogram(cap=16)
    payloads = [bytes(p) for p in G["payloads"]]
    assert len(payloads) == len(times) >= 4
    for i, p in enumerate(payloads):
        t_ms, start, stop, step, size = struct.unpack("<QiiiI", p[:24])
        assert (t_ms, start, stop, step, size) == (int(times[i]), cfg.center_hz - fs // 2, cfg.center_hz + fs // 2, fs // 256, 256)
        if i >= 1:
            assert p == b2s.pack_spectrogram_message(int(times[i]), int(centers[i]), fs, rows[i])
