"""The band's host tracker (rtl-sdr-scanner-cpp_b200/csrc/tracker.h — addSignals / getBestIndex / updateSignals / clearSignals /
getSortedTransmissions of transmission.cpp:70-176, driven by sparse detection entries and K2's watch data) checked on the
CPU: b2s_host_transmission_* feeds it from dense rows, here the oracle's own boxcar / NoiseLearner rows, and the per-frame
lists must equal the oracle's (which test_oracle_vs_reference_blocks.py pins to the reference's compiled Transmission)."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_b2s
from test_oracle_chain import scene

b2s = load_b2s()


def _lists(frame_tx):
    return [[(f, fl, k, np.float32(p)) for f, fl, k, p in fr] for fr in frame_tx]


@pytest.mark.parametrize("use_watch", [False, True])
@pytest.mark.parametrize("chunks", [1, 7])
def test_host_tracker_reproduces_the_oracle_lists(use_watch, chunks):
    cfg, tones, iq, period = scene(n=1024, frames=400, learn=40)
    r = ol.OracleChain(cfg).push(iq, 400, 0, period, dense=("noise_sub_db", "box_db"))
    want = _lists(r.frame_tx)
    assert sum(len(x) for x in want) > 300 and any(fl for fr in want for _, fl, _, _ in fr)
    h = b2s.HostTransmission(cfg)
    got = []
    edges = np.linspace(0, 400, chunks + 1).astype(int)
    for a, b in zip(edges[:-1], edges[1:]):
        t0 = int(np.floor(a * period + 0.5))
        # frame k of the chunk must carry the stamp of frame a + k of the whole run: push chunk by chunk with the exact stamps
        part = h.push(r.box_db[a:b], r.noise_sub_db[a:b], 0, period, use_watch=use_watch) if chunks == 1 else _push_at(h, r, a, b, period, use_watch)
        got += part
    assert _lists(got) == want


def _push_at(h, r, a, b, period, use_watch):
    """Chunk [a, b) one frame at a time with the stamp floor(k * period + 0.5) of the frame's GLOBAL index (a chunk-level t0
    plus a chunk-local k would round differently for a non-integer period)."""
    out = []
    for k in range(a, b):
        out += h.push(r.box_db[k : k + 1], r.noise_sub_db[k : k + 1], int(np.floor(k * period + 0.5)), period, use_watch=use_watch)
    return out


def test_reset_drops_signals_and_ring():
    cfg, tones, iq, period = scene(n=1024, frames=300, learn=40)
    oc = ol.OracleChain(cfg)
    r = oc.push(iq, 300, 0, period, dense=("noise_sub_db", "box_db"))
    h = b2s.HostTransmission(cfg)
    first = h.push(r.box_db[:200], r.noise_sub_db[:200], 0, period)
    assert first[199], "a carrier is up at frame 199 of the standard scene"
    h.reset()
    quiet = np.full((5, cfg.fft_size), -3.0, dtype=np.float32)
    assert h.push(quiet, quiet, 1000, period) == [[], [], [], [], []]
