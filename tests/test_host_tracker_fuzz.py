"""Randomised scenes for the band's host tracker (tracker.h) against the oracle: PSD rows with carriers that start, stop,
drift, overlap and sit near the band edges, ignored ranges and scan-range limits, odd and even group sizes. The oracle turns
the PSD rows into NoiseLearner / boxcar rows and per-frame lists; the tracker, fed with those dense rows through
b2s_host_transmission_*, must produce the same lists — with K2's watch data emulated and without, in one call and in chunks."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_b2s

b2s = load_b2s()


def _scene(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([256, 512]))
    fs = 2_048_000
    frames, learn = 260, 24
    period = 4.0
    bw = int(rng.choice([6, 9, 16, 25])) * fs // n  # group size in bins (odd and even)
    kw = {}
    center = 100_000_000
    if rng.random() < 0.5:  # two ignored ranges somewhere in the band
        lo = center - fs // 2 + int(rng.integers(0, fs - 200_000))
        kw["ignored"] = [(lo, lo + int(rng.integers(20_000, 150_000)))]
    cfg = b2s.make_config(n, fs, center_hz=center, learn_frames=learn, recording_bandwidth_hz=bw, min_time_ms=int(rng.choice([0, 20, 60])),
                          timeout_ms=int(rng.choice([12, 40, 100])), **kw)
    if rng.random() < 0.4:  # scan range narrower than the band: only bins inside may START a signal
        cfg.range_lo_hz = center - int(rng.integers(100_000, fs // 2))
        cfg.range_hi_hz = center + int(rng.integers(100_000, fs // 2))
    psd = (-60.0 + 1.5 * rng.standard_normal((frames, n))).astype(np.float32)
    for _ in range(int(rng.integers(3, 9))):
        c = float(rng.integers(0, n))
        a, b = sorted(int(x) for x in rng.integers(learn, frames, 2))
        level = float(rng.uniform(12.0, 70.0))  # after the 21-bin boxcar some end up hovering around the start / stop levels
        width = float(rng.uniform(3.0, 14.0))
        drift = float(rng.uniform(-0.05, 0.05))
        for t in range(a, b):
            cc = c + drift * (t - a) + 2.0 * np.sin(0.7 * t)
            bins = np.arange(n)
            psd[t] += (level * np.exp(-0.5 * ((bins - cc) / width) ** 2)).astype(np.float32)
            if rng.random() < 0.03:
                psd[t] -= np.float32(level)  # a dropout frame
    return cfg, psd, frames, period


@pytest.mark.parametrize("seed", range(24))
def test_random_scene(seed):
    cfg, psd, frames, period = _scene(seed)
    r = ol.OracleChain(cfg).push(psd, frames, 0, period, dense=("noise_sub_db", "box_db"), psd_rows=True)
    want = [[(f, fl, k, np.float32(p)) for f, fl, k, p in fr] for fr in r.frame_tx]
    for use_watch in (False, True):
        h = b2s.HostTransmission(cfg)
        got = h.push(r.box_db, r.noise_sub_db, 0, period, use_watch=use_watch)
        got = [[(f, fl, k, np.float32(p)) for f, fl, k, p in fr] for fr in got]
        bad = [k for k in range(frames) if got[k] != want[k]]
        assert not bad, f"seed {seed} watch {use_watch}: first differing frame {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"
    # chunked (integer period: chunk-local stamps equal global ones), watch data refreshed per chunk like the band does
    h = b2s.HostTransmission(cfg)
    got = []
    for a in range(0, frames, 37):
        b = min(frames, a + 37)
        got += h.push(r.box_db[a:b], r.noise_sub_db[a:b], int(a * period), period, use_watch=True)
    got = [[(f, fl, k, np.float32(p)) for f, fl, k, p in fr] for fr in got]
    bad = [k for k in range(frames) if got[k] != want[k]]
    assert not bad, f"seed {seed} chunked: first differing frame {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"
    _RECORDS.append(sum(len(fr) for fr in want))


_RECORDS = []


def test_the_scenes_were_not_empty():
    """Runs after the parametrised cases (file order): most scenes must have produced transmissions, several concurrent ones."""
    assert len(_RECORDS) == 24 and sum(1 for x in _RECORDS if x > 50) >= 16, _RECORDS
