"""Behavioural known-answer tests for the restated block chain (oracle/scan_oracle.cpp) — the area the reference's
tests leave unpinned (SURVEY.md §8c "New KATs the build must add") — plus an independent pure-Python restatement of
NoiseLearner/Transmission/Signal/Spectrogram (written from the reference sources, not from the C++ oracle) that must
agree with the C++ oracle frame by frame. CPU only."""
import math

import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_b2s

b2s = load_b2s()
import importlib.util, os, sys  # noqa: E402

_spec = importlib.util.spec_from_file_location("synth", os.path.join(os.path.dirname(b2s.__file__), "synth.py"))
synth = importlib.util.module_from_spec(_spec)
sys.modules["synth"] = synth
_spec.loader.exec_module(synth)


def scene(n=1024, fs=2_048_000, frames=400, learn=40, **kw):
    cfg = b2s.make_config(n, fs, learn_frames=learn, recording_bandwidth_hz=16 * fs // n, min_time_ms=20, timeout_ms=30, **kw)
    tones = synth.standard_scene(n, frames, learn)
    iq = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(0), quiet_frames=learn)
    return cfg, tones, iq, synth.frame_period_ms(n, fs)


# ------------------------------------------------------------------------------------------------------------
# independent Python restatement (reference file:line in comments)
# ------------------------------------------------------------------------------------------------------------
class PyChain:
    def __init__(self, cfg):
        self.c = cfg
        self.n = cfg.fft_size
        self.thr = np.full(self.n, -np.finfo(np.float32).max, np.float32)
        self.samples, self.ready = 0, False
        y = cfg.grouping_y
        self.rows = [np.zeros(self.n, np.float32) for _ in range(y)]
        self.sum = np.zeros(self.n, np.float32)
        self.frames = 0
        self.signals = {}  # key -> [first, last, power]
        self.step = float(cfg.sample_rate_hz) / cfg.fft_size
        self.spec_sum = None
        self.sent = []

    def shift(self, i):  # sdr_device.cpp:154
        return int(self.step * (i + 0.5)) - self.c.sample_rate_hz // 2

    def freq(self, i):  # sdr_device.cpp:153
        return self.c.center_hz + self.shift(i)

    @staticmethod
    def tuned(f, step):  # radio_utils.cpp:86-96 (C++ % truncates toward zero)
        rest = int(math.fmod(f, step))
        if f < 0:
            rest += step
        down = f - rest
        return down if rest < step - rest else down + step

    def max_index(self, d, i, g):  # collection_utils.h:9-14
        lo, hi = max(0, i - g // 2), min(self.n, i + g // 2 + 1)
        return lo + int(np.argmax(d[lo:hi]))

    def boxcar(self, x, g):  # utils.cpp:31-53
        a = g // 2
        out = np.zeros(self.n, np.float32)
        s, cnt = np.float32(0), 0
        for i in range(-a, self.n + a - 1):
            first, last = i - a - 1, i + a
            if 0 <= first < self.n:
                s = np.float32(s - x[first])
                cnt -= 1
            if 0 <= last < self.n:
                s = np.float32(s + x[last])
                cnt += 1
            if 0 <= i < self.n:
                out[i] = np.float32(s) / np.float32(cnt)
        return out

    def frame(self, psd, now):
        c, n = self.c, self.n
        # Spectrogram, spectrogram.cpp:29-75
        m = c.spectrogram_out_size
        if m > 0:
            d = n // m
            if self.spec_sum is None:
                self.spec_sum, self.spec_cnt, self.spec_last = np.zeros(m, np.float32), 0, now
            if d == 1:
                self.spec_sum = (self.spec_sum + psd).astype(np.float32)
            else:
                acc = np.zeros(m, np.float32)
                for j in range(d):
                    acc = (acc + psd[j::d]).astype(np.float32)
                self.spec_sum = (self.spec_sum + (acc / np.float32(d)).astype(np.float32)).astype(np.float32)
            self.spec_cnt += 1
            if self.spec_last + c.spectrogram_interval_ms < now:
                self.sent.append((now, np.trunc(self.spec_sum / np.float32(self.spec_cnt)).astype(np.int8)))
                self.spec_sum[:] = 0
                self.spec_cnt, self.spec_last = 0, now
        # NoiseLearner, noise_learner.cpp:36-67
        if not self.ready:
            self.thr = np.maximum(self.thr, psd)
            self.samples += 1
            if c.learn_frames <= self.samples:
                self.ready = True
            q = np.full(n, -100, np.float32)
        else:
            q = (psd - self.thr).astype(np.float32)
        # Averager::push, averager.cpp:14-25
        self.frames = min(self.frames + 1, c.grouping_y)
        old = self.rows.pop(0)
        self.sum = (self.sum - old).astype(np.float32)
        self.sum = (self.sum + q).astype(np.float32)
        self.rows.append(q.copy())
        avg = (self.sum / np.float32(c.grouping_y)).astype(np.float32) if self.frames >= c.grouping_y else np.full(n, -100, np.float32)
        box = self.boxcar(avg, c.grouping_x)
        # addSignals, transmission.cpp:88-111
        g = c.group_size_bins
        cand = [i for i in range(n) if c.start_level <= box[i] and c.range_lo_hz <= self.freq(i) <= c.range_hi_hz and not any(c.ignored_lo_hz[r] <= self.freq(i) <= c.ignored_hi_hz[r] for r in range(c.n_ignored))]
        cand.sort(key=lambda i: (-box[i], i))
        sub = g // 2 if g % 2 == 0 else g // 2 + 1
        for i in cand:
            if not any(i - sub <= k <= i + sub for k in self.signals):
                votes = []
                for row in self.rows[len(self.rows) // 2 :]:
                    b = self.max_index(row, i, g)
                    if c.start_level <= row[b]:
                        votes.append(b)
                if votes:
                    vals, counts = np.unique(votes, return_counts=True)
                    tied = vals[counts == counts.max()]
                    key = int(tied[len(tied) // 2])
                else:
                    key = i
                self.signals.setdefault(key, [now, now, np.float32(0)])
        for k, s in self.signals.items():  # updateSignals + Signal::newData
            b = self.max_index(box, k, g)
            s[2] = box[b]
            if c.stop_level <= box[b]:
                s[1] = now
        for k in [k for k, s in self.signals.items() if s[1] + c.timeout_ms <= now or s[0] + c.max_time_ms <= now]:
            del self.signals[k]
        keys = sorted(self.signals, key=lambda k: (-self.signals[k][2], k))
        return [(self.tuned(self.shift(k), c.tuning_step_hz), int(self.signals[k][1] == now and self.signals[k][0] + c.min_time_ms <= now), k) for k in keys], q, avg, box


def test_learning_frames_output_no_data_and_threshold_is_row_max():
    cfg, tones, iq, period = scene(frames=60, learn=40)
    ch = ol.OracleChain(cfg)
    r = ch.push(iq, 60, 1000, period)
    assert np.all(r.noise_sub_db[:40] == -100.0)  # including the frame that completes learning (noise_learner.cpp:45-51)
    assert np.all(r.peak_index[:40] == -1)
    thr, samples, ready = ch.get_noise()
    assert ready and samples == 40
    assert np.array_equal(thr, r.psd_db[:40].max(axis=0))
    assert np.array_equal(r.noise_sub_db[40:], r.psd_db[40:] - thr)
    assert np.array_equal(r.peak_index[40:], np.argmax(r.psd_db[40:], axis=1))


def test_averager_warm_up_blocks_detection():
    cfg, tones, iq, period = scene(frames=80, learn=5)
    ch = ol.OracleChain(cfg)
    r = ch.push(iq, 80, 0, period)
    y = cfg.grouping_y
    assert np.all(r.avg_db[: y - 1] == -100.0) and np.all(r.tx_count[: y - 1] == 0)
    assert not np.any(r.avg_db[y - 1 :] == -100.0)


def test_cpp_oracle_matches_independent_python_restatement():
    cfg, tones, iq, period = scene(n=256, frames=260, learn=30)
    cfg.spectrogram_interval_ms = 7
    cfg.spectrogram_out_size = 64
    ch = ol.OracleChain(cfg)
    r = ch.push(iq, 260, 5000, period)
    py = PyChain(cfg)
    seen_tx = 0
    for k in range(260):
        now = 5000 + int(math.floor(k * period + 0.5))
        tx, q, avg, box = py.frame(r.psd_db[k], now)
        assert np.array_equal(q, r.noise_sub_db[k]), k
        assert np.array_equal(avg, r.avg_db[k]), k
        assert np.array_equal(box, r.box_db[k]), k
        assert [(f, fl, key) for f, fl, key, _ in r.frame_tx[k]] == tx, k
        seen_tx += len(tx)
    assert seen_tx > 100  # the scene really exercises the tracker
    t_c, _, rows_c = ch.get_spectrogram()
    assert len(py.sent) == len(t_c) and len(t_c) >= 1
    for (t, row), tc, rc in zip(py.sent, t_c, rows_c):
        assert t == tc and np.array_equal(row, rc)


def test_detection_lifecycle_on_standard_scene():
    n, fs, frames, learn = 1024, 2_048_000, 400, 40
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    ch = ol.OracleChain(cfg)
    r = ch.push(iq, frames, 0, period)
    step = fs / n
    for t in tones:
        expect_shift = t.bin_offset * step
        on0, on1 = t.on_frames[0]
        mid = (on0 + on1) // 2
        freqs = [f for f, _, _, _ in r.frame_tx[mid]]
        assert any(abs(f - expect_shift) <= 10 * step + cfg.tuning_step_hz for f in freqs), (t.bin_offset, freqs)
        # a flush is requested once the signal has lasted min_time (Signal::needFlush, signal.cpp:32)
        if (on1 - on0) * period > 3 * cfg.min_time_ms:
            assert any(fl for k in range(on0, on1) for f, fl, _, _ in r.frame_tx[k] if abs(f - expect_shift) <= 10 * step + cfg.tuning_step_hz)
    # everything has timed out some frames after the last key-off (Signal::isTimeout, signal.cpp:30)
    last_off = max(b for t in tones for _, b in t.on_frames)
    gone = last_off + cfg.grouping_y + int(math.ceil(cfg.timeout_ms / period)) + 2
    if gone < frames:
        assert r.tx_count[gone:].max() == 0
    # nothing before the averager is warm
    assert r.tx_count[: learn + cfg.grouping_y - 1].max() == 0


def test_hop_reset_keeps_noise_and_clears_averager():
    cfg, tones, iq, period = scene(frames=200, learn=40)
    ch = ol.OracleChain(cfg)
    ch.push(iq, 150, 0, period, dense=())
    thr0, _, ready0 = ch.get_noise()
    ch.reset()  # Transmission::resetBuffers, transmission.cpp:42-55
    s, a, ring, f = ch.get_averager()
    assert f == 0 and not s.any() and not ring.any() and np.all(a == -100.0)
    thr1, _, ready1 = ch.get_noise()
    assert ready0 and ready1 and np.array_equal(thr0, thr1)
    n = cfg.fft_size
    r = ch.push(iq[150 * n * 2 :], 50, 1000, period)
    assert np.all(r.avg_db[: cfg.grouping_y - 1] == -100.0)  # warm-up again
    assert not np.any(r.noise_sub_db == -100.0)  # no re-learning
    # a different centre frequency learns its own threshold (map keyed by centre, noise_learner.cpp:41-42)
    ch.set_center(cfg.center_hz + 2_000_000, cfg.range_lo_hz + 2_000_000, cfg.range_hi_hz + 2_000_000)
    r2 = ch.push(iq, 45, 2000, period)
    assert np.all(r2.noise_sub_db[:40] == -100.0) and not np.any(r2.noise_sub_db[40:] == -100.0)
    ch.set_center(cfg.center_hz, cfg.range_lo_hz, cfg.range_hi_hz)
    r3 = ch.push(iq, 3, 3000, period)
    assert not np.any(r3.noise_sub_db == -100.0)  # the first centre's threshold was kept


def test_spectrogram_decimates_then_accumulates_and_truncates():
    n, fs = 1024, 2_048_000
    cfg = b2s.make_config(n, fs, learn_frames=5, spectrogram_out_size=256)
    rng = np.random.default_rng(1)
    frames = 2200  # 0.5 ms per frame -> one row per ~1000 ms
    iq = rng.integers(-60, 60, frames * n * 2).astype(np.int8)
    ch = ol.OracleChain(cfg)
    period = synth.frame_period_ms(n, fs)
    r = ch.push(iq, frames, 0, period, dense=("psd_db",))
    times, centers, rows = ch.get_spectrogram()
    assert len(times) == 1 and centers[0] == cfg.center_hz
    k_send = int(np.searchsorted(np.floor(np.arange(frames) * period + 0.5), 1000, side="right"))  # first now > 1000
    assert times[0] == int(math.floor(k_send * period + 0.5))
    acc = np.zeros(256, np.float32)
    for k in range(k_send + 1):
        s = np.zeros(256, np.float32)
        for j in range(4):
            s = (s + r.psd_db[k, j::4]).astype(np.float32)
        acc = (acc + (s / np.float32(4)).astype(np.float32)).astype(np.float32)
    expect = np.trunc(acc / np.float32(k_send + 1)).astype(np.int8)  # float -> int8 truncates toward zero
    assert np.array_equal(rows[0], expect)
    assert np.all(rows[0] < 0)  # dB/Hz of 8-bit noise is negative: truncation toward zero != floor, and it shows


def test_range_and_ignored_frequencies_gate_new_signals_only():
    n, fs, frames, learn = 1024, 2_048_000, 300, 40
    base, tones, iq, period = scene(n, fs, frames, learn)
    step = fs / n
    f_tone0 = base.center_hz + int(tones[0].bin_offset * step)
    cfg = b2s.make_config(n, fs, learn_frames=learn, recording_bandwidth_hz=16 * fs // n, min_time_ms=20, timeout_ms=30,
                          ignored=[(f_tone0 - 40 * int(step), f_tone0 + 40 * int(step))])
    r = ol.OracleChain(cfg).push(iq, frames, 0, period, dense=())
    shifts = {f for k in range(frames) for f, _, _, _ in r.frame_tx[k]}
    assert not any(abs(f - tones[0].bin_offset * step) < 30 * step for f in shifts)
    assert any(abs(f - tones[1].bin_offset * step) < 4 * step + 2500 for f in shifts)
