"""GPU parity tests (run with -m gpu on a B200): every call goes through the C-ABI (libb2s.so) and is compared with the
CPU oracle on the same seeded inputs.

Bars (DESIGN.md "Parity criterion"):
  * raw PSD rows: |dB error| <= 2e-3 dB; linear power bins: floored criterion |p-p_ref| <= 1e-5*max(p_ref, median)
    pass fraction >= 99.5 %, worst <= 1e-4, L2-relative <= 1e-6 (an fp32 FFT — FFTW3f included — cannot do better
    against the exact result; the strict per-bin fraction is printed).
  * Averager state (m_sum, ring, m_average, m_frames): BIT-EXACT when both sides see identical rows.
  * detection: per-frame FrequencyFlush lists, signal keys and peak indices EQUAL on margin-safe scenes.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT, load_b2s
from test_oracle_chain import scene, synth
from test_oracle_kats import run_averager_fixture, _run_averager_script, G
from test_oracle_spectrum import power_parity_stats

b2s = load_b2s()
pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------------------
# K1: unpack + window + FFT + PSD
# ------------------------------------------------------------------------------------------------------------
def _noise_tones(n, frames, seed, stride=None):
    tones = [synth.Tone(1234.1 * n / 16384), synth.Tone(-3000.1 * n / 16384), synth.Tone(77.1 * n / 16384, amplitude=25), synth.Tone(6000.1 * n / 16384, fm_dev_bins=5)]
    return synth.make_iq_int8(n, frames, tones, seed=seed, stride=stride)


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144])
def test_psd_rows_match_oracle(engine, n):
    """N <= 2048: k_spectrum; 4096..16384: k_spectrum3; 32768..262144: k_spectrum3's split mode (S = N/16384 residue classes)."""
    frames = 6 if n <= 16384 else 3
    fs = 20_000_000 if n >= 8192 else 2_048_000
    cfg = b2s.make_config(n, fs)
    iq = _noise_tones(n, frames, seed=n)
    psd, lin = engine.psd(cfg, iq, frames, want_linear=True)
    ref = np.empty_like(psd)
    ref_lin = np.empty_like(lin)
    for k in range(frames):
        ref[k], ref_lin[k] = ol.oracle_psd_frame(cfg, iq[k * 2 * n : (k + 1) * 2 * n], want_linear=True)
    ol.assert_db_rows_close(psd, ref, f"N={n}")
    st = power_parity_stats(lin, ref_lin)
    print(f"\nN={n}: floored pass {st['pass_frac']:.5f} worst {st['worst']:.2e} strict pass {st['strict_frac']:.4f} L2rel {st['l2_rel']:.2e}")
    assert st["pass_frac"] >= 0.995 and st["worst"] <= ol.worst_tolerance(n) and st["l2_rel"] <= 1e-6, st
    assert np.array_equal(np.argmax(psd, axis=1), np.argmax(ref, axis=1))


@pytest.mark.parametrize("n", [4096, 32768, 262144])
def test_psd_peaks_through_the_band(engine, n):
    """peak_index / peak_value (first maximum of the raw row, noise_learner.cpp:53-59): with the split mode the S residue classes of
    a frame are reduced through a packed atomic maximum."""
    fs, frames = 20_000_000, 5
    cfg = b2s.make_config(n, fs, learn_frames=2, spectrogram_out_size=0, max_frames_per_push=8)
    iq = _noise_tones(n, frames, seed=n + 1)
    got = b2s.Band(engine, cfg).push(iq, frames, 0, 1.0, per_frame=True, dense=("psd_db",))
    assert np.array_equal(got.peak_index, np.argmax(got.psd_db, axis=1))
    assert np.array_equal(got.peak_value, got.psd_db.max(axis=1))


def test_psd_known_answers_on_gpu(engine):
    n, fs = 4096, 2_048_000
    cfg = b2s.make_config(n, fs, iq_scale=1.0)
    w = ol.hamming(n).astype(np.float64)
    iq = np.zeros((3, 2 * n), np.int8)
    iq[0, 0] = 100  # impulse -> flat
    iq[1, 0::2] = 50  # DC -> main lobe at N/2
    k = np.arange(n)
    z = 60 * np.exp(2j * np.pi * k / 4)  # +fs/4 -> 3N/4
    iq[2, 0::2], iq[2, 1::2] = np.rint(z.real), np.rint(z.imag)
    p = engine.psd(cfg, iq.reshape(-1), 3)
    assert np.max(np.abs(p[0] - 10 * np.log10((100 * w[0]) ** 2 / fs))) < 1e-3
    assert int(np.argmax(p[1])) == n // 2 and abs(p[1, n // 2] - 10 * np.log10((50 * w.sum()) ** 2 / fs)) < 1e-3
    assert int(np.argmax(p[2])) == 3 * n // 4


@pytest.mark.parametrize("n", [2048, 8192, 16384, 32768, 131072])
def test_psd_input_variants(engine, n):
    """CF32 input, decimated frames (stride r*N) and a stride that defeats the 16-byte TMA path all give the same rows — for
    k_spectrum (2048), k_spectrum3's three load modes (8192, 16384) and the split mode's pre-pass (32768, 131072)."""
    fs, frames = 2_048_000, 5 if n <= 16384 else 3
    iq = _noise_tones(n, frames, seed=3)
    base = engine.psd(b2s.make_config(n, fs), iq, frames)
    ref = np.stack([ol.oracle_psd_frame(b2s.make_config(n, fs), iq[k * 2 * n : (k + 1) * 2 * n]) for k in range(frames)])
    ol.assert_db_rows_close(base, ref, f"cs8 N={n}")
    f32 = (iq.astype(np.float32) * np.float32(1 / 127.0)).astype(np.float32)
    cfg_f = b2s.make_config(n, fs, iq_format=b2s.IQ_CF32)
    cf = engine.psd(cfg_f, f32, frames)
    ol.assert_db_rows_close(cf, base, "cf32 vs cs8")  # only the unpack rounding differs (scale folded into the window for CS8)
    ref_f = np.stack([ol.oracle_psd_frame(cfg_f, f32[k * 2 * n : (k + 1) * 2 * n]) for k in range(frames)])
    ol.assert_db_rows_close(cf, ref_f, f"cf32 N={n} vs oracle")
    # decimator: keep the first N samples of every 3N (decimator.h:16-22)
    wide = np.zeros((frames, 3 * n * 2), np.int8)
    wide[:, : 2 * n] = iq.reshape(frames, 2 * n)
    wide[:, 2 * n :] = 77
    dec = engine.psd(b2s.make_config(n, fs, decimator=3), wide.reshape(-1), frames)
    assert np.array_equal(dec, base)
    cfg = b2s.make_config(n, fs)
    cfg.frame_stride_samples = n + 3  # 2N+6 bytes: not a multiple of 16 -> direct-load kernel
    odd = np.zeros((frames, (n + 3) * 2), np.int8)
    odd[:, : 2 * n] = iq.reshape(frames, 2 * n)
    assert np.array_equal(engine.psd(cfg, odd.reshape(-1), frames), base)


def test_parseval_property_at_full_size(engine):
    """Size-independent property at BASELINE config-2 geometry (N=16384): sum_k |X_k|^2 = N * sum_n |x_n w_n|^2."""
    n, fs, frames = 16384, 20_000_000, 64
    cfg = b2s.make_config(n, fs)
    iq = _noise_tones(n, frames, seed=99)
    _, lin = engine.psd(cfg, iq, frames, want_linear=True)
    w = ol.hamming(n).astype(np.float64)
    x = iq.reshape(frames, n, 2).astype(np.float64) * (1.0 / 127.0)
    energy = np.sum((x[..., 0] ** 2 + x[..., 1] ** 2) * w**2, axis=1) * n / fs
    assert np.max(np.abs(lin.astype(np.float64).sum(axis=1) / energy - 1.0)) < 2e-6


# ------------------------------------------------------------------------------------------------------------
# Averager / average() operators — the reference's own unit tests, run against device-backed operators
# ------------------------------------------------------------------------------------------------------------
def test_device_averager_reference_kats(engine):
    g = G["averager"]
    _run_averager_script(b2s.Averager(engine, g["size"], g["group"]))
    f = G["averager_fixture"]
    size, group = f["size"], f["group"]
    run_averager_fixture(b2s.Averager(engine, size, group), size, group, f["simple"])
    rows = list(f["big"]["prefix"]) + [[i * 11 + j * 7 for j in range(size)] for i in range(*f["big"]["ramp_i"])]
    run_averager_fixture(b2s.Averager(engine, size, group), size, group, rows)


@pytest.mark.parametrize("size,group,batch", [(5, 3, 1), (4096, 21, 1), (4096, 21, 7), (16384, 21, 50), (1000, 4, 3)])
def test_device_averager_bitwise_vs_oracle(engine, size, group, batch):
    rng = np.random.default_rng(size + group + batch)
    dev, cpu = b2s.Averager(engine, size, group), ol.CpuAverager(size, group, "orc")
    for step in range(6):
        if step == 4:
            dev.reset(), cpu.reset()
        rows = (rng.standard_normal((batch, size)) * 30 - 20).astype(np.float32)
        dev.push(rows if batch > 1 else rows[0])
        for r in rows:
            cpu.push(r)
        s_d, f_d = dev.sum()
        s_c, f_c = cpu.sum()
        assert f_d == f_c and s_d.tobytes() == s_c.tobytes()
        assert dev.average().tobytes() == cpu.average().tobytes()
        assert dev.data().tobytes() == cpu.data().tobytes()


def _window_mean64(row, group):
    n, a = len(row), group // 2
    if a == 0:
        out = row.astype(np.float64).copy()
        out[-1] = 0.0  # reference quirk: groupSize 1 never writes the last element (utils.cpp:38)
        return out
    c = np.concatenate([[0.0], np.cumsum(row.astype(np.float64))])
    lo, hi = np.maximum(0, np.arange(n) - a), np.minimum(n - 1, np.arange(n) + a)
    return (c[hi + 1] - c[lo]) / (hi - lo + 1)


def test_device_average_operator(engine):
    g = G["average"]
    x = np.array(g["input"], np.float32)
    for exact in (False, True):
        np.testing.assert_allclose(engine.average(x, g["group"], exact=exact), np.array(g["expect"], np.float32), rtol=4e-7)
    rng = np.random.default_rng(4)
    for size, group in [(9, 5), (100, 21), (4096, 21), (16384, 21), (50, 1), (10, 21), (7, 4)]:
        x = (rng.standard_normal((3, size)) * 20 - 7).astype(np.float32)
        ref = np.stack([ol.cpu_average(r, group) for r in x])
        assert engine.average(x, group, exact=True).tobytes() == ref.tobytes()  # serial form: bit-exact
        fused = engine.average(x, group, exact=False)
        exact_mean = np.stack([_window_mean64(r, group) for r in x])
        # the reference's single running sum drifts by rounding (up to ~2e-4 at the far end of a 16384-bin row);
        # the fused form restarts per bin, so it stays within a few ulp of the exact window mean
        assert np.max(np.abs(fused - ref)) <= 1e-3
        assert np.max(np.abs(fused - exact_mean)) <= 2e-5
    x = np.full((1, 4096), -100.0, np.float32)
    assert np.max(np.abs(engine.average(x, 21) - ol.cpu_average(x[0], 21))) <= 1e-3


# ------------------------------------------------------------------------------------------------------------
# the whole band chain
# ------------------------------------------------------------------------------------------------------------
def _assert_peaks_equal(got_idx, ref_idx, ref_psd):
    """argmax of the raw PSD row: equal, except where the oracle's own two best bins tie within fp32 rounding."""
    for k in np.nonzero(got_idx != ref_idx)[0]:
        assert abs(float(ref_psd[k, got_idx[k]]) - float(ref_psd[k, ref_idx[k]])) <= 1e-4, (k, got_idx[k], ref_idx[k])
    assert np.mean(got_idx == ref_idx) >= 0.98


def _tx(frames):
    return [[(f, fl, k) for f, fl, k, _ in fr] for fr in frames]


DENSE = ("psd_db", "noise_sub_db", "avg_db", "box_db")


def _run_both(engine, cfg, iq, frames, period, t0=0, splits=None):
    band = b2s.Band(engine, cfg)
    ref = ol.OracleChain(cfg).push(iq, frames, t0, period)
    if splits is None:
        got = band.push(iq, frames, t0, period, per_frame=True, dense=DENSE)
        return band, got, ref
    raise NotImplementedError


@pytest.mark.parametrize("n,fs,frames,learn", [(1024, 2_048_000, 400, 40), (4096, 2_048_000, 260, 30), (256, 2_048_000, 300, 30),
                                               (8192, 2_048_000, 400, 40), (32768, 20_000_000, 400, 40)])
def test_band_matches_oracle_end_to_end(engine, n, fs, frames, learn):
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    band, got, ref = _run_both(engine, cfg, iq, frames, period, t0=1000)
    ol.assert_db_rows_close(got.psd_db, ref.psd_db, "psd")
    assert np.array_equal(got.noise_sub_db[:learn], ref.noise_sub_db[:learn])  # -100 rows
    main = ref.psd_db[learn:] >= np.median(ref.psd_db[learn:], axis=1, keepdims=True) - 10.0
    assert np.max(np.abs(got.noise_sub_db[learn:] - ref.noise_sub_db[learn:])[main]) <= 4e-3
    assert np.max(np.abs(got.avg_db - ref.avg_db)) <= 4e-3
    assert np.max(np.abs(got.box_db - ref.box_db)) <= 4e-3
    _assert_peaks_equal(got.peak_index[learn:], ref.peak_index[learn:], ref.psd_db[learn:])
    assert _tx(got.frame_tx) == _tx(ref.frame_tx)
    assert sum(len(x) for x in ref.frame_tx) > 50
    for a, b in zip(got.frame_tx, ref.frame_tx):
        for (_, _, _, pa), (_, _, _, pb) in zip(a, b):
            assert abs(pa - pb) <= 4e-3
    # noise threshold and spectrogram against the oracle
    thr_g, samples_g, ready_g = band.get_noise()
    o = ol.OracleChain(cfg)
    o.push(iq, frames, 1000, period, dense=())
    thr_o, samples_o, ready_o = o.get_noise()
    assert (samples_g, ready_g) == (samples_o, ready_o) and np.max(np.abs(thr_g - thr_o)) <= 2e-3


def test_generic_grouping_parameters(engine):
    """GROUPING_X / GROUPING_Y other than the reference's 21/21 run the generic (runtime-parameter) kernel path."""
    n, fs, frames, learn = 1024, 2_048_000, 300, 40
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    for gx, gy in ((9, 7), (1, 1), (33, 40)):
        cfg.grouping_x, cfg.grouping_y = gx, gy
        band = b2s.Band(engine, cfg)
        got = band.push(iq, frames, 0, period, per_frame=True, dense=DENSE)
        ref = ol.OracleChain(cfg).push(iq, frames, 0, period)
        assert np.max(np.abs(got.avg_db - ref.avg_db)) <= 4e-3 and np.max(np.abs(got.box_db - ref.box_db)) <= 4e-3, (gx, gy)
        assert _tx(got.frame_tx) == _tx(ref.frame_tx), (gx, gy)
        cpu = ol.CpuAverager(n, gy, "orc")
        for k in range(frames):
            cpu.push(got.noise_sub_db[k])
        s, a, ring, f = band.get_averager()
        assert s.tobytes() == cpu.sum()[0].tobytes() and ring.tobytes() == cpu.data().tobytes() and a.tobytes() == cpu.average().tobytes()


def test_averager_state_is_bit_exact_on_identical_rows(engine):
    """Feed the oracle's Averager the GPU's own noise-subtracted rows: m_sum, ring, m_average, m_frames and every
    per-frame average row must then be bit-identical (operator-level parity inside the fused chain)."""
    n, fs, frames, learn = 4096, 2_048_000, 130, 20
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    band = b2s.Band(engine, cfg)
    got = band.push(iq, frames, 0, period, dense=DENSE)
    thr, _, _ = band.get_noise()
    q_expect = np.where(np.arange(frames)[:, None] < learn, np.float32(-100), got.psd_db - thr).astype(np.float32)
    assert got.noise_sub_db.tobytes() == q_expect.tobytes()  # NoiseLearner arithmetic, exact
    assert np.array_equal(thr, got.psd_db[:learn].max(axis=0))
    cpu = ol.CpuAverager(n, cfg.grouping_y, "orc")
    for k in range(frames):
        cpu.push(got.noise_sub_db[k])
        assert cpu.average().tobytes() == got.avg_db[k].tobytes(), k
    s, a, ring, f = band.get_averager()
    cs, cf = cpu.sum()
    assert f == cf and s.tobytes() == cs.tobytes() and a.tobytes() == cpu.average().tobytes() and ring.tobytes() == cpu.data().tobytes()
    # boxcar: serial reference form on the same rows differs from the fused form by rounding only
    ref_box = np.stack([ol.cpu_average(r, cfg.grouping_x) for r in got.avg_db])
    assert np.max(np.abs(got.box_db - ref_box)) <= 1e-3


def test_fast_path_equals_dense_path_bitwise(engine):
    """The specialised steady-state tiles (no dense rows requested) and the generic tiles (dense rows requested) must
    leave bit-identical Averager state, detections and spectrogram rows."""
    n, fs, frames, learn = 4096, 2_048_000, 400, 30
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    cfg.spectrogram_interval_ms = 50
    fast, slow = b2s.Band(engine, cfg), b2s.Band(engine, cfg)
    gf = fast.push(iq, frames, 0, period, per_frame=True)
    gs = slow.push(iq, frames, 0, period, per_frame=True, dense=DENSE)
    assert _tx(gf.frame_tx) == _tx(gs.frame_tx) and gf.n_detect_entries == gs.n_detect_entries > 0
    for a, b in zip(fast.get_averager(), slow.get_averager()):
        assert np.array_equal(a, b)
    tf_, _, rf = fast.get_spectrogram()
    ts_, _, rs = slow.get_spectrogram()
    assert np.array_equal(tf_, ts_) and np.array_equal(rf, rs) and len(tf_) >= 2


def test_async_mode_equals_sync_mode(engine):
    """B2S_FLAG_ASYNC (bookkeeping of push k on a worker thread while push k+1 is in the kernels) must leave the same
    state and deliver the same mailbox, signals and spectrogram rows as the synchronous mode, for host and device input."""
    import torch

    n, fs, frames, learn = 1024, 1_024_000, 600, 40
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    cfg.spectrogram_interval_ms = 40
    cfg.max_frames_per_push = 64
    sync = b2s.Band(engine, cfg)
    acfg = b2s.BandConfig.from_buffer_copy(cfg)
    acfg.flags |= b2s.FLAG_ASYNC
    asyn = b2s.Band(engine, acfg)
    dcfg = b2s.BandConfig.from_buffer_copy(acfg)
    dcfg.flags |= b2s.FLAG_IQ_ON_DEVICE
    adev = b2s.Band(engine, dcfg)
    iq_dev = torch.from_numpy(iq).cuda()
    sizes, k, i = [50, 64, 7, 120, 33, 64, 200], 0, 0
    mail_sync = None
    while k < frames:
        m = min(sizes[i % len(sizes)], frames - k)
        i += 1
        mail_sync = sync.push(iq[k * 2 * n :], m, 500 + k, period).transmissions
        asyn.push_raw(iq[k * 2 * n :].ctypes.data, m, 500 + k, period)
        adev.push_raw(iq_dev.data_ptr() + k * 2 * n, m, 500 + k, period)
        k += m
    for band in (asyn, adev):
        res = band.sync()
        mail = [(t.shift_hz, t.flush, t.key, t.power) for t in res.transmissions[: res.n_transmissions]]
        assert mail == mail_sync
        for a, b in zip(sync.get_averager(), band.get_averager()):
            assert np.array_equal(a, b)
        for a, b in zip(sync.get_signals(), band.get_signals()):
            assert np.array_equal(a, b)
        ts, _, rs = sync.get_spectrogram(consume=False)
        ta, _, ra = band.get_spectrogram()
        assert np.array_equal(ts, ta) and np.array_equal(rs, ra) and len(ts) >= 5
    assert len(sync.get_signals()[0]) >= 0 and mail_sync is not None
    with pytest.raises(b2s.B2SError):
        asyn.push(iq, 10, 5000, period)  # async mode takes no per-push result structure


def test_chunked_pushes_equal_one_push(engine):
    """State carried across b2s_band_push calls (ring hand-over, noise, tracker, spectrogram) — pushes of 1..97 frames,
    some shorter than the Averager depth — must reproduce the single-push result bit for bit."""
    n, fs, frames, learn = 1024, 1_024_000, 400, 40  # 1 ms per frame: the frame clock is additive at any split point
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    assert period == 1.0
    cfg.spectrogram_interval_ms = 40
    one = b2s.Band(engine, cfg)
    whole = one.push(iq, frames, 500, period, per_frame=True, dense=DENSE)
    t_w, c_w, rows_w = one.get_spectrogram()
    many = b2s.Band(engine, cfg)
    sizes, k, i = [1, 2, 5, 20, 21, 22, 97, 3, 60, 1, 80], 0, 0
    tx, parts = [], {d: [] for d in DENSE}
    while k < frames:
        m = min(sizes[i % len(sizes)], frames - k)
        i += 1
        r = many.push(iq[k * 2 * n :], m, 500 + k, period, per_frame=True, dense=DENSE)
        tx += r.frame_tx
        for d in DENSE:
            parts[d].append(getattr(r, d))
        k += m
    for d in DENSE:
        assert np.concatenate(parts[d]).tobytes() == getattr(whole, d).tobytes(), d
    assert _tx(tx) == _tx(whole.frame_tx) and sum(len(x) for x in tx) > 50
    for a, b in zip(one.get_averager(), many.get_averager()):
        assert np.array_equal(a, b)
    t_m, c_m, rows_m = many.get_spectrogram()
    assert np.array_equal(t_w, t_m) and np.array_equal(rows_w, rows_m) and len(t_w) >= 3


def test_spectrogram_rows_are_exact_on_identical_rows(engine):
    n, fs, frames = 1024, 2_048_000, 4100
    cfg = b2s.make_config(n, fs, learn_frames=5, spectrogram_out_size=256)
    rng = np.random.default_rng(1)
    iq = rng.integers(-60, 60, frames * n * 2).astype(np.int8)
    period = synth.frame_period_ms(n, fs)
    band = b2s.Band(engine, cfg)
    got = band.push(iq, frames, 0, period, dense=("psd_db",))
    times, centers, rows = band.get_spectrogram()
    assert len(times) == 2 and got.n_spectrogram_rows == 2
    now = np.floor(np.arange(frames) * period + 0.5).astype(np.int64)
    last, start = 0, 0
    for t_sent, row in zip(times, rows):
        k_send = int(np.argmax(now > last + 1000))
        assert t_sent == now[k_send]
        acc = np.zeros(256, np.float32)
        for k in range(start, k_send + 1):
            s = np.zeros(256, np.float32)
            for j in range(4):
                s = (s + got.psd_db[k, j::4]).astype(np.float32)
            acc = (acc + (s / np.float32(4)).astype(np.float32)).astype(np.float32)
        assert np.array_equal(row, np.trunc(acc / np.float32(k_send + 1 - start)).astype(np.int8))
        last, start = now[k_send], k_send + 1
    # and against the oracle end to end: off-by-one only where the mean sits on an integer boundary
    o = ol.OracleChain(cfg)
    o.push(iq, frames, 0, period, dense=())
    t_o, _, rows_o = o.get_spectrogram()
    assert np.array_equal(times, t_o) and np.max(np.abs(rows.astype(int) - rows_o.astype(int))) <= 1
    # d == 1 path
    cfg1 = b2s.make_config(n, fs, learn_frames=5, spectrogram_out_size=n)
    b1 = b2s.Band(engine, cfg1)
    g1 = b1.push(iq, 2100, 0, period, dense=("psd_db",))
    t1, _, r1 = b1.get_spectrogram()
    k_send = int(np.argmax(now > 1000))
    acc = np.zeros(n, np.float32)
    for k in range(k_send + 1):
        acc = (acc + g1.psd_db[k]).astype(np.float32)
    assert len(t1) == 1 and np.array_equal(r1[0], np.trunc(acc / np.float32(k_send + 1)).astype(np.int8))


def test_reset_and_retune_semantics(engine):
    cfg, tones, iq, period = scene(frames=200, learn=40)
    n = cfg.fft_size
    band, o = b2s.Band(engine, cfg), ol.OracleChain(cfg)
    band.push(iq, 150, 0, period)
    o.push(iq, 150, 0, period, dense=())
    thr0, _, ready0 = band.get_noise()
    band.reset(), o.reset()  # Transmission::resetBuffers: signals + averager cleared, noise kept
    s, a, ring, f = band.get_averager()
    assert f == 0 and not s.any() and not ring.any() and np.all(a == -100.0)
    assert len(band.get_signals()[0]) == 0
    thr1, _, ready1 = band.get_noise()
    assert ready0 and ready1 and np.array_equal(thr0, thr1)
    g = band.push(iq[150 * 2 * n :], 50, 1000, period, per_frame=True, dense=("noise_sub_db", "avg_db"))
    r = o.push(iq[150 * 2 * n :], 50, 1000, period)
    assert np.all(g.avg_db[: cfg.grouping_y - 1] == -100.0) and not np.any(g.noise_sub_db == -100.0)
    assert _tx(g.frame_tx) == _tx(r.frame_tx)
    # another centre frequency learns its own threshold; the first one is kept (noise_learner.cpp:41-42)
    band.set_center(cfg.center_hz + 2_000_000, cfg.range_lo_hz + 2_000_000, cfg.range_hi_hz + 2_000_000)
    g2 = band.push(iq, 45, 2000, period, dense=("noise_sub_db",))
    assert np.all(g2.noise_sub_db[:40] == -100.0) and not np.any(g2.noise_sub_db[40:] == -100.0)
    band.set_center(cfg.center_hz, cfg.range_lo_hz, cfg.range_hi_hz)
    g3 = band.push(iq, 3, 3000, period, dense=("noise_sub_db",))
    assert not np.any(g3.noise_sub_db == -100.0)
    assert np.array_equal(band.get_noise()[0], thr0)


def test_noise_learning_on_the_frame_clock_under_a_hop_schedule(engine):
    """noise_learning_ms (NoiseLearner's own wall-clock rule, noise_learner.cpp:11,23) through the engine: two centres visited alternately for
    500 ms each; the engine's learning frames, rows and per-frame lists equal the oracle's (which is pinned against the compiled reference
    NoiseLearner under the same schedule, tests/test_oracle_vs_reference_blocks.py)."""
    n, fs, dwell, hops = 1024, 2_048_000, 20, 14
    period = 25.0
    cfg = b2s.make_config(n, fs, learn_frames=10_000, noise_learning_ms=2000, recording_bandwidth_hz=16 * fs // n, min_time_ms=200, timeout_ms=300)
    frames = dwell * hops
    iq = synth.make_iq_int8(n, frames, synth.standard_scene(n, frames, 0), seed=synth.seed_for(11), quiet_frames=0)
    band, o = b2s.Band(engine, cfg), ol.OracleChain(cfg)
    centres = [cfg.center_hz, cfg.center_hz + 3_000_000]
    real_rows = 0
    for h in range(hops):
        c = centres[h % 2]
        band.reset(), o.reset()
        band.set_center(c, c - fs // 2, c + fs // 2), o.set_center(c, c - fs // 2, c + fs // 2)
        k0 = h * dwell
        t0 = int(np.floor(k0 * period + 0.5))
        part = iq[k0 * 2 * n : (k0 + dwell) * 2 * n]
        g = band.push(part, dwell, t0, period, per_frame=True, dense=("noise_sub_db",))
        r = o.push(part, dwell, t0, period, dense=("psd_db", "noise_sub_db"))
        learning_g = np.all(g.noise_sub_db == -100.0, axis=1)
        learning_r = np.all(r.noise_sub_db == -100.0, axis=1)
        assert np.array_equal(learning_g, learning_r), f"hop {h}: learning frames {np.nonzero(learning_g)[0]} vs {np.nonzero(learning_r)[0]}"
        if (~learning_r).any():
            main = r.psd_db[~learning_r] >= np.median(r.psd_db[~learning_r], axis=1, keepdims=True) - 10.0
            assert np.max(np.abs(g.noise_sub_db[~learning_r] - r.noise_sub_db[~learning_r])[main]) <= 4e-3
            real_rows += int((~learning_r).sum())
        assert _tx(g.frame_tx) == _tx(r.frame_tx), f"hop {h}"
        assert band.get_noise()[1:] == o.get_noise()[1:]  # (samples, ready) of the current centre
    assert real_rows == frames - 82  # 41 learning frames per centre, not 2 x 81 frames of dwell


def test_ignored_ranges_and_scan_range(engine):
    n, fs, frames, learn = 1024, 2_048_000, 300, 40
    base, tones, iq, period = scene(n, fs, frames, learn)
    step = fs / n
    f0 = base.center_hz + int(tones[0].bin_offset * step)
    cfg = b2s.make_config(n, fs, learn_frames=learn, recording_bandwidth_hz=16 * fs // n, min_time_ms=20, timeout_ms=30,
                          ignored=[(f0 - 40 * int(step), f0 + 40 * int(step))], range_hz=(base.center_hz - 900_000, base.center_hz + 1_000_000))
    band = b2s.Band(engine, cfg)
    got = band.push(iq, frames, 0, period, per_frame=True)
    ref = ol.OracleChain(cfg).push(iq, frames, 0, period, dense=())
    assert _tx(got.frame_tx) == _tx(ref.frame_tx) and sum(len(x) for x in ref.frame_tx) > 20


def test_full_size_geometry_detects_and_agrees_on_a_sample(engine):
    """BASELINE config-2 geometry (N=16384, fs=20 MS/s, reference levels/timeouts): 700 frames through the oracle
    (seconds on CPU) and through the engine; detections, keys and flush flags equal."""
    n, fs, frames, learn = 16384, 20_000_000, 700, 100
    cfg = b2s.make_config(n, fs, learn_frames=learn, min_time_ms=100, timeout_ms=120)
    tones = synth.standard_scene(n, frames, learn)
    iq = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(2), quiet_frames=learn)
    period = synth.frame_period_ms(n, fs)
    band = b2s.Band(engine, cfg)
    got = band.push(iq, frames, 0, period, per_frame=True, dense=("psd_db",))
    ref = ol.OracleChain(cfg).push(iq, frames, 0, period, dense=("psd_db",))
    ol.assert_db_rows_close(got.psd_db, ref.psd_db, "psd")
    assert _tx(got.frame_tx) == _tx(ref.frame_tx) and sum(len(x) for x in ref.frame_tx) > 200
    _assert_peaks_equal(got.peak_index[learn:], ref.peak_index[learn:], ref.psd_db[learn:])


# ------------------------------------------------------------------------------------------------------------
# error behaviour of the boundary
# ------------------------------------------------------------------------------------------------------------
def test_invalid_arguments_return_codes_not_crashes(engine):
    for bad in (dict(fft_size=1000), dict(fft_size=524288), dict(learn_frames=0), dict(tuning_step_hz=0)):
        kw = dict(fft_size=1024, learn_frames=10, tuning_step_hz=2500)
        kw.update(bad)
        cfg = b2s.make_config(kw.pop("fft_size"), 2_048_000, **kw)
        with pytest.raises(b2s.B2SError) as e:
            b2s.Band(engine, cfg)
        assert "b2s error -1" in str(e.value)
    cfg, tones, iq, period = scene(frames=120, learn=20)
    small = b2s.BandConfig.from_buffer_copy(cfg)
    small.detect_capacity = 8
    band = b2s.Band(engine, small)
    with pytest.raises(b2s.B2SError) as e:
        band.push(iq, 120, 0, period)
    assert "b2s error -4" in str(e.value)  # B2S_E_OVERFLOW, loud — not a silent truncation


# ------------------------------------------------------------------------------------------------------------
# the benchmark's own scene, N = 32768 chains, several engines, exact constant division
# ------------------------------------------------------------------------------------------------------------
def _bench_scene(n, fs, frames, learn, seed):
    import bench

    tones = bench.bench_tones(synth, n, frames, learn)
    return synth.make_iq_int8(n, frames, tones, seed=seed, quiet_frames=learn)


def _mailbox(res):
    return [(t.shift_hz, t.flush, t.key) for t in res.transmissions[: res.n_transmissions]]


def test_bench_scene_matches_oracle_async(engine):
    """bench.py's workload through the path bench.py times: N = 16384, fs = 20 MS/s, reference levels / time-outs, B2S_FLAG_ASYNC,
    device-resident IQ, two pushes of 2048 frames (the second starts with live signals). Mailbox after every push, live signal
    map and noise threshold against the oracle; Averager state bit-exact against an oracle Averager fed the engine's own rows."""
    import torch

    n, fs, frames, learn = 16384, 20_000_000, 2048, 100
    period = synth.frame_period_ms(n, fs)
    iq = _bench_scene(n, fs, frames, learn, synth.seed_for(2))
    cfg = b2s.make_config(n, fs, learn_frames=learn, max_frames_per_push=frames, flags=b2s.FLAG_ASYNC | b2s.FLAG_IQ_ON_DEVICE)
    band = b2s.Band(engine, cfg)
    ocfg = b2s.make_config(n, fs, learn_frames=learn)
    o = ol.OracleChain(ocfg)
    iq_dev = torch.from_numpy(iq).cuda()
    t0 = 0
    for rep in range(2):
        band.push_raw(iq_dev.data_ptr(), frames, t0, period)
        res = band.sync()
        ref = o.push(iq, frames, t0, period, dense=())
        assert _mailbox(res) == [(f, fl, k) for f, fl, k, _ in ref.frame_tx[-1]], rep
        assert res.n_detect_entries > 1000
        t0 += int(frames * period)
    keys_g, first_g, last_g, _ = band.get_signals()
    keys_o, first_o, last_o, _ = o.get_signals()
    assert np.array_equal(keys_g, keys_o) and np.array_equal(first_g, first_o) and np.array_equal(last_g, last_o)
    thr_g, _, ready_g = band.get_noise()
    thr_o, _, ready_o = o.get_noise()
    assert ready_g and ready_o and np.max(np.abs(thr_g - thr_o)) <= 2e-3
    # Averager state: bit-exact on identical rows (a synchronous twin with the same push history provides the rows of a third push)
    twin = b2s.Band(engine, b2s.make_config(n, fs, learn_frames=learn, max_frames_per_push=frames))
    twin.push(iq, frames, 0, period)
    twin.push(iq, frames, int(frames * period), period)
    got = twin.push(iq[: 64 * 2 * n], 64, t0, period, dense=("noise_sub_db",))
    band.push_raw(iq_dev.data_ptr(), 64, t0, period)
    band.sync()
    for a, b in zip(band.get_averager(), twin.get_averager()):
        assert np.array_equal(a, b)
    cpu = ol.CpuAverager(n, 21, "orc")
    for r in got.noise_sub_db[-21:]:
        cpu.push(r)
    s_, a_, ring, f = band.get_averager()
    assert ring.tobytes() == cpu.data().tobytes() and f == 21


def test_two_engines_in_one_process(engine):
    """Function attributes (opt-in shared memory) are per device: a second engine — on another GPU when the box has one, else on
    the same — must be able to launch the large-shared-memory kernels, and gives the same result."""
    import torch

    dev = 1 if torch.cuda.device_count() > 1 else 0
    other = b2s.Engine(dev)
    n, fs, frames, learn = 16384, 20_000_000, 160, 30
    cfg = b2s.make_config(n, fs, learn_frames=learn, min_time_ms=20, timeout_ms=30)
    iq = synth.make_iq_int8(n, frames, synth.standard_scene(n, frames, learn), seed=5, quiet_frames=learn)
    period = synth.frame_period_ms(n, fs)
    a = b2s.Band(engine, cfg).push(iq, frames, 0, period, per_frame=True, dense=("psd_db",))
    b = b2s.Band(other, cfg).push(iq, frames, 0, period, per_frame=True, dense=("psd_db",))
    assert np.array_equal(a.psd_db, b.psd_db) and _tx(a.frame_tx) == _tx(b.frame_tx) and sum(len(x) for x in a.frame_tx) > 20
    other.close()


def test_exact_constant_division_exhaustive(engine):
    """div_const<D> (detect.cuh, Markstein's 3-instruction sequence) against IEEE division for EVERY float with an exponent in the
    guarded range [2^-60, 2^60], both signs, D = 21 (GROUPING_X = GROUPING_Y) and the other divisors the fast paths use."""
    for d in (21, 2, 3, 5, 7, 9, 11, 13, 15, 17, 19):
        bad = engine.check_div_const(d)
        assert bad == 0, (d, bad)


@pytest.mark.parametrize("n,fs,frames,learn", [(1024, 1_024_000, 700, 40), (4096, 4_096_000, 420, 30)])
def test_device_tracker_mailbox_after_every_push(engine, n, fs, frames, learn):
    """K4 (the signal map on the device, csrc/track.cuh) against the oracle: the scene is pushed in pieces of 1..97 frames, many of
    them single frames, so the mailbox after each push is the oracle's list of that frame — starts, flushes, stops and time-outs
    included. Every fifth push asks for per-frame lists instead, which hands the map to the host tracker and back."""
    cfg, tones, iq, period = scene(n, fs, frames, learn)
    assert period == 1.0  # the frame clock is additive at any split point
    o = ol.OracleChain(cfg)
    ref = o.push(iq, frames, 500, period, dense=())
    band = b2s.Band(engine, cfg)
    sizes, k, i, checked = [1, 1, 1, 2, 5, 1, 20, 21, 1, 22, 97, 3, 1, 60, 1, 1, 80, 33, 1, 7], 0, 0, 0
    while k < frames:
        m = min(sizes[i % len(sizes)], frames - k)
        if i % 5 == 4:
            r = band.push(iq[k * 2 * n :], m, 500 + k, period, per_frame=True)
            assert _tx(r.frame_tx) == _tx(ref.frame_tx[k : k + m]), (k, m)
            got = r.transmissions
        else:
            res = band.push_raw(iq[k * 2 * n :].ctypes.data, m, 500 + k, period)
            got = [(t.shift_hz, t.flush, t.key, t.power) for t in res.transmissions[: res.n_transmissions]]
        want = ref.frame_tx[k + m - 1]
        if [(f, fl, key) for f, fl, key, _ in got] != [(f, fl, key) for f, fl, key, _ in want]:
            oo = ol.OracleChain(cfg)  # the oracle's map at the same frame, for the failure message
            oo.push(iq, k + m, 500, period, dense=())
            raise AssertionError((k, m, got, want, [x.tolist() for x in band.get_signals()[:3]], [x.tolist() for x in oo.get_signals()[:3]]))
        for (_, _, _, pa), (_, _, _, pb) in zip(got, want):
            assert abs(pa - pb) <= 4e-3
        checked += len(want)
        i += 1
        k += m
    assert checked > 50
    for a, b in zip(band.get_signals()[:3], o.get_signals()[:3]):
        assert np.array_equal(a, b)


def test_more_signals_than_the_result_struct_holds(engine):
    """The signal map is not limited to B2S_MAX_TX (the reference's std::map is unbounded): 80 simultaneous weak carriers (1 LSB each
    under 2 LSB of noise, so their summed Hamming sidelobes stay below the levels) are all tracked; the embedded array holds the 64
    strongest and says so, b2s_band_get_transmissions returns the whole list."""
    n, fs, frames, learn, count = 8192, 8_192_000, 80, 20, 80
    rng = np.random.default_rng(3)
    tones = [synth.Tone(-3900.1 + 97 * i, amplitude=1.0 + 0.02 * (i % 7), fm_dev_bins=4.0, phase=float(rng.uniform(0, 6.28))) for i in range(count)]
    iq = synth.make_iq_int8(n, frames, tones, seed=9, quiet_frames=learn, sigma=2.0)
    cfg = b2s.make_config(n, fs, learn_frames=learn, group_size_bins=16, min_time_ms=10, timeout_ms=30, detect_capacity=4096, start_level=3.0, stop_level=2.0)
    band = b2s.Band(engine, cfg)
    period = synth.frame_period_ms(n, fs)
    res = band.push_raw(iq.ctypes.data, frames, 0, period)
    o = ol.OracleChain(cfg)
    o.push(iq, frames, 0, period, dense=())
    want = [(f, fl, k) for f, fl, k, _ in o.get_transmissions()]
    assert len(want) >= count
    assert res.n_transmissions_total == len(want) and res.n_transmissions == b2s.MAX_TX
    assert [(f, fl, k) for f, fl, k, _ in band.get_transmissions()] == want
    assert _mailbox(res) == want[: b2s.MAX_TX]


def test_eight_concurrent_bands_with_hops(engine):
    """BASELINE config 3: eight bands (N = 8192, fs = 2.048 MS/s, one CUDA stream each) in ONE process, pushed asynchronously so
    their kernels overlap, with Transmission::resetBuffers (the Scanner's hop, scanner.cpp:46-60 / sdr_device.cpp:74) between some
    pushes and not others. Every band must follow ITS OWN oracle chain: mailbox after every push, signal map, noise, Averager ring."""
    n, fs, learn, piece, pieces = 8192, 2_048_000, 30, 125, 4
    period = synth.frame_period_ms(n, fs)  # 4 ms
    frames = piece * pieces
    bands, oracles, iqs = [], [], []
    for b in range(8):
        tones = [synth.Tone(-3000.1 + 700 * b, amplitude=60.0, fm_dev_bins=6.0, on_frames=[(learn + 20 + 3 * b, learn + 150 + 10 * b), (300 + 5 * b, 460)]),
                 synth.Tone(1000.1 + 300 * b, amplitude=50.0, fm_dev_bins=5.0, on_frames=[(learn + 60, 280 + 7 * b)], phase=1.0 + b)]
        iq = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(3, b), quiet_frames=learn)
        cfg = b2s.make_config(n, fs, center_hz=int(bench_centres()[b] * 1e6), learn_frames=learn, min_time_ms=80, timeout_ms=120, max_frames_per_push=piece,
                              flags=b2s.FLAG_ASYNC)
        ocfg = b2s.make_config(n, fs, center_hz=int(bench_centres()[b] * 1e6), learn_frames=learn, min_time_ms=80, timeout_ms=120)
        bands.append(b2s.Band(engine, cfg))
        oracles.append(ol.OracleChain(ocfg))
        iqs.append(iq)
    checked = 0
    for p in range(pieces):
        t0 = 1000 + int(p * piece * period)
        for band, iq in zip(bands, iqs):
            band.push_raw(iq[p * piece * 2 * n :].ctypes.data, piece, t0, period)
        for b, (band, o, iq) in enumerate(zip(bands, oracles, iqs)):
            res = band.sync()
            ref = o.push(iq[p * piece * 2 * n :], piece, t0, period, dense=())
            assert _mailbox(res) == [(f, fl, k) for f, fl, k, _ in ref.frame_tx[-1]], (p, b)
            checked += len(ref.frame_tx[-1])
        if p in (0, 2):  # hop: resetBuffers on every band (not drained first: the reset is ordered behind the pushes in flight)
            for band, o in zip(bands, oracles):
                band.reset()
                o.reset()
    assert checked > 20
    for band, o in zip(bands, oracles):
        for x, y in zip(band.get_signals()[:3], o.get_signals()[:3]):
            assert np.array_equal(x, y)
        thr_g, _, ready_g = band.get_noise()
        thr_o, _, ready_o = o.get_noise()
        assert ready_g and ready_o and np.max(np.abs(thr_g - thr_o)) <= 2e-3


def bench_centres():
    import bench

    return bench.HOP_CENTRES_MHZ
