"""Oracle checks for the part of the path that lives outside /root/reference (GNU Radio window + FFT + shift) and for
PSD::work (sources/radio/blocks/psd.cpp:18-20). The reference has no fixture here ("parity unpinned"),
so the restatement is cross-checked against numpy.fft, against the semantics of the reference's own offline tool
(scripts/converter.py:17-21: fft -> |x|^2/fs -> 10 log10 -> fftshift) and against closed-form known answers."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_b2s

b2s = load_b2s()


def _fft(n, x, f32=False):
    inp = np.ascontiguousarray(np.stack([x.real, x.imag], -1).astype(np.float32))
    out = np.empty_like(inp)
    fn = ol.oracle().orc_fft_f32 if f32 else ol.oracle().orc_fft_f64
    fn(n, inp.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out[:, 0].astype(np.float64) + 1j * out[:, 1].astype(np.float64)


@pytest.mark.parametrize("n", [2, 8, 256, 4096, 16384, 32768])
def test_fft_f64_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.fft(x.astype(np.complex128))
    got = _fft(n, x)
    scale = np.sqrt(np.mean(np.abs(ref) ** 2))
    assert np.max(np.abs(got - ref)) <= 2e-7 * scale * 4  # only the final fp32 rounding separates them


@pytest.mark.parametrize("n", [4, 8, 1024, 4096, 8192, 16384, 32768])
def test_fft_f32_baseline_is_a_correct_fft(n):
    rng = np.random.default_rng(n + 1)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.fft(x.astype(np.complex128))
    got = _fft(n, x, f32=True)
    assert np.sqrt(np.mean(np.abs(got - ref) ** 2)) <= 2e-6 * np.sqrt(np.mean(np.abs(ref) ** 2))


def test_hamming_window():
    for n in (16, 4096, 16384):
        w = ol.hamming(n)
        k = np.arange(n)
        expect = (0.54 - 0.46 * np.cos(2 * np.pi * k / (n - 1))).astype(np.float32)
        assert np.array_equal(w, expect)
        assert np.array_equal(w, w[::-1])  # symmetric
        assert abs(w[0] - 0.08) < 1e-6 and (n < 4096 or abs(w[n // 2] - 1.0) < 1e-3)


def _cfg(n, fs=2_048_000, fmt=b2s.IQ_CS8, scale=1.0 / 127.0):
    return b2s.make_config(n, fs, iq_format=fmt, iq_scale=scale)


def test_psd_matches_converter_semantics_without_window():
    """scripts/converter.py:17-21 with a rectangular window == oracle with user taps of ones."""
    n, fs = 1024, 2_048_000
    rng = np.random.default_rng(3)
    iq = rng.integers(-100, 100, 2 * n).astype(np.int8)
    cfg = _cfg(n, fs, scale=1.0 / 127.5)  # converter.py:33 reads cs8 as int8/127.5
    ones = np.ones(n, np.float32)
    got = ol.oracle_psd_frame(cfg, iq, window=ones)
    x = (iq[0::2].astype(np.float32) * np.float32(1 / 127.5)) + 1j * (iq[1::2].astype(np.float32) * np.float32(1 / 127.5))
    d = np.fft.fft(x.astype(np.complex128))
    ref = np.fft.fftshift(10.0 * np.log10(np.abs(d**2.0) / np.float32(fs)))
    assert np.max(np.abs(got - ref)) < 2e-4  # dB


def test_psd_known_answers():
    n, fs = 4096, 2_048_000
    cfg = _cfg(n, fs, scale=1.0)
    w = ol.hamming(n).astype(np.float64)
    # impulse at n=0 -> flat spectrum |w[0]|^2
    iq = np.zeros(2 * n, np.int8)
    iq[0] = 100
    p = ol.oracle_psd_frame(cfg, iq)
    expect = 10 * np.log10((100 * w[0]) ** 2 / fs)
    assert np.max(np.abs(p - expect)) < 1e-3
    # DC -> Hamming main lobe centred at out[N/2], peak = (A * sum w)^2 / fs
    iq = np.zeros(2 * n, np.int8)
    iq[0::2] = 50
    p = ol.oracle_psd_frame(cfg, iq)
    assert int(np.argmax(p)) == n // 2
    assert abs(p[n // 2] - 10 * np.log10((50 * w.sum()) ** 2 / fs)) < 1e-3
    # complex tone at +fs/4 lands at out[3N/4]
    k = np.arange(n)
    z = 60 * np.exp(2j * np.pi * k / 4)
    iq = np.empty(2 * n, np.int8)
    iq[0::2] = np.rint(z.real)
    iq[1::2] = np.rint(z.imag)
    p = ol.oracle_psd_frame(cfg, iq)
    assert int(np.argmax(p)) == 3 * n // 4
    # fs only shifts the level by -10 log10(fs)
    p2 = ol.oracle_psd_frame(_cfg(n, 20_000_000, scale=1.0), iq)
    assert np.max(np.abs((p - p2) - 10 * np.log10(20_000_000 / fs))[np.isfinite(p)]) < 2e-3


def test_cf32_and_cs8_inputs_agree():
    n = 2048
    rng = np.random.default_rng(5)
    iq = rng.integers(-128, 128, 2 * n).astype(np.int8)
    a = ol.oracle_psd_frame(_cfg(n, scale=1.0 / 127.0), iq)
    f = (iq.astype(np.float32) * np.float32(1.0 / 127.0)).astype(np.float32)
    b = ol.oracle_psd_frame(_cfg(n, fmt=b2s.IQ_CF32), f)
    assert np.array_equal(a, b)


def power_parity_stats(got_lin, ref_lin):
    """The power-bin parity statistics used by the GPU parity tests (DESIGN.md "Parity criterion"):
    floored criterion |p - p_ref| <= 1e-5 * max(p_ref, median(p_ref)) per bin -> pass fraction and worst ratio,
    plus the strict per-bin fraction and the L2-relative error."""
    ref, got = np.asarray(ref_lin, np.float64), np.asarray(got_lin, np.float64)
    floor = np.maximum(ref, np.median(ref, axis=-1, keepdims=True))
    rel = np.abs(got - ref) / floor
    strict = np.abs(got - ref) / np.maximum(ref, 1e-300)
    return {
        "pass_frac": float(np.mean(rel <= 1e-5)),
        "worst": float(rel.max()),
        "strict_frac": float(np.mean(strict <= 1e-5)),
        "l2_rel": float(np.sqrt(np.sum((got - ref) ** 2) / np.sum(ref**2))),
    }


def test_fp32_error_cloud_vs_the_stated_tolerance():
    """SURVEY.md §7 hard part 1, measured: ANY fp32 FFT (here the oracle's own fp32 variant; scipy's pocketfft in fp32
    behaves the same: 99.91 % / worst 3.0e-5) leaves a ~0.1 % tail of noise-level bins above the floored 1e-5 criterion.
    The GPU parity tests therefore assert: floored pass fraction >= 99.5 %, worst <= 1e-4, L2-relative <= 1e-6 and
    report the strict per-bin fraction. This test pins those numbers for a plain fp32 CPU FFT."""
    n, fs = 16384, 20_000_000
    rng = np.random.default_rng(9)
    k = np.arange(n)
    z = 8 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for b in (1234.1, -3000.1, 77.1, 6000.1):
        z += 40 * np.exp(2j * np.pi * b * k / n)
    iq = np.empty(2 * n, np.int8)
    iq[0::2] = np.clip(np.rint(z.real), -128, 127)
    iq[1::2] = np.clip(np.rint(z.imag), -128, 127)
    cfg = _cfg(n, fs)
    _, ref = ol.oracle_psd_frame(cfg, iq, want_linear=True)
    cfg32 = b2s.BandConfig.from_buffer_copy(cfg)
    cfg32.flags |= 1
    _, got = ol.oracle_psd_frame(cfg32, iq, want_linear=True)
    st = power_parity_stats(got, ref)
    assert st["pass_frac"] >= 0.995 and st["worst"] <= 1e-4 and st["l2_rel"] <= 1e-6, st
