"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64) of the reference's recorder DSP chain, SURVEY.md §8(f)#1.

Follows sources/radio/recorder.cpp:22-40,58-73 of the reference:
    rotator_cc(phase_inc = 2 pi (-shift) / fs)  ->  rational_resampler(f1, f2) for each pair of
    getResamplersFactors(fs, bandwidth, RESAMPLER_THRESHOLD)  ->  complex_to_interleaved_char(vector, 127.0)
getResamplersFactors / getPrimeFactors / split are in the reference tree (sources/utils/radio_utils.cpp:9-35,105-152) and are pinned by
its gtest vectors (tests/test_radio_utils.cpp:28-69, tests/golden/reference_kats.json) and by the compiled reference object.
The resampler itself is GNU Radio 3.10 (gr-filter: rational_resampler.cc design_resampler_filter, firdes.cc low_pass, gr-fft window.cc
kaiser / Izero; gr-blocks rotator, VOLK volk_32f_s32f_convert_8i) — NOT in the reference tree and not installable here: PARITY
UNPINNED for the taps, the FIR and the rounding; restated from those sources' published algorithms. History is zero at startRecording."""
import math

import numpy as np


def prime_factors(n):  # radio_utils.cpp:105-127
    if n == 1:
        return [1]
    f = []
    while n % 2 == 0:
        f.append(2)
        n //= 2
    i = 3
    while i <= math.sqrt(n):
        while n % i == 0:
            f.append(i)
            n //= i
        i += 2
    if n > 2:
        f.append(n)
    return f


def _split(value, out, threshold):  # radio_utils.cpp:9-35
    if threshold < value and len(prime_factors(value)) != 1:
        f1, f2 = 1, value
        for i in range(int(math.sqrt(value)), 0, -1):
            if value % i == 0:
                f1, f2 = i, value // i
                break
        for f in (f1, f2):
            if threshold < f:
                _split(f, out, threshold)
            else:
                out.append(f)
    else:
        out.append(value)


def resamplers_factors(sample_rate, bandwidth, threshold=125):  # radio_utils.cpp:129-152
    g = math.gcd(sample_rate, bandwidth)
    left, right = [], []
    _split(bandwidth // g, left, threshold)
    _split(sample_rate // g, right, threshold)
    while len(left) < len(right):
        left.append(1)
    while len(right) < len(left):
        right.append(1)
    return list(zip(sorted(left), sorted(right)))


def _izero(x):  # gr::fft::window Izero, IzeroEPSILON = 1e-21
    s = u = n = 1.0
    halfx = x / 2.0
    while True:
        t = halfx / n
        n += 1.0
        t *= t
        u *= t
        s += u
        if u < 1e-21 * s:
            return s


def design_resampler_taps(interpolation, decimation, fractional_bw=0.4):
    """rational_resampler design_resampler_filter -> firdes::low_pass(I, I, mid_transition_band, trans_width, WIN_KAISER, 7.0)."""
    f32 = np.float32
    beta, halfband = f32(7.0), f32(0.5)
    rate = f32(interpolation) / f32(decimation)
    if rate >= 1.0:
        trans_width = f32(halfband - f32(fractional_bw))
        mid = f32(halfband - f32(trans_width / f32(2.0)))
    else:
        trans_width = f32(rate * f32(halfband - f32(fractional_bw)))
        mid = f32(f32(rate * halfband) - f32(trans_width / f32(2.0)))
    gain = fs = float(interpolation)
    atten = float(beta) / 0.1102 + 8.7
    ntaps = int(atten * fs / (22.0 * float(trans_width)))
    if ntaps % 2 == 0:
        ntaps += 1
    ibeta, inm1 = 1.0 / _izero(float(beta)), 1.0 / (ntaps - 1)
    w = np.array([f32(_izero(float(beta) * math.sqrt(1.0 - (2 * i * inm1 - 1) ** 2)) * ibeta) for i in range(ntaps)], dtype=np.float32)
    m = (ntaps - 1) // 2
    fwt0 = 2 * math.pi * float(mid) / fs
    taps = np.zeros(ntaps, np.float32)
    for n in range(-m, m + 1):
        taps[n + m] = f32(fwt0 / math.pi * float(w[n + m])) if n == 0 else f32(math.sin(n * fwt0) / (n * math.pi) * float(w[n + m]))
    fmax = float(taps[m]) + 2.0 * float(np.sum(taps[m + 1 :].astype(np.float64)))
    return (taps.astype(np.float64) * (gain / fmax)).astype(np.float32)


def rotate(x, sample_rate, shift, start=0):
    """rotator_cc with phase_inc = 2 pi (-shift) / fs: x[n] * exp(i phase_inc n), the phase reduced exactly (integers) before the trig."""
    n = np.arange(start, start + len(x), dtype=np.int64)
    k = (-(shift) * n) % sample_rate  # turns * fs
    ang = 2.0 * np.pi * k.astype(np.float64) / float(sample_rate)
    return x * (np.cos(ang) + 1j * np.sin(ang))


def resample(x, taps, interp, decim):
    """y[m] = sum_k h[k] u[m D - k], u = x upsampled by I with zeros, zero history; outputs whose newest input sample has arrived."""
    from scipy.signal import upfirdn

    y = upfirdn(taps.astype(np.float64), x, up=interp, down=decim)
    n_out = (len(x) * interp - 1) // decim + 1
    return y[:n_out]


def recorder_chain(x, sample_rate, bandwidth, shift):
    """x: complex128 samples at sample_rate (already scaled to +-1). Returns interleaved int8 I/Q at `bandwidth` samples/s."""
    y = rotate(np.asarray(x, np.complex128), sample_rate, shift)
    for interp, decim in resamplers_factors(sample_rate, bandwidth):
        y = resample(y, design_resampler_taps(interp, decim), interp, decim)
    out = np.empty(2 * len(y), np.int8)
    out[0::2] = np.clip(np.rint(y.real * 127.0), -128, 127).astype(np.int8)  # volk_32f_s32f_convert_8i: rint, saturate
    out[1::2] = np.clip(np.rint(y.imag * 127.0), -128, 127).astype(np.int8)
    return out
