"""Pins the oracle (and the wire-format helpers) against the REFERENCE'S OWN COMPILED BLOCKS.

oracle/_ref/libref.so is built by `make -C oracle ref` from the unmodified sources under /root/reference:
sources/radio/blocks/{psd,noise_learner,transmission,spectrogram}.cpp, sources/radio/{signal,averager}.cpp,
sources/network/data_controller.cpp, sources/utils/*.cpp (see oracle/ref_blocks_shim.cpp for the name-only stand-ins of
GNU Radio / Paho / Config and for the injected clock). The same PSD rows go through the reference objects and through
oracle/scan_oracle.cpp with the same frame clock; everything must agree exactly:
  * PSD::work                      bit for bit on identical spectra
  * NoiseLearner::work             every output row bit for bit, the length of the learning phase
  * Transmission::work             the FrequencyFlush list handed to TransmissionNotification, every frame
  * Spectrogram::work + DataController::pushSpectrogram   every published payload, byte for byte
The window and the FFT itself are gr::fft::fft_v (GNU Radio + FFTW, not in the reference tree): those stay unpinned."""
import ctypes as C
import struct

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT, load_b2s

b2s = load_b2s()
import sys  # noqa: E402

sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

synth = ge.load_synth()

pytestmark = pytest.mark.skipif(not (ol.have_ref() and ol.have_ref_blocks()), reason="oracle/_ref/libref.so (reference blocks) not built: needs /root/reference")

NOISE_LEARNING_MS = 2000  # NOISE_LEARNING_TIME, sources/config.h:24 (compiled into the reference objects)
PERIOD_MS = 25.0          # frame clock of the test: 2000 ms of learning = 81 frames
T0 = 1_700_000_000_000


def _scene(n=1024, fs=2_048_000, frames=520, **kw):
    learn = b2s.lib().b2s_learn_frames_from_ms(NOISE_LEARNING_MS, PERIOD_MS)
    bw = 16 * fs // n
    cfg = b2s.make_config(n, fs, learn_frames=learn, recording_bandwidth_hz=bw, min_time_ms=200, timeout_ms=300, **kw)
    tones = synth.standard_scene(n, frames, learn)
    iq = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(7), quiet_frames=learn)
    return cfg, iq, frames, learn, bw


def _now(k):
    return T0 + int(np.floor(k * PERIOD_MS + 0.5))


def test_psd_work_bit_exact():
    rng = np.random.default_rng(3)
    n, items, fs = 512, 5, 2_048_000
    x = (rng.standard_normal(2 * n * items) * rng.choice([1e-3, 1.0, 300.0], 2 * n * items)).astype(np.float32)
    x[:4] = 0.0  # |z| = 0 -> -inf in both
    got = np.empty(n * items, dtype=np.float32)
    want = np.empty(n * items, dtype=np.float32)
    ol.ref().ref_psd_work(n, fs, x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), items)
    ol.oracle().orc_psd_from_spectrum(n, fs, x.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), items)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.isneginf(got[0])


def test_noise_learner_and_transmission_follow_the_reference_objects_frame_by_frame():
    cfg, iq, frames, learn, bw = _scene()
    orc = ol.OracleChain(cfg)
    r = orc.push(iq, frames, T0, PERIOD_MS, dense=("psd_db", "noise_sub_db"))
    ref = ol.RefBlocksChain(cfg, _now(0), bw, with_spectrogram=False)
    n_tx_frames, flushes = 0, 0
    for k in range(frames):
        q, tx = ref.push_row(r.psd_db[k], _now(k))
        assert np.array_equal(q.view(np.uint32), r.noise_sub_db[k].view(np.uint32)), f"NoiseLearner row {k}"
        assert tx == [(f, fl) for f, fl, _, _ in r.frame_tx[k]], f"transmission list of frame {k}"
        n_tx_frames += 1 if tx else 0
        flushes += sum(fl for _, fl in tx)
    # the learning phase had the reference's own length: -100 rows, then real ones
    assert np.all(r.noise_sub_db[learn - 1] == -100.0) and not np.all(r.noise_sub_db[learn] == -100.0)
    assert n_tx_frames > 100 and flushes > 20, (n_tx_frames, flushes)  # the scene really exercises start / flush / stop / time-out


def test_hop_and_reset_follow_the_reference_objects():
    cfg, iq, frames, learn, bw = _scene(frames=420)
    orc = ol.OracleChain(cfg)
    ref = ol.RefBlocksChain(cfg, _now(0), bw, with_spectrogram=False)
    half = 300
    r = orc.push(iq[: half * cfg.fft_size * 2], half, T0, PERIOD_MS, dense=("psd_db", "noise_sub_db"))
    for k in range(half):
        q, tx = ref.push_row(r.psd_db[k], _now(k))
        assert tx == [(f, fl) for f, fl, _, _ in r.frame_tx[k]]
    # SdrDevice::setFrequencyRange (sdr_device.cpp:66-77): resetBuffers, new centre -> a fresh noise entry is learned
    orc.reset()
    ref.reset()
    new_c = cfg.center_hz + 1_000_000
    orc.set_center(new_c, new_c - cfg.sample_rate_hz // 2, new_c + cfg.sample_rate_hz // 2)
    ref.set_center(new_c, new_c - cfg.sample_rate_hz // 2, new_c + cfg.sample_rate_hz // 2)
    rest = frames - half
    t1 = _now(half)
    r2 = orc.push(iq[half * cfg.fft_size * 2 :], rest, t1, PERIOD_MS, dense=("psd_db", "noise_sub_db"))
    for k in range(rest):
        q, tx = ref.push_row(r2.psd_db[k], t1 + int(np.floor(k * PERIOD_MS + 0.5)))
        assert np.array_equal(q.view(np.uint32), r2.noise_sub_db[k].view(np.uint32)), f"row {k} after the hop"
        assert tx == [(f, fl) for f, fl, _, _ in r2.frame_tx[k]], f"frame {k} after the hop"


def test_noise_learning_follows_the_reference_clock_under_a_hop_schedule():
    """Scanner hops every 500 ms (RANGE_SCANNING_TIME) between two centres: the reference's Noise finishes NOISE_LEARNING_TIME after the FIRST
    visit of a centre (noise_learner.cpp:11,23), the time spent on the other centre included — not after 2 s of dwell. The oracle's clock rule
    (noise_learning_ms) must give the same learning frames, rows and transmission lists as the compiled NoiseLearner / Transmission."""
    n, fs, dwell, hops = 1024, 2_048_000, 20, 14  # 20 frames x 25 ms = 500 ms per visit
    bw = 16 * fs // n
    cfg = b2s.make_config(n, fs, learn_frames=10_000, noise_learning_ms=NOISE_LEARNING_MS, recording_bandwidth_hz=bw, min_time_ms=200, timeout_ms=300)
    frames = dwell * hops
    tones = synth.standard_scene(n, frames, 0)
    iq = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(11), quiet_frames=0)
    centres = [cfg.center_hz, cfg.center_hz + 3_000_000]
    orc = ol.OracleChain(cfg)
    ref = ol.RefBlocksChain(cfg, _now(0), bw, with_spectrogram=False)
    learning_frames = {c: 0 for c in centres}
    first_real = {}
    for h in range(hops):
        c = centres[h % 2]
        orc.reset(), ref.reset()  # SdrDevice::setFrequencyRange: resetBuffers on every retune (sdr_device.cpp:74)
        orc.set_center(c, c - fs // 2, c + fs // 2), ref.set_center(c, c - fs // 2, c + fs // 2)
        k0 = h * dwell
        r = orc.push(iq[k0 * 2 * n : (k0 + dwell) * 2 * n], dwell, _now(k0), PERIOD_MS, dense=("psd_db", "noise_sub_db"))
        for k in range(dwell):
            q, tx = ref.push_row(r.psd_db[k], _now(k0) + int(np.floor(k * PERIOD_MS + 0.5)))
            assert np.array_equal(q.view(np.uint32), r.noise_sub_db[k].view(np.uint32)), f"hop {h} frame {k}"
            assert tx == [(f, fl) for f, fl, _, _ in r.frame_tx[k]], f"hop {h} frame {k}"
            if np.all(q == -100.0):
                learning_frames[c] += 1
            else:
                first_real.setdefault(c, k0 + k)
    # centre 0: first frame at 0 ms, visits at 0, 1000, 2000 ms -> the frame stamped 2000 ms (frame 80) is its last learning frame;
    # centre 1: first frame at 500 ms -> the frame stamped 2500 ms (frame 100). 41 learning frames each instead of 81 frames of dwell.
    assert first_real == {centres[0]: 81, centres[1]: 101}, first_real
    assert learning_frames == {centres[0]: 41, centres[1]: 41}, learning_frames


@pytest.mark.parametrize("n,fs", [(2048, 2_048_000), (8192, 2_048_000)])
def test_spectrogram_payloads_byte_for_byte(n, fs):
    """Output size = min(16384, getFft(fs, 1000)) = 2048 (spectrogram.cpp:14). n = 2048: decimator factor 1; n = 8192 (the
    reference's own detection FFT at 2.048 MS/s, getFft(fs, 250)): factor 4, mean of adjacent bins (spectrogram.cpp:50-58).
    (N < 2048 is not a valid geometry for the reference at this rate: m_inputSize / m_outputSize would be 0.)"""
    cfg, iq, frames, learn, bw = _scene(n=n, fs=fs, frames=260)
    assert cfg.spectrogram_out_size == b2s.get_fft(fs, 1000) == 2048
    orc = ol.OracleChain(cfg)
    r = orc.push(iq, frames, T0, PERIOD_MS, dense=("psd_db",))
    ref = ol.RefBlocksChain(cfg, _now(0), bw, with_spectrogram=True)
    for k in range(frames):
        ref.push_row(r.psd_db[k], _now(k))
    pub = [p for t, p in ref.published() if t == "sdr/dev/spectrogram"]
    times, centers, rows = orc.get_spectrogram(cap=64)
    assert len(pub) == len(times) >= 5
    m = cfg.spectrogram_out_size
    # Spectrogram::Container::m_counter is never initialised (spectrogram.cpp:9): the divisor of the FIRST row is whatever the
    # allocator left there. Every later row is well defined; compare those strictly and the first one's header only.
    for i, payload in enumerate(pub):
        t_ms, start, stop, step, size = struct.unpack("<QiiiI", payload[:24])
        assert (t_ms, size) == (int(times[i]), m)
        assert (start, stop, step) == (cfg.center_hz - fs // 2, cfg.center_hz + fs // 2, fs // m)
        if i >= 1:
            assert payload[24:] == rows[i].tobytes(), f"spectrogram row {i}"
            assert payload == b2s.pack_spectrogram_message(int(times[i]), int(centers[i]), fs, rows[i])


def test_wire_messages_against_the_reference_data_controller():
    rng = np.random.default_rng(11)
    R = ol.ref()
    for size in (3, 1000, 16384):
        R.ref_published_clear()
        t_ms, f, fs = int(rng.integers(1, 2**41)), int(rng.integers(10**8, 10**9)), 20_000_000
        row = rng.integers(-128, 128, size).astype(np.int8)
        iq = rng.integers(-128, 128, 2 * size).astype(np.int8)
        R.ref_push_spectrogram(t_ms, f, fs, row.ctypes.data_as(C.c_void_p), size)
        R.ref_push_transmission(t_ms, f, fs, iq.ctypes.data_as(C.c_void_p), size)
        got = {}
        for i in range(R.ref_published_count()):
            topic = C.create_string_buffer(128)
            buf = np.empty(1 << 16, dtype=np.uint8)
            k = R.ref_published_get(i, topic, 128, buf.ctypes.data_as(C.c_void_p), buf.size)
            got[topic.value.decode()] = buf[:k].tobytes()
        assert got["sdr/dev/spectrogram"] == b2s.pack_spectrogram_message(t_ms, f, fs, row)
        assert got["sdr/dev/transmission/uint8"] == b2s.pack_transmission_message(t_ms, f, fs, iq)
