"""The compiled drop-in boundary: GpuScanChain (rtl-sdr-scanner-cpp_b200/host/gpu_scan_chain.h) — a gr::sync_block
with the work() shape of the reference's blocks — built against the reference's OWN headers (config.h, notification.h,
radio/help_structures.h, network/data_controller.h) and driven the way GNU Radio and SdrDevice drive the chain it replaces
(sources/radio/sdr_device.cpp:161-171): CF32 items of fftSize * decimatorFactor samples into work(), the transmission list out of
the TransmissionNotification mailbox, spectrogram rows out of DataController / Mqtt.

Compared, frame by frame, with the REFERENCE'S OWN compiled NoiseLearner / Transmission / Spectrogram objects fed the PSD rows the
GPU computes for the same items (b2s_psd: the same kernel), under the same injected clock.

The harness (oracle/gpu_chain_shim.cpp) lives in oracle/_ref/libref.so, which is built in the container that has /root/reference."""
import ctypes as C
import json
import struct

import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_b2s
from test_oracle_chain import synth

b2s = load_b2s()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (ol.have_ref() and hasattr(ol.ref(), "gpuchain_create")), reason="oracle/_ref/libref.so with the wrapper harness is not built")]

T0 = 1_700_000_000_000
FS, BANDWIDTH, CENTER = 2_048_000, 32_000, 145_000_000  # the reference derives N = 8192, decimator 5, indexStep 128 from these


def _harness():
    L = ol.ref()
    L.gpuchain_create.restype = C.c_void_p
    L.gpuchain_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64]
    L.gpuchain_destroy.argtypes = [C.c_void_p]
    L.gpuchain_last_error.restype = C.c_char_p
    L.gpuchain_item_bytes.restype = C.c_long
    L.gpuchain_fft_size.argtypes = [C.c_void_p]
    L.gpuchain_decimator.argtypes = [C.c_void_p]
    L.gpuchain_reset.argtypes = [C.c_void_p]
    L.gpuchain_set_center.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.gpuchain_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    return L


def _published(L):
    out = []
    for i in range(L.ref_published_count()):
        topic = C.create_string_buffer(128)
        buf = np.empty(1 << 16, dtype=np.uint8)
        k = L.ref_published_get(i, topic, 128, buf.ctypes.data_as(C.c_void_p), buf.size)
        out.append((topic.value.decode(), buf[:k].tobytes()))
    L.ref_published_clear()
    return out


def test_gpu_scan_chain_follows_the_reference_chain(engine):
    L = _harness()
    n = b2s.get_fft(FS, 250)
    r = b2s.lib().b2s_decimator_factor(FS, n)
    assert (n, r) == (8192, 5) and L.gpuchain_item_bytes(FS) == 8 * n * r  # sdr_device.cpp:149-152,161-162
    period = 1000.0 * n * r / FS  # 20 ms per item
    learn = b2s.lib().b2s_learn_frames_from_ms(2000, period)  # NOISE_LEARNING_TIME
    frames = 360
    conf = {"ignored": [[CENTER + 300_000, CENTER + 340_000]], "bandwidth": BANDWIDTH, "min_time_ms": 200, "timeout_ms": 300, "tuning_step": 2500}
    lo, hi = CENTER - 900_000, CENTER + FS // 2
    # the scene: keyed FM carriers; frame k of the engine = the first N samples of item k (Decimator, decimator.h:16-22)
    tones = synth.standard_scene(n, frames, learn)
    tones.append(synth.Tone((320_000 / (FS / n)) + 0.1, amplitude=50.0, fm_dev_bins=5.0))   # inside the ignored range
    tones.append(synth.Tone((-950_000 / (FS / n)) + 0.1, amplitude=50.0, fm_dev_bins=5.0))  # below the scanned range
    i8 = synth.make_iq_int8(n, frames, tones, seed=synth.seed_for(8), quiet_frames=learn, stride=n * r)
    items = (i8.astype(np.float32) * np.float32(1.0 / 127.0)).reshape(frames, 2 * n * r)  # what a CF32 SoapySDR stream carries
    now = [T0 + int(np.floor(k * period + 0.5)) for k in range(frames)]

    h = C.c_void_p(L.gpuchain_create(json.dumps(conf).encode(), FS, CENTER, lo, hi, 8.0, 5.0, now[0]))
    assert h, L.gpuchain_last_error()
    assert (L.gpuchain_fft_size(h), L.gpuchain_decimator(h)) == (n, r)
    L.ref_published_clear()
    got_lists = []
    freq, flush = np.zeros(256, np.int32), np.zeros(256, np.int32)
    k = 0
    for m in [1, 1, 2, 3, 1, 4, 8, 1] * 100:  # GNU Radio hands work() a few items at a time; the mailbox holds the last frame's list
        if k >= frames:
            break
        m = min(m, frames - k)
        if m > 1:  # the injected clock stamps the first frame of a call; later frames follow by the frame period
            pass
        cnt = L.gpuchain_work(h, items[k].ctypes.data_as(C.c_void_p), m, now[k], freq.ctypes.data_as(C.c_void_p), flush.ctypes.data_as(C.c_void_p), 256)
        assert cnt >= 0
        got_lists.append((k + m - 1, [(int(freq[i]), int(flush[i])) for i in range(cnt)]))
        k += m
    got_pub = [p for t, p in _published(L) if t == "sdr/dev/spectrogram"]
    L.gpuchain_destroy(h)

    # the reference's own blocks on the GPU's PSD rows of the same items
    cfg = b2s.make_config(n, FS, center_hz=CENTER, decimator=r, iq_format=b2s.IQ_CF32, learn_frames=learn, recording_bandwidth_hz=BANDWIDTH, min_time_ms=200,
                          timeout_ms=300, ignored=[tuple(conf["ignored"][0])], range_hz=(lo, hi))
    psd = engine.psd(cfg, items.reshape(-1), frames)
    ref = ol.RefBlocksChain(cfg, now[0], BANDWIDTH, with_spectrogram=True)
    ref_lists = []
    for k in range(frames):
        _, tx = ref.push_row(psd[k], now[k])
        ref_lists.append(tx)
    ref_pub = [p for t, p in ref.published() if t == "sdr/dev/spectrogram"]

    busy = 0
    for last, lst in got_lists:
        assert lst == ref_lists[last], (last, lst, ref_lists[last])
        busy += 1 if lst else 0
    assert busy > 40 and sum(fl for _, l in got_lists for _, fl in l) > 10  # starts, flushes and stops really happened
    all_ref = {f for l in ref_lists for f, _ in l}
    assert not any(abs(f - 320_000) < 30_000 for f in all_ref) and not any(f < -900_000 for f in all_ref)  # ignored / out of range stay silent
    # spectrogram messages: same times, same headers; rows byte for byte from the second on (the reference's first divisor is uninitialised, spectrogram.cpp:9)
    assert len(got_pub) == len(ref_pub) >= 5
    for i, (a, b) in enumerate(zip(got_pub, ref_pub)):
        assert a[:24] == b[:24], i
        assert struct.unpack("<QiiiI", a[:24])[4] == 2048
        if i >= 1:
            assert a[24:] == b[24:], f"spectrogram row {i}"


def test_gpu_scan_chain_reset_and_retune(engine):
    """SdrDevice::setFrequencyRange (sdr_device.cpp:66-77): resetBuffers on every hop, noise learnt per centre frequency."""
    L = _harness()
    n, r = 8192, 5
    period = 1000.0 * n * r / FS
    learn = b2s.lib().b2s_learn_frames_from_ms(2000, period)
    frames = learn + 80
    conf = {"ignored": [], "bandwidth": BANDWIDTH, "min_time_ms": 100, "timeout_ms": 200, "tuning_step": 2500}
    tones = [synth.Tone(500.1, amplitude=60.0, fm_dev_bins=6.0)]
    i8 = synth.make_iq_int8(n, frames, tones, seed=5, quiet_frames=learn, stride=n * r)
    items = (i8.astype(np.float32) * np.float32(1.0 / 127.0)).reshape(frames, 2 * n * r)
    h = C.c_void_p(L.gpuchain_create(json.dumps(conf).encode(), FS, CENTER, CENTER - FS // 2, CENTER + FS // 2, 8.0, 5.0, T0))
    assert h, L.gpuchain_last_error()
    freq, flush = np.zeros(64, np.int32), np.zeros(64, np.int32)

    def work(k0, m, t):
        return L.gpuchain_work(h, items[k0].ctypes.data_as(C.c_void_p), m, t, freq.ctypes.data_as(C.c_void_p), flush.ctypes.data_as(C.c_void_p), 64)

    assert work(0, frames, T0) >= 1  # the carrier is tracked after learning
    L.gpuchain_reset(h)
    t1 = T0 + int(frames * period)
    assert work(learn, 5, t1) == 0  # signals dropped, Averager empty: nothing before GROUPING_Y frames have passed again
    assert work(learn + 5, 40, t1 + int(5 * period)) >= 1  # the noise floor of this centre was kept: detection resumes without re-learning
    L.gpuchain_reset(h)
    L.gpuchain_set_center(h, CENTER + 5_000_000, CENTER + 5_000_000 - FS // 2, CENTER + 5_000_000 + FS // 2)
    t2 = t1 + 10_000
    assert work(learn, 60, t2) == 0  # a new centre learns its own noise floor first (noise_learner.cpp:41-42)
    L.gpuchain_destroy(h)
