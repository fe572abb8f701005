"""CPU-side checks of the product boundary (no GPU needed): the C-ABI library loads, exports every symbol that
include/b2s.h declares, its host helpers obey the reference's known-answer vectors, and — because there is no CPU
fallback — engine creation fails loudly when no B200 is present."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_b2s

b2s = load_b2s()
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))


def _have_lib():
    return os.path.exists(b2s.LIB_PATH)


pytestmark = pytest.mark.skipif(not _have_lib(), reason="libb2s.so not built; run __graft_entry__.build()")


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "b2s.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 30
    lib = C.CDLL(b2s.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_the_header():
    """ctypes mirror vs the C compiler's view of include/b2s.h (sizeof + a few offsets through a tiny probe)."""
    import subprocess, tempfile

    src = r"""
#include <stdio.h>
#include <stddef.h>
#include "b2s.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(b2s_band_config), offsetof(b2s_band_config, window_taps), offsetof(b2s_band_config, min_time_ms),
         offsetof(b2s_band_config, detect_capacity), sizeof(b2s_result), offsetof(b2s_result, frame_tx_count), sizeof(b2s_transmission));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    B, R = b2s.BandConfig, b2s.Result
    assert got == [C.sizeof(B), B.window_taps.offset, B.min_time_ms.offset, B.detect_capacity.offset, C.sizeof(R), R.frame_tx_count.offset, C.sizeof(b2s.Transmission)]


def test_host_helpers_obey_reference_kats():
    for fs, step, expect in G["get_fft"]["cases"]:
        assert b2s.get_fft(fs, step) == expect
    for f, step, expect in G["get_tuned_frequency"]["cases"]:
        assert b2s.get_tuned_frequency(f, step) == expect
    g = G["get_max_index"]
    for index, group, expect in g["cases"]:
        assert b2s.get_max_index(np.array(g["data"], np.float32), index, group) == expect
    g = G["contains_with_margin"]
    for index, margin, expect in g["cases"]:
        assert (b2s.contains_with_margin(g["keys"], index, margin) is not None) == expect
    for data, expect in G["most_frequent_value"]["cases"]:
        assert b2s.most_frequent_value(data) == expect


def test_host_helpers_match_oracle_on_random_inputs():
    import oracle_lib as ol

    rng = np.random.default_rng(21)
    O = ol.oracle()
    for _ in range(300):
        n = int(rng.integers(1, 300))
        d = rng.integers(-4, 5, n).astype(np.float32)
        idx, grp = int(rng.integers(0, n)), int(rng.integers(0, 130))
        assert b2s.get_max_index(d, idx, grp) == O.orc_get_max_index(d.ctypes.data_as(C.c_void_p), n, idx, grp)
        v = rng.integers(0, 9, int(rng.integers(1, 24))).astype(np.int32)
        assert b2s.most_frequent_value(v) == O.orc_most_frequent_value(v.ctypes.data_as(C.c_void_p), len(v))
        f, step = int(rng.integers(-3_000_000, 3_000_000)), int(rng.integers(1, 50_000))
        assert b2s.get_tuned_frequency(f, step) == O.orc_get_tuned_frequency(f, step)


def test_wire_messages_match_the_reference_layout():
    """MQTT payloads of DataController (network/data_controller.cpp:27-57): an independent struct.pack statement of the
    layout, the oracle restatement and the library agree byte for byte; short buffers are refused with the needed size."""
    import struct

    import oracle_lib as ol

    rng = np.random.default_rng(5)
    O = ol.oracle()
    for size in (1, 7, 256, 8192):
        t_ms = int(rng.integers(1, 2**40))
        f = int(rng.integers(50_000_000, 1_500_000_000))
        fs = int(rng.choice([2_048_000, 20_000_000, 40_000_000]))
        row = rng.integers(-128, 128, size).astype(np.int8)
        want = struct.pack("<Qiii", t_ms, f - fs // 2, f + fs // 2, fs // size) + struct.pack("<I", size) + row.tobytes()
        assert b2s.pack_spectrogram_message(t_ms, f, fs, row) == want
        buf = np.empty(24 + size, dtype=np.uint8)
        n = O.orc_spectrogram_message(t_ms, f, fs, row.ctypes.data_as(C.c_void_p), size, buf.ctypes.data_as(C.c_void_p))
        assert buf[:n].tobytes() == want

        iq = rng.integers(-128, 128, 2 * size).astype(np.int8)
        want = struct.pack("<QiiI", t_ms, f - fs // 2, f + fs // 2, fs) + (iq.view(np.uint8) ^ 0x80).tobytes()
        assert b2s.pack_transmission_message(t_ms, f, fs, iq) == want
        buf = np.empty(20 + 2 * size, dtype=np.uint8)
        n = O.orc_transmission_message(t_ms, f, fs, iq.ctypes.data_as(C.c_void_p), size, buf.ctypes.data_as(C.c_void_p))
        assert buf[:n].tobytes() == want
    # the extremes of int8 map to 0x00 / 0xff / 0x80 (offset binary), as the sdr-hub decoder expects
    edge = np.array([-128, 127, 0, -1], dtype=np.int8)
    assert b2s.pack_transmission_message(0, 0, 0, edge)[20:] == bytes([0x00, 0xFF, 0x80, 0x7F])
    lib = C.CDLL(b2s.LIB_PATH)
    lib.b2s_pack_spectrogram_message.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    need = C.c_size_t(0)
    row = np.zeros(100, dtype=np.int8)
    small = np.zeros(10, dtype=np.uint8)
    assert lib.b2s_pack_spectrogram_message(1, 2, 3, row.ctypes.data_as(C.c_void_p), 100, small.ctypes.data_as(C.c_void_p), 10, C.byref(need)) != 0
    assert need.value == 124


def test_default_config_follows_setup_chains():
    """sdr_device.cpp:148-152 + config.h:24-38 for the two sample rates the reference's tests use."""
    cfg = b2s.BandConfig()
    b2s.lib().b2s_default_config(C.byref(cfg), 2_048_000, 144_000_000, 32_000)
    assert (cfg.fft_size, cfg.frame_stride_samples // cfg.fft_size, cfg.group_size_bins) == (8192, 5, 128)
    assert (cfg.grouping_x, cfg.grouping_y, cfg.start_level, cfg.stop_level) == (21, 21, 8.0, 5.0)
    assert cfg.spectrogram_out_size == 2048
    b2s.lib().b2s_default_config(C.byref(cfg), 20_000_000, 150_000_000, 32_000)
    assert (cfg.fft_size, cfg.frame_stride_samples // cfg.fft_size, cfg.group_size_bins) == (131072, 3, 210)
    # learning: first frame k with t_k >= 2000 ms completes it (noise_learner.cpp:23)
    period = 8192 * 5 * 1000.0 / 2_048_000
    assert b2s.lib().b2s_learn_frames_from_ms(2000, period) == 101
    assert cfg.noise_learning_ms == 2000  # NOISE_LEARNING_TIME as the reference's clock rule (config.h:24, noise_learner.cpp:23)
    assert C.sizeof(b2s.BandConfig) % 8 == 0 and b2s.BandConfig.noise_learning_ms.offset == C.sizeof(b2s.BandConfig) - 8  # appended: older fields keep their offsets


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(b2s.B2SError) as e:
        b2s.Engine(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package may include, import or link it."""
    pkg = os.path.join(ROOT, "rtl-sdr-scanner-cpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "scan_oracle" not in text and "liboracle" not in text and "oracle_lib" not in text, os.path.join(dirpath, f)
