"""Pins the CPU oracle (oracle/scan_oracle.cpp):
  (1) against the reference's own gtest known-answer vectors (tests/golden/reference_kats.json), and
  (2) against the reference's own objects compiled from /root/reference (oracle/_ref/libref.so), bit for bit,
      on random inputs — for every hot-path piece that the reference can be built for here.
CPU only (-m "not gpu")."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT

G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
KINDS = ["orc"] + (["ref"] if ol.have_ref() else [])


def _lib(kind):
    return ol.oracle() if kind == "orc" else ol.ref()


@pytest.mark.parametrize("kind", KINDS)
def test_average_kat(kind):
    g = G["average"]
    out = ol.cpu_average(np.array(g["input"], np.float32), g["group"], kind)
    np.testing.assert_allclose(out, np.array(g["expect"], np.float32), rtol=4e-7)  # EXPECT_FLOAT_EQ = 4 ulp


@pytest.mark.parametrize("kind", KINDS)
def test_contains_with_margin_kat(kind):
    g = G["contains_with_margin"]
    keys = np.array(g["keys"], np.int32)
    for index, margin, expect in g["cases"]:
        found = C.c_int()
        got = getattr(_lib(kind), f"{kind}_contains_with_margin")(keys.ctypes.data_as(C.c_void_p), len(keys), index, margin, C.byref(found))
        assert bool(got) == expect, (index, margin)


@pytest.mark.parametrize("kind", KINDS)
def test_most_frequent_value_kat(kind):
    for data, expect in G["most_frequent_value"]["cases"]:
        v = np.array(data, np.int32)
        assert getattr(_lib(kind), f"{kind}_most_frequent_value")(v.ctypes.data_as(C.c_void_p), len(v)) == expect


@pytest.mark.parametrize("kind", KINDS)
def test_get_max_index_kat(kind):
    g = G["get_max_index"]
    d = np.array(g["data"], np.float32)
    for index, group, expect in g["cases"]:
        assert getattr(_lib(kind), f"{kind}_get_max_index")(d.ctypes.data_as(C.c_void_p), len(d), index, group) == expect


@pytest.mark.parametrize("kind", KINDS)
def test_get_fft_and_tuned_kat(kind):
    L = _lib(kind)
    for fs, step, expect in G["get_fft"]["cases"]:
        assert getattr(L, f"{kind}_get_fft")(fs, step) == expect
    for f, step, expect in G["get_tuned_frequency"]["cases"]:
        assert getattr(L, f"{kind}_get_tuned_frequency")(f, step) == expect


def _run_averager_script(av):
    g = G["averager"]
    size, group = g["size"], g["group"]
    for st in g["steps"]:
        if st["op"] == "push":
            av.push(np.full(size, st["v"], np.float32))
        elif st["op"] == "reset":
            av.reset()
        assert np.array_equal(av.average(), np.full(size, st["avg"], np.float32)), st
        assert np.array_equal(av.data(), np.array([[r] * size for r in st["rows"]], np.float32)), st


@pytest.mark.parametrize("kind", KINDS)
def test_averager_kat(kind):
    g = G["averager"]
    _run_averager_script(ol.CpuAverager(g["size"], g["group"], kind))


def _fixture_mean(raw, group):
    s = np.zeros(raw[0].shape, np.float32)
    for r in raw:  # same accumulation order as AveragerTest::average(), tests/test_averager.cpp:29-40
        s = (s + r).astype(np.float32)
    return (s / np.float32(group)).astype(np.float32)


def run_averager_fixture(av, size, group, rows_iter):
    """AveragerTest fixture (tests/test_averager.cpp:13-44): exact equality of average() and data() after every add."""
    raw = [np.zeros(size, np.float32) for _ in range(group)]
    pushed = 0
    for row in rows_iter:
        row = np.asarray(row, np.float32)
        av.push(row)
        raw.append(row)
        raw = raw[-group:]
        pushed += 1
        if pushed < group:
            assert np.array_equal(av.average(), np.full(size, -100, np.float32))
        else:
            assert np.array_equal(av.average(), _fixture_mean(raw, group))
        assert np.array_equal(av.data(), np.stack(raw))


@pytest.mark.parametrize("kind", KINDS)
def test_averager_fixture_kat(kind):
    g = G["averager_fixture"]
    size, group = g["size"], g["group"]
    run_averager_fixture(ol.CpuAverager(size, group, kind), size, group, g["simple"])
    rows = list(g["big"]["prefix"]) + [[i * 11 + j * 7 for j in range(size)] for i in range(*g["big"]["ramp_i"])]
    run_averager_fixture(ol.CpuAverager(size, group, kind), size, group, rows)


# ---------------- oracle == compiled reference, bit for bit, on random data ----------------
@needs_ref
@pytest.mark.parametrize("size,group", [(5, 3), (64, 21), (1000, 21), (4096, 21), (33, 1)])
def test_averager_matches_reference_bitwise(size, group):
    rng = np.random.default_rng(size * 131 + group)
    a, b = ol.CpuAverager(size, group, "orc"), ol.CpuAverager(size, group, "ref")
    for step in range(3 * group + 7):
        if step == 2 * group + 1:
            a.reset(), b.reset()
        row = (rng.standard_normal(size) * 30 - 20).astype(np.float32)
        a.push(row), b.push(row)
        assert a.average().tobytes() == b.average().tobytes()
        assert a.data().tobytes() == b.data().tobytes()


@needs_ref
@pytest.mark.parametrize("size,group", [(9, 5), (100, 21), (4096, 21), (16384, 21), (50, 1), (10, 21), (7, 4)])
def test_average_matches_reference_bitwise(size, group):
    rng = np.random.default_rng(size + group)
    for scale in (1.0, 40.0):
        x = (rng.standard_normal(size) * scale - 7).astype(np.float32)
        assert ol.cpu_average(x, group, "orc").tobytes() == ol.cpu_average(x, group, "ref").tobytes()
    x = np.full(size, -100, np.float32)
    assert ol.cpu_average(x, group, "orc").tobytes() == ol.cpu_average(x, group, "ref").tobytes()


@needs_ref
def test_collection_utils_match_reference_random():
    rng = np.random.default_rng(7)
    O, R = ol.oracle(), ol.ref()
    for _ in range(300):
        n = int(rng.integers(1, 200))
        d = rng.integers(-5, 6, n).astype(np.float32)  # many ties: first-maximum rule matters
        idx, grp = int(rng.integers(0, n)), int(rng.integers(0, 64))
        p = d.ctypes.data_as(C.c_void_p)
        assert O.orc_get_max_index(p, n, idx, grp) == R.ref_get_max_index(p, n, idx, grp)
        v = rng.integers(0, 8, int(rng.integers(1, 30))).astype(np.int32)
        pv = v.ctypes.data_as(C.c_void_p)
        assert O.orc_most_frequent_value(pv, len(v)) == R.ref_most_frequent_value(pv, len(v))
        keys = np.unique(rng.integers(0, 500, int(rng.integers(0, 12)))).astype(np.int32)
        pk = keys.ctypes.data_as(C.c_void_p)
        fo, fr = C.c_int(-1), C.c_int(-1)
        index, margin = int(rng.integers(-10, 510)), int(rng.integers(0, 140))
        ro = O.orc_contains_with_margin(pk, len(keys), index, margin, C.byref(fo))
        rr = R.ref_contains_with_margin(pk, len(keys), index, margin, C.byref(fr))
        assert ro == rr and (not ro or fo.value == fr.value)


@needs_ref
def test_radio_utils_match_reference_random():
    rng = np.random.default_rng(11)
    O, R = ol.oracle(), ol.ref()
    for _ in range(2000):
        f, step = int(rng.integers(-2_000_000, 2_000_000)), int(rng.integers(1, 100_000))
        assert O.orc_get_tuned_frequency(f, step) == R.ref_get_tuned_frequency(f, step)
    for fs in (250_000, 1_024_000, 2_048_000, 2_400_000, 10_000_000, 20_000_000, 40_000_000, 61_440_000):
        for step in (100, 250, 625, 1000):
            assert O.orc_get_fft(fs, step) == R.ref_get_fft(fs, step)


@needs_ref
def test_reference_only_kats():
    """Vectors for reference helpers that sit beside the hot path (range splitting, resampler factors)."""
    R = ol.ref()
    for fs, expect in G["range_split_sample_rate"]["cases"]:
        assert R.ref_get_range_split_sample_rate(fs) == expect
    buf = np.zeros(64, np.int32)
    for lo, hi, fs, expect in G["split_range"]["cases"]:
        n = R.ref_split_range(lo, hi, fs, buf.ctypes.data_as(C.c_void_p), 32)
        assert buf[: 2 * n].reshape(-1, 2).tolist() == expect
    for fs, bw, thr, expect in G["resamplers"]["cases"]:
        n = R.ref_get_resamplers_factors(fs, bw, thr, buf.ctypes.data_as(C.c_void_p), 32)
        assert buf[: 2 * n].reshape(-1, 2).tolist() == expect
    for v, f, e in G["round"]["up"]:
        assert R.ref_round_up(v, f) == e
    for v, f, e in G["round"]["down"]:
        assert R.ref_round_down(v, f) == e
