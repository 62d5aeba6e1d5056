"""CPU tests: pin the oracle against the reference's own known answers (tests/golden/reference_kats.json,
each with its reference file:line) and against independent scipy implementations."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("precision", [32, 64, 0])
def test_correlate_kat(orc, kats, precision):
    k = kats["correlate_valid"]                      # audio_matcher.rs:489-517
    w = np.arange(*k["within_range"], dtype=np.float32)
    got = orc.correlate(w, np.array(k["sample"], dtype=np.float32), orc.MODE_VALID, precision)
    assert got.shape == (18,)
    assert np.abs(got - np.array(k["expect"])).max() < k["abs_tol"]


def test_find_peaks_kat(orc, kats):
    k = kats["find_peaks"]                           # audio_matcher.rs:167-185
    pk = orc.find_peaks(k["y"], k["min_prominence"])
    assert [p.start for p in pk] == k["expect_starts_in_order"]          # descending height
    assert np.allclose([p.prominence for p in pk], k["expect_prominences"], atol=k["abs_tol"])
    assert [p.end - p.start for p in pk] == [1, 1, 1]


def test_is_overshadowed_kat(orc, kats):
    k = kats["is_overshadowed"]                      # audio_matcher.rs:187-218
    p1, p2, p3 = orc.find_peaks(kats["find_peaks"]["y"], 0.0)
    named = {"p1": p1, "p2": p2, "p3": p3, None: None}
    for e, o, dist, expect in k["cases"]:
        assert orc.is_overshadowed(named[e], named[o], k["sr"], float(dist)) is expect, (e, o, dist)


def test_bench_shapes(orc, kats):
    k = kats["bench_shapes"]                         # benches/my_benchmark.rs:29-79
    s = np.arange(*k["sample_range"], dtype=np.float32)
    w = np.arange(*k["within_range"], dtype=np.float32)
    c = orc.correlate(w, s, orc.MODE_VALID, 0)           # direct sums: exact for these integers
    assert c.size == k["valid_len"] and c[0] == k["first_value"]
    assert abs(orc.correlate(w, s, orc.MODE_VALID, 64)[0] / k["first_value"] - 1) < 1e-12
    assert abs(1.0 / orc.inv_autocorr(s, exact=True) - k["sum_squares"]) < 1e-6
    assert abs(1.0 / orc.inv_autocorr(s) / k["sum_squares"] - 1) < 1e-5   # f32 FFT like the reference


@pytest.mark.parametrize("n,m", [(20, 3), (50, 4), (257, 31), (1000, 64)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_correlate_modes_vs_scipy(orc, n, m, mode):
    from scipy import signal
    rng = np.random.default_rng(n * 131 + m)
    w, s = rng.standard_normal(n).astype(np.float32), rng.standard_normal(m).astype(np.float32)
    ref = signal.correlate(w.astype(np.float64), s.astype(np.float64), mode=["full", "same", "valid"][mode], method="direct")
    for precision, tol in ((0, 1e-12), (64, 1e-11), (32, 2e-5)):
        got = orc.correlate(w, s, mode, precision)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max())


def test_correlate_degenerate(orc):
    assert orc.out_len(5, 9, orc.MODE_VALID) == 0     # window shorter than the snippet: no outputs
    assert orc.out_len(0, 3, orc.MODE_FULL) == 0
    assert orc.correlate(np.ones(4, np.float32), np.ones(9, np.float32), orc.MODE_VALID).size == 0


def test_find_peaks_vs_scipy(orc):
    from scipy import signal
    rng = np.random.default_rng(7)
    y = rng.standard_normal(20000).astype(np.float32)
    pk = orc.find_peaks(y, 0.8)
    sp, props = signal.find_peaks(y, prominence=0.8)
    assert sorted(p.start for p in pk) == list(sp)
    byp = {p.start: p for p in pk}
    assert np.allclose([byp[i].prominence for i in sp], props["prominences"], atol=1e-6)
    assert all(a.height >= b.height for a, b in zip(pk, pk[1:]))         # ordered by height


def test_find_peaks_plateaus_and_edges(orc):
    y = [0, 1, 1, 1, 0, 2, 2, 3, 3, 0, 5]
    pk = orc.find_peaks(y, 0.0)
    assert sorted((p.start, p.end) for p in pk) == [(1, 4), (7, 9)]      # plateau ranges; rising plateau and the end are not peaks
    assert orc.find_peaks([1, 1, 1, 1], 0.0) == [] and orc.find_peaks([3, 2], 0.0) == []
    assert orc.find_peaks([0, 0, 0, 0, 0], 0.0) == []                    # silence -> flat correlation -> no peaks


def test_find_peaks_min_distance(orc):
    y = np.zeros(100, np.float32)
    y[[10, 14, 30, 33, 60]] = [1.0, 2.0, 3.0, 2.5, 0.5]
    assert sorted(p.start for p in orc.find_peaks(y, 0.0, 0)) == [10, 14, 30, 33, 60]
    assert sorted(p.start for p in orc.find_peaks(y, 0.0, 5)) == [14, 30, 60]       # greedy by height, strict <
    assert sorted(p.start for p in orc.find_peaks(y, 0.0, 4)) == [10, 14, 30, 60]   # distance 4 is not < 4
    assert sorted(p.start for p in orc.find_peaks(y, 0.0, 17)) == [10, 30, 60]  # 14 is within 17 of 30, 10 is not


def test_pcm_scale(orc):
    pcm = np.array([[32767, -32768], [100, 50], [-1, 1]], dtype=np.int16)
    got = orc.pcm16_to_f32(pcm, 2)                   # mp3_reader.rs:12,35
    f = np.float32(1.0) / np.float32(65535.0)
    exp = [(np.float32(l) + np.float32(r)) * np.float32(0.5) * f for l, r in pcm]
    assert np.array_equal(got, np.array(exp, dtype=np.float32))
    mono = orc.pcm16_to_f32(np.array([8191, -8192], np.int16), 1)
    assert np.array_equal(mono, np.array([np.float32(8191) * f, np.float32(-8192) * f], np.float32))


def test_synth_is_deterministic_and_bounded(orc):
    a = orc.synth_pcm16(orc.SEED_STREAM, 1000, 5000)
    b = orc.synth_pcm16(orc.SEED_STREAM, 0, 6000)[1000:]
    assert np.array_equal(a, b) and a.min() >= -8192 and a.max() <= 8191
    assert orc.lib().orc_hash64(0, 0) == 0 and orc.lib().orc_hash64(1, 2) == orc.lib().orc_hash64(1, 2)


def test_calc_chunks_quirks(orc):
    """Chunk geometry quirks of audio_matcher.rs:99-131: gap when ov < m - 1, duplicates when ov > m,
    an occurrence exactly on a chunk boundary is an array endpoint and therefore never a peak."""
    with open(os.path.join(GOLD, "oracle_cases.json")) as f:
        cases = {c["name"]: c for c in json.load(f)}
    assert 160000 not in [p[0] for p in cases["mono_8k"]["peaks"]]       # 4C: endpoint of chunks 3 and 4
    assert 160000 in [p[0] for p in cases["dup_overlap_long"]["peaks"]]  # ov > m: inside chunk 3's window


def test_golden_oracle_cases_reproduce(orc):
    with open(os.path.join(GOLD, "oracle_cases.json")) as f:
        cases = json.load(f)
    for c in cases[:2]:
        pcm, snip, planted = orc.synth_case(c["sr"], c["stream_s"], c["snippet_s"], channels=c["channels"],
                                            chunk_s=c["chunk_s"], plant_period_s=c["chunk_s"] * 2.5,
                                            plant_jitter_s=c["chunk_s"] / 2)
        assert int(pcm.astype(np.int64).sum()) == c["pcm_checksum"]
        x, s = orc.pcm16_to_f32(pcm, c["channels"]), orc.pcm16_to_f32(snip, 1)
        cfg = orc.make_config(c["chunk_s"], len(s) / c["sr"] if c["overlap_s"] is None else c["overlap_s"],
                              c["distance_s"], c["prominence"])
        for precision in (64, 32):
            got = orc.calc_chunks(x, s, c["sr"], cfg, scale=True, precision=precision)
            assert [p.start for p in got] == [p[0] for p in c["peaks"]]
            assert np.allclose([p.height for p in got], [p[2] for p in c["peaks"]], rtol=1e-4)


def test_optimised_cpu_variant_matches(orc):
    """precision 33 (power-of-two transforms, cached snippet spectrum: the 'optimised CPU' baseline line) must find
    the same peaks as the exact-length restatement."""
    pcm, snip, _ = orc.synth_case(8000, 47.0, 1.0, chunk_s=5.0, plant_period_s=12.5, plant_jitter_s=2.5)
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    cfg = orc.make_config(5.0, 1.0, 2.0, 0.13)
    a = orc.calc_chunks(x, s, 8000, cfg, precision=32)
    b = orc.calc_chunks(x, s, 8000, cfg, precision=33)
    assert len(a) >= 3 and [p.start for p in a] == [p.start for p in b]
    assert np.allclose([p.height for p in a], [p.height for p in b], rtol=1e-4)


def test_sharded_oracle_equals_whole(orc):
    pcm, snip, _ = orc.synth_case(8000, 60.0, 0.5, chunk_s=5.0, plant_period_s=12.5, plant_jitter_s=2.5)
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    cfg = orc.make_config(5.0, 0.5, 2.0, 0.13)
    whole = orc.calc_chunks(x, s, 8000, cfg)
    n = orc.num_chunks(len(x), 8000, cfg)
    parts = []
    for r in range(3):
        a, b = r * n // 3, (r + 1) * n // 3
        parts += orc.calc_chunks(x, s, 8000, cfg, first_chunk=a, n_chunks=b - a, final_filter=False)
    merged = orc.merge_peaks(parts, 8000, 2.0)
    assert [(p.start, p.height) for p in merged] == [(p.start, p.height) for p in whole]
