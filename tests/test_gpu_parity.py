"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs, against the committed golden vectors, and -- at BASELINE.json's full sizes --
through size-independent properties.  Offsets must be bit-exact; scores within 1e-4 relative (fp32)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-4          # north_star: correlation scores within 1e-4 relative in fp32


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(float(np.abs(b).max()), 1e-30))


def test_reference_kats_on_gpu(am):
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        kats = json.load(f)
    k = kats["correlate_valid"]                                            # audio_matcher.rs:489-517
    for no_direct in ("", "1"):                                            # direct path and FFT path
        os.environ["AM_NO_DIRECT"] = no_direct
        if not no_direct:
            os.environ.pop("AM_NO_DIRECT")
        algo = am.CudaConvolve(np.array(k["sample"], np.float32), sr=1)
        got = algo.correlate_with_sample(am.test_data(range(*k["within_range"])), am.Mode.Valid, False)
        assert np.abs(got - np.array(k["expect"])).max() < k["abs_tol"]
        assert abs(algo.inverse_sample_auto_correlation() - 1 / 14) < 1e-8
        algo.close()
    os.environ.pop("AM_NO_DIRECT", None)
    b = kats["bench_shapes"]                                               # benches/my_benchmark.rs:29-79
    algo = am.CudaConvolve(am.test_data(range(*b["sample_range"])), sr=1)
    c = algo.correlate_with_sample(am.test_data(range(*b["within_range"])), am.Mode.Valid, False)
    assert c.size == b["valid_len"] and abs(c[0] / b["first_value"] - 1) < 1e-6
    assert abs(1 / algo.inverse_sample_auto_correlation() / b["sum_squares"] - 1) < 1e-6
    algo.close()


def test_find_peaks_kat_through_calc_chunks(am, orc):
    """audio_matcher.rs:167-185 pushed through the whole device path: a 1-tap snippet makes the
    correlation equal to the stream, one chunk covers it."""
    y = np.array([0, 0.7, 0.5, 1.0, 0.5, 0.8, 0.0], np.float32)
    conf = am.Config(chunk_size=7.0, overlap_length=0.0, peak_config=am.PeakConfig(0.0, 0.0))
    algo = am.CudaConvolve(np.ones(1, np.float32), sr=1, config=conf)
    got = am.calc_chunks(1, y, algo, True, conf)
    algo.close()
    assert [p.position.start for p in got] == [1, 3, 5]                    # calc_chunks sorts by start
    assert np.allclose([p.prominence for p in got], [0.2, 1.0, 0.3], atol=1e-6)
    assert np.allclose([p.height for p in got], [0.7, 1.0, 0.8], atol=1e-7)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n,m,log2", [(4000, 50, 0), (4000, 50, 9), (20000, 777, 0), (25810, 4096, 14), (99538, 16384, 16),
                                      (1574098, 262144, 20), (6292690, 480000, 22), (9000000, 1323000, 23)])
def test_correlate_vs_oracle(am, orc, n, m, log2, mode):
    rng = np.random.default_rng(n + m)
    w, s = rng.standard_normal(n).astype(np.float32), rng.standard_normal(m).astype(np.float32)
    if log2 == 9:
        os.environ["AM_NO_DIRECT"] = "1"
    try:
        algo = am.CudaConvolve(s, sr=1000, config=am.Config(fft_log2=log2))
        got = algo.correlate_with_sample(w, am.Mode(mode), False)
        st = algo.stats()
        algo.close()
    finally:
        os.environ.pop("AM_NO_DIRECT", None)
    if n * m <= 4e6:
        ref = orc.correlate(w, s, mode, 0)
    else:
        from scipy import signal
        ref = signal.correlate(w.astype(np.float64), s.astype(np.float64), mode=["full", "same", "valid"][mode], method="fft")
    assert got.shape == ref.shape and st["kernel_launches"] > 0
    assert _rel(got, ref) < 5e-6


def test_correlate_formats_and_scale(am, orc):
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32767, size=(30000, 2), dtype=np.int16)
    snip = rng.integers(-20000, 20000, size=900, dtype=np.int16)
    x, s = orc.pcm16_to_f32(pcm, 2), orc.pcm16_to_f32(snip, 1)
    ref = orc.correlate(x, s, orc.MODE_VALID, 0) * orc.inv_autocorr(s, exact=True)
    a_pcm = am.CudaConvolve(snip, sr=8000)
    a_f32 = am.CudaConvolve(s, sr=8000)
    got_pcm = a_pcm.correlate_with_sample(pcm, am.Mode.Valid, True)        # int16 stereo, downmix fused into the load
    got_f32 = a_f32.correlate_with_sample(x, am.Mode.Valid, True)
    assert np.array_equal(got_pcm, got_f32)                                # identical samples after the fused conversion
    assert _rel(got_pcm, ref) < 5e-6
    unscaled = a_f32.correlate_with_sample(x, am.Mode.Valid, False)
    a_f32.scale(unscaled)                                                  # CorrelateAlgo::scale, :73-75
    assert _rel(unscaled, ref) < 5e-6
    import torch
    got_dev = a_pcm.correlate_with_sample(torch.from_numpy(pcm).cuda(), am.Mode.Valid, True)
    assert np.array_equal(got_dev, got_pcm)
    a_pcm.close(); a_f32.close()


def _run_case(am, orc, c, device=False, fft_log2=0):
    pcm, snip, _ = orc.synth_case(c["sr"], c["stream_s"], c["snippet_s"], channels=c["channels"], chunk_s=c["chunk_s"],
                                  plant_period_s=c["chunk_s"] * 2.5, plant_jitter_s=c["chunk_s"] / 2)
    assert int(pcm.astype(np.int64).sum()) == c["pcm_checksum"]
    conf = am.Config(chunk_size=c["chunk_s"], overlap_length=-1.0 if c["overlap_s"] is None else c["overlap_s"],
                     peak_config=am.PeakConfig(c["distance_s"], c["prominence"]), fft_log2=fft_log2)
    algo = am.CudaConvolve(snip, sr=c["sr"], config=conf)
    stream = pcm.reshape(-1, 2) if c["channels"] == 2 else pcm
    if device:
        import torch
        stream = torch.from_numpy(stream).cuda()
    got = am.calc_chunks(c["sr"], stream, algo, True, conf)
    algo.close()
    return got


def _assert_peaks(got, ref_rows):
    assert [p.position.start for p in got] == [r[0] for r in ref_rows]     # bit-exact offsets after suppression
    assert [p.position.stop for p in got] == [r[1] for r in ref_rows]
    for p, r in zip(got, ref_rows):
        assert abs(p.height - r[2]) <= REL_TOL * abs(r[2])
        assert abs(p.prominence - r[3]) <= REL_TOL * abs(r[3])
        assert p.chunk == r[4]


@pytest.mark.parametrize("device", [False, True])
def test_golden_cases(am, orc, device):
    with open(os.path.join(GOLD, "oracle_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        _assert_peaks(_run_case(am, orc, c, device=device), c["peaks"])


@pytest.mark.parametrize("fft_log2", [13, 14, 17])
def test_golden_case_other_block_lengths(am, orc, fft_log2):
    with open(os.path.join(GOLD, "oracle_cases.json")) as f:
        c = json.load(f)[0]
    _assert_peaks(_run_case(am, orc, c, fft_log2=fft_log2), c["peaks"])


@pytest.mark.parametrize("row_log2", [13, 12, 11])
def test_summary_mode_small_cases(am, orc, row_log2):
    """Summary mode (run records instead of the dense correlation) needs 16-column tiles, i.e. blocks of >= 2^20:
    force that block length on the small golden cases (column lengths 128 / 256 / 512) and check the mode was used."""
    with open(os.path.join(GOLD, "oracle_cases.json")) as f:
        cases = json.load(f)
    os.environ["AM_ROW_LOG2"] = str(row_log2)
    try:
        for c in cases:
            pcm, snip, _ = orc.synth_case(c["sr"], c["stream_s"], c["snippet_s"], channels=c["channels"], chunk_s=c["chunk_s"],
                                          plant_period_s=c["chunk_s"] * 2.5, plant_jitter_s=c["chunk_s"] / 2)
            conf = am.Config(chunk_size=c["chunk_s"], overlap_length=-1.0 if c["overlap_s"] is None else c["overlap_s"],
                             peak_config=am.PeakConfig(c["distance_s"], c["prominence"]), fft_log2=20)
            algo = am.CudaConvolve(snip, sr=c["sr"], config=conf)
            got = am.calc_chunks(c["sr"], pcm.reshape(-1, 2) if c["channels"] == 2 else pcm, algo, True, conf)
            st = algo.stats()
            algo.close()
            _assert_peaks(got, c["peaks"])
            # short snippets give noisy scores: a chunk minimum below theta - prominence legitimately forces the
            # dense repeat (mode 2); the 1 s snippet case must stay in summary mode
            assert st["summary_mode"] in (1, 2), (c["name"], st)
            if c["name"] == "mono_16k_1s":
                assert st["summary_mode"] == 1, st
    finally:
        os.environ.pop("AM_ROW_LOG2", None)


def test_summary_mode_falls_back_to_dense(am, orc):
    """Geometries / data the run records cannot represent exactly must be done densely, not approximated:
    (a) full windows whose last run holds more than one valid output (ov = m + 4): no summary pass at all,
    (b) a chunk minimum below theta - prominence (stream much louder than the snippet): only the chunks whose kept
        peaks do not cover them are repeated densely (am_stats.dense_chunks)."""
    sr, m = 8000, 4000
    pcm = orc.synth_pcm16(41, 0, sr * 43)
    snip = orc.synth_pcm16(42, 0, m)
    for k, o in enumerate([9000, 120000, 260000]):
        orc.synth_plant(pcm, 1, snip, o, k % 2)
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    ov_s = (m + 4) / sr
    ref = orc.calc_chunks(x, s, sr, orc.make_config(5.0, ov_s, 2.0, 0.13), scale=True, precision=64)
    conf = am.Config(chunk_size=5.0, overlap_length=ov_s, peak_config=am.PeakConfig(2.0, 0.13), fft_log2=20)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    got = am.calc_chunks(sr, pcm, algo, True, conf)
    assert algo.stats()["summary_mode"] == 0
    _assert_peaks(got, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])
    algo.close()
    quiet = (snip // 4).astype(np.int16)                                   # snippet 12 dB below the stream: noisy scores
    sq = orc.pcm16_to_f32(quiet)
    for dist, pkcap in ((2.0, 0), (2.0, 48), (0.5, 200), (0.0, 0)):        # small caps force the descent through many height bands
        ref = orc.calc_chunks(x, sq, sr, orc.make_config(2.5, m / sr, dist, 0.3), scale=True, precision=64, cap=1 << 18)
        conf = am.Config(chunk_size=2.5, peak_config=am.PeakConfig(dist, 0.3), fft_log2=20, max_peaks_per_chunk=pkcap)
        algo = am.CudaConvolve(quiet, sr=sr, config=conf)
        got = am.calc_chunks(sr, pcm, algo, True, conf, cap=1 << 18)
        st = algo.stats()
        assert st["summary_mode"] in (1, 2) and (st["summary_mode"] == 2) == (st["dense_chunks"] > 0), st
        assert st["dense_chunks"] <= st["chunks"]
        assert len(ref) > 3
        _assert_peaks(got, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])
        algo.close()


def _coloured_noise(rng, n, rho, rms):
    """AR(1) noise x[i] = rho x[i-1] + e[i], scaled to `rms` (int16 PCM)."""
    from scipy import signal
    e = rng.standard_normal(n)
    x = signal.lfilter([1.0], [1.0, -rho], e)
    x *= rms / np.sqrt(np.mean(x * x))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("kind", ["loud_white", "coloured", "tonal"])
def test_loud_and_coloured_streams_vs_oracle(am, orc, kind):
    """Programme material is rarely white noise at the snippet's level: scores are normalised by the snippet's energy
    only (audio_matcher.rs:321-329), so a stream 12 dB louder, an AR(1) (coloured) spectrum or a tonal stream push the
    chunk minimum far below theta - prominence and fill every chunk with local maxima.  Results must still equal the
    oracle's, through summary mode (block 2^20) and through the dense path (block 2^15), without raising
    max_peaks_per_chunk."""
    rng = np.random.default_rng({"loud_white": 11, "coloured": 12, "tonal": 13}[kind])
    sr, m, n = 8000, 4000, 8000 * 42 + 123
    if kind == "loud_white":
        snip = _coloured_noise(rng, m, 0.0, 1500.0)
        pcm = _coloured_noise(rng, n, 0.0, 6000.0)                          # stream RMS = 4 x snippet RMS
    elif kind == "coloured":
        snip = _coloured_noise(rng, m, 0.97, 2500.0)
        pcm = _coloured_noise(rng, n, 0.97, 7000.0)
    else:
        t = np.arange(n)
        snip = np.round(3000 * np.sin(2 * np.pi * 440.0 / sr * np.arange(m)) + 300 * rng.standard_normal(m)).astype(np.int16)
        pcm = np.round(6000 * np.sin(2 * np.pi * 440.0 / sr * t + 0.3) + 2000 * np.sin(2 * np.pi * 97.0 / sr * t)
                       + 500 * rng.standard_normal(n)).astype(np.int16)
    for k, o in enumerate([11003, 100000, 170500, 290001]):                # real occurrences on top of the programme
        seg = pcm[o:o + m].astype(np.int32) // 2 + snip.astype(np.int32) * (2 if k % 2 == 0 else 1)
        pcm[o:o + m] = np.clip(seg, -32768, 32767).astype(np.int16)
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    for dist, prom in ((480.0, 0.13), (1.0, 0.13), (3.0, 0.5)):
        ref = orc.calc_chunks(x, s, sr, orc.make_config(5.0, m / sr, dist, prom), scale=True, precision=64, cap=1 << 18)
        assert len(ref) >= 2                                                # (the global neighbour filter thins them at distance 480 s)
        for log2 in (20, 15):
            conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(dist, prom), fft_log2=log2)
            algo = am.CudaConvolve(snip, sr=sr, config=conf)
            got = am.calc_chunks(sr, pcm, algo, True, conf, cap=1 << 18)
            st = algo.stats()
            algo.close()
            assert (st["summary_mode"] != 0) == (log2 == 20), st
            _assert_peaks(got, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


@pytest.mark.parametrize("seed,dist,prom,maxpk", [(1, 0.0, 0.09, 0), (2, 1.0, 0.09, 32), (3, 3.0, 0.11, 0), (4, 480.0, 0.13, 8)])
def test_calc_chunks_random_vs_oracle(am, orc, seed, dist, prom, maxpk):
    sr = 8000
    pcm = orc.synth_pcm16(1000 + seed, 0, sr * 33 + 17 * seed)             # ragged tail window
    snip = orc.synth_pcm16(2000 + seed, 0, 3000 + 100 * seed)
    for k, o in enumerate([5000, 70001, 140000, 199999]):
        orc.synth_plant(pcm, 1, snip, o, k % 3)
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    ref = orc.calc_chunks(x, s, sr, orc.make_config(5.0, len(s) / sr, dist, prom), scale=True, precision=64, cap=1 << 18)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(dist, prom), max_peaks_per_chunk=maxpk)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    got = am.calc_chunks(sr, pcm, algo, True, conf, cap=1 << 18)
    algo.close()
    assert len(ref) > 0
    _assert_peaks(got, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


def test_unaligned_device_buffer_and_dc_offset(am, orc):
    """The 128-bit staged PCM loads need 16-byte aligned rows; a device view that starts at an odd sample must take
    the scalar load path and give the same peaks.  A strong DC offset in the stream must not hurt the scores."""
    import torch
    sr, m = 8000, 4000
    pcm = orc.synth_pcm16(31, 0, sr * 37 + 1)
    snip = orc.synth_pcm16(32, 0, m)
    for k, o in enumerate([7001, 90000, 170003, 250000]):
        orc.synth_plant(pcm, 1, snip, o, k % 3)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13), fft_log2=15)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    dev = torch.from_numpy(pcm).cuda()
    for view, host in ((dev[1:], pcm[1:]), (dev[:-1], pcm[:-1])):            # odd start / aligned start
        got = am.calc_chunks(sr, view, algo, True, conf)
        x, s = orc.pcm16_to_f32(np.ascontiguousarray(host)), orc.pcm16_to_f32(snip)
        ref = orc.calc_chunks(x, s, sr, orc.make_config(5.0, m / sr, 2.0, 0.13), scale=True, precision=64)
        assert len(ref) >= 3
        _assert_peaks(got, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])
    shifted = (pcm.astype(np.int32) // 2 + 12000).astype(np.int16)             # large DC component
    got = am.calc_chunks(sr, shifted, algo, True, conf)
    x = orc.pcm16_to_f32(shifted)
    ref = orc.calc_chunks(x, orc.pcm16_to_f32(snip), sr, orc.make_config(5.0, m / sr, 2.0, 0.13), scale=True, precision=64)
    assert [p.position.start for p in got] == [p.start for p in ref] and len(ref) >= 1
    for a, b in zip(got, ref):
        assert abs(a.height - b.height) <= REL_TOL * abs(b.height)
    algo.close()


def test_batch_equals_independent_runs(am, orc):
    """BASELINE config 3 semantics: a batch of snippets == one reference-semantics run per snippet."""
    sr, m = 8000, 3000
    pcm = orc.synth_pcm16(77, 0, sr * 41)
    snips = [orc.synth_pcm16(500 + i, 0, m) for i in range(3)]
    for k, o in enumerate([9000, 100123, 200500, 280000]):
        orc.synth_plant(pcm, 1, snips[k % 3], o, k % 2)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13), fft_log2=14)
    batch = am.CudaConvolve(np.stack([orc.pcm16_to_f32(s) for s in snips]), sr=sr, config=conf, batch=True)
    got = am.calc_chunks(sr, pcm, batch, True, conf)
    assert batch.n_snippets == 3 and [p.snippet_id for p in got] == sorted(p.snippet_id for p in got)
    x = orc.pcm16_to_f32(pcm)
    for i, s in enumerate(snips):
        sf = orc.pcm16_to_f32(s)
        ref = orc.calc_chunks(x, sf, sr, orc.make_config(5.0, m / sr, 2.0, 0.13), scale=True, precision=64)
        mine = [p for p in got if p.snippet_id == i]
        assert len(ref) >= 1
        _assert_peaks(mine, [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])
        batch.select_snippet(i)                                            # am_correlate on snippet i of the batch
        c = batch.correlate_with_sample(x[:40000], am.Mode.Valid, True)
        assert _rel(c, orc.correlate(x[:40000], sf, orc.MODE_VALID, 64) * orc.inv_autocorr(sf, exact=True)) < 5e-6
    batch.close()


def _two_chunk_case(orc, sr, snippet_s, channels, seed, plant_at):
    """A stream of two logical 60 s chunks (+ overlap) at one of BASELINE.json's geometries with the snippet planted
    at `plant_at` (frame offsets, gains 1, 1/2, 1/4 cycling)."""
    m, Cs = int(round(snippet_s * sr)), 60 * sr
    frames = 2 * Cs + m
    pcm = orc.synth_pcm16(orc.SEED_STREAM + seed, 0, frames * channels)
    snip = orc.synth_pcm16(orc.SEED_SNIP + seed, 0, m)
    for k, o in enumerate(plant_at):
        orc.synth_plant(pcm, channels, snip, o, k % 3)
    return pcm, snip, m, frames


def test_cfg4_geometry_vs_oracle(am, orc):
    """BASELINE.json configs[3] geometry: 44.1 kHz, 30 s snippet (m = 1,323,000), block 2^23 = 512 x 16384
    (k_row32<14>, k_col_inv<9,4,32,14>, the TMA-fed forward column kernel with 16384-frame rows): calc_chunks over two
    logical chunks against the oracle (audio_matcher.rs:88-141, 221-230)."""
    import torch
    sr = 44100
    pcm, snip, m, frames = _two_chunk_case(orc, sr, 30.0, 1, 4, [700_001, 60 * sr + 1_234_567])
    conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(480.0, 0.13), fft_log2=23)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    got = algo._calc(torch.from_numpy(pcm).cuda(), True, None, 0, 0, None, False, 1 << 12)
    st = algo.stats()
    algo.close()
    assert (st["fft_log2"], st["log2_n1"], st["log2_n2"]) == (23, 9, 14) and st["summary_mode"] == 1
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    ref = orc.calc_chunks(x, s, sr, orc.make_config(60.0, m / sr, 480.0, 0.13), scale=True, precision=64, threads=3,
                          final_filter=False)
    assert len(ref) == 2
    _assert_peaks(sorted(got, key=lambda p: (p.chunk, -p.height)), [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


def test_cfg5_geometry_vs_oracle(am, orc):
    """BASELINE.json configs[4] geometry: stereo 96 kHz, 2 s snippet (m = 192,000), block 2^20 = 128 x 8192, downmix
    (l + r) * 0.5 / 65535 fused into the TMA-fed column loads (mp3_reader.rs:12,35)."""
    import torch
    sr = 96000
    pcm, snip, m, frames = _two_chunk_case(orc, sr, 2.0, 2, 5, [3_000_017, 60 * sr + 2_500_000, 60 * sr - 100_000])
    conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(480.0, 0.13), fft_log2=20)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    got = algo._calc(torch.from_numpy(pcm.reshape(-1, 2)).cuda(), True, None, 0, 0, None, False, 1 << 12)
    st = algo.stats()
    algo.close()
    assert (st["fft_log2"], st["log2_n1"], st["log2_n2"]) == (20, 7, 13)
    x, s = orc.pcm16_to_f32(pcm, 2), orc.pcm16_to_f32(snip)
    ref = orc.calc_chunks(x, s, sr, orc.make_config(60.0, m / sr, 480.0, 0.13), scale=True, precision=64, threads=3,
                          final_filter=False)
    assert len(ref) >= 2
    _assert_peaks(sorted(got, key=lambda p: (p.chunk, -p.height)), [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


def test_batch_of_8_snippets_at_full_block_length(am, orc):
    """BASELINE.json configs[2] at its real geometry: 8 snippets of 10 s at 48 kHz, block 2^22, i.e. the shared forward
    passes + the default persistent inverse-only row kernel (k_row32_stream<13, ROW_INVERSE>) per snippet.  Every
    snippet must find exactly its planted copies; three of them are checked against the oracle score for score."""
    import torch
    sr, m, Cs = 48000, 480000, 60 * 48000
    frames = 2 * Cs + m
    pcm = orc.synth_pcm16(orc.SEED_STREAM + 8, 0, frames)
    snips = [orc.synth_pcm16(orc.SEED_SNIP + 80 + i, 0, m) for i in range(8)]
    planted = {}
    for i in range(8):                                                     # one copy per snippet, spread over both chunks
        o = 200_003 + i * 610_007
        orc.synth_plant(pcm, 1, snips[i], o, i % 2)
        planted[i] = o
    conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(480.0, 0.13), fft_log2=22)
    batch = am.CudaConvolve(np.stack([orc.pcm16_to_f32(s) for s in snips]), sr=sr, config=conf, batch=True)
    got = batch._calc(torch.from_numpy(pcm).cuda(), True, None, 0, 0, None, False, 1 << 12)
    st = batch.stats()
    batch.close()
    assert st["fft_log2"] == 22 and st["log2_n2"] == 13
    for i in range(8):
        assert [p.position.start for p in got if p.snippet_id == i] == [planted[i]], (i, [(p.snippet_id, p.position.start) for p in got])
    x = orc.pcm16_to_f32(pcm)
    for i in (0, 3, 7):
        ref = orc.calc_chunks(x, orc.pcm16_to_f32(snips[i]), sr, orc.make_config(60.0, m / sr, 480.0, 0.13), scale=True,
                              precision=64, threads=3, final_filter=False)
        _assert_peaks([p for p in got if p.snippet_id == i], [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


def test_progress_callback_and_shard_geometry(am, orc):
    """The progress callbacks of audio_matcher.rs:102-117,129 (segment submitted / call finished) and the library's own
    shard geometry (am_shard_frames == the Python helper, including a .5 rounding case)."""
    sr = 8000
    pcm, snip, _ = orc.synth_case(sr, 40.0, 0.5, chunk_s=5.0, plant_period_s=12.5, plant_jitter_s=2.5)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13))
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    seen = []
    algo.set_progress(lambda phase, c0, nc: seen.append((phase, c0, nc)))
    got = am.calc_chunks(sr, pcm, algo, True, conf)
    algo.set_progress(None)
    assert len(got) > 0 and seen[-1] == (1, 0, 8) and sum(nc for ph, c0, nc in seen if ph == 0) == 8
    from audio_matcher_b200.matcher import shard_frames
    for conf2 in (conf, am.Config(chunk_size=2.5, overlap_length=53 / 128, peak_config=am.PeakConfig(2.0, 0.13))):
        algo.set_config(conf2)                                            # 53/128 s * 8000 = 3312.5 exactly: f64::round gives 3313, round-half-even 3312
        C_, ov = algo.chunk_geometry()
        assert (C_, ov) == ((40000, 4000) if conf2 is conf else (20000, 3313))
        for c0, nc in ((0, 3), (2, 2), (5, 100)):
            assert algo.shard_frames(c0, nc, len(pcm)) == shard_frames(c0, min(nc, algo.num_chunks(len(pcm)) - c0), len(pcm), sr, conf2, algo.m)
    algo.close()


@pytest.mark.parametrize("piece", [1152, 100_003, 9_000_000])
def test_push_session_equals_one_shot(am, orc, native, piece):
    """am_stream_begin / push / finish (the lazy sample iterator of mp3_reader.rs:13-66 feeding calc_chunks,
    matcher/mod.rs:71-83): pieces of an MP3 frame (1152 samples), of an odd size and larger than a pinned slot, segments
    forced small (AM_SEGMENT_MB is read per call) so that several segments, both device buffers and the ov-frame halo
    copy are exercised; the claimed length may exceed the true one.  Must equal am_calc_chunks on the whole stream."""
    sr, m = 8000, 4000
    pcm, snip, planted = orc.synth_case(sr, 1503.7, 0.5, chunk_s=5.0, plant_period_s=61.0, plant_jitter_s=7.0)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13))
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    default_segments = am.calc_chunks(sr, pcm, algo, True, conf)
    assert len(default_segments) > 10
    os.environ["AM_SEGMENT_MB"] = "3"                                      # 3 MB of correlation per segment: ~18 chunks
    try:
        whole = am.calc_chunks(sr, pcm, algo, True, conf)                  # same segments => same block tiling => identical bits
        assert [p.position.start for p in whole] == [p.position.start for p in default_segments]
        for claimed in (len(pcm), len(pcm) + 12345):
            got = am.calc_chunks_streamed(sr, (pcm[i:i + piece] for i in range(0, len(pcm), piece)), claimed, algo, True, conf)
            assert [(p.position.start, p.position.stop, p.height, p.prominence, p.chunk) for p in got] == \
                   [(p.position.start, p.position.stop, p.height, p.prominence, p.chunk) for p in whole]
        st = algo.stats()
        assert st["h2d_bytes"] >= 2 * len(pcm) and st["chunks"] == algo.num_chunks(len(pcm))
        # stereo and f32 sessions, a session on a stream shorter than the snippet, misuse
        stereo = np.stack([pcm, pcm], axis=1)
        got = am.calc_chunks_streamed(sr, (stereo[i:i + 70001] for i in range(0, len(stereo), 70001)), len(stereo), algo, True, conf,
                                      fmt=native.FMT_I16_STEREO)
        assert [p.position.start for p in got] == [p.position.start for p in whole]
        x = orc.pcm16_to_f32(pcm)
        got = am.calc_chunks_streamed(sr, (x[i:i + 333333] for i in range(0, len(x), 333333)), len(x), algo, True, conf, fmt=native.FMT_F32_MONO)
        assert [p.position.start for p in got] == [p.position.start for p in whole]
        assert am.calc_chunks_streamed(sr, [pcm[:1000]], 1000, algo, True, conf) == []
        sess = am.StreamSession(algo, 5000)
        with pytest.raises(native.NativeError):
            am.calc_chunks(sr, pcm, algo, True, conf)                       # the handle is busy with the session
        with pytest.raises(native.NativeError):
            sess.push(pcm[:6000])                                          # more than announced
        sess.abort()
        assert [p.position.start for p in am.calc_chunks(sr, pcm, algo, True, conf)] == [p.position.start for p in whole]
    finally:
        os.environ.pop("AM_SEGMENT_MB", None)
        algo.close()


def _peaks_from_correlation(am, native, c, m, C_, ov, prom, dist, summary, maxpk=0):
    """tests-only hook: per-chunk peak kernels on a supplied correlation (sr = 1, so seconds == samples)."""
    conf = am.Config(chunk_size=float(C_), overlap_length=float(ov), peak_config=am.PeakConfig(float(dist), prom),
                     max_peaks_per_chunk=maxpk)
    algo = am.CudaConvolve(np.ones(m, np.float32), sr=1, config=conf)
    c = np.ascontiguousarray(c, np.float32)
    cap = c.size // 2 + 8
    buf = (native.AmPeak * cap)()
    got, mode = C.c_size_t(), C.c_uint32()
    native.check(native.lib().am_debug_peaks_from_correlation(algo._h, c.ctypes.data, c.size, int(summary), buf, cap,
                                                              C.byref(got), C.byref(mode)))
    algo.close()
    return [(buf[i].chunk, buf[i].start, buf[i].end, buf[i].height, buf[i].prominence, buf[i].left_diff, buf[i].right_diff)
            for i in range(got.value)], mode.value


def _oracle_peaks_from_correlation(orc, c, m, C_, ov, prom, dist):
    L = c.size + m - 1
    out = []
    for i in range((L + C_ - 1) // C_):
        n = min(C_ + ov, L - C_ * i)
        if n < m:
            continue
        y = c[C_ * i: C_ * i + n - m + 1]
        for p in orc.find_peaks(y, prom, dist):
            out.append((i, p.start + C_ * i, p.end + C_ * i, p.height, p.prominence, p.left_diff, p.right_diff))
    return out


@pytest.mark.parametrize("summary", [0, 1])
@pytest.mark.parametrize("seed", range(6))
def test_peak_kernels_on_adversarial_correlations(am, orc, native, seed, summary):
    """The peak kernels (dense and run-record paths) against the oracle's find_peaks on arrays no transform would
    produce: heavy quantisation (plateaus inside runs, across run and tile boundaries), peaks next to chunk edges,
    flat and monotone stretches.  Everything must match bit for bit; in summary mode unstored runs are poisoned."""
    rng = np.random.default_rng(100 + seed)
    C_, m = 4096, 37
    ov = [m, m - 1, m + 15, m, m + 31, m][seed]                      # V = C+1, C, C+16, C+1, C+32, C+1 (partial runs of 0/1 outputs)
    n = 5 * C_ + [1, 0, 16, 777, 32, 2049][seed]
    levels = [4, 9, 33, 3, 17, 65][seed]
    c = np.round(rng.random(n) * levels) / levels * 0.2 - 0.02       # quantised noise in [-0.02, 0.18]: many plateaus
    for pos in rng.integers(2, n - 2, size=40):                      # bumps of various widths and heights
        w = int(rng.integers(1, 40))
        c[pos:pos + w] = np.maximum(c[pos:pos + w], rng.choice([0.3, 0.5, 0.5, 0.8, 1.0]))
    c[1] = 0.9; c[C_ - 1] = 0.95; c[C_] = 0.95; c[C_ + 1] = 0.7       # next to array / chunk edges, plateau over a chunk boundary
    c[2 * C_ + 15:2 * C_ + 17] = 0.6                                  # plateau across a run boundary
    c[3 * C_ + 1023:3 * C_ + 1026] = 0.65                             # plateau across a tile boundary
    c[4 * C_:4 * C_ + 300] = np.linspace(0.0, 0.17, 300)              # monotone stretch
    c = c.astype(np.float32)
    # (the quantised background puts hundreds of exactly equal maxima into a chunk, which height bands cannot split: the
    # small-cap rows use a prominence only the bumps reach; real correlations have no such ties)
    for prom, dist, maxpk in ((0.13, 0, 4000), (0.25, 100, 4000), (0.13, 3000, 4000), (0.25, 3000, 16), (0.25, 400, 24)):
        ref = _oracle_peaks_from_correlation(orc, c, m, C_, ov, prom, dist)
        got, mode = _peaks_from_correlation(am, native, c, m, C_, ov, prom, dist, summary, maxpk=maxpk)
        assert mode == (1 if summary else 0)
        assert len(ref) > 5
        assert sorted(got) == sorted(ref)


def test_peak_kernels_summary_rejections(am, orc, native):
    """Run records cannot represent (a) a partial last run with more than one output inside a segment and (b) a chunk
    minimum below theta - prominence: both must come back as mode 2 (dense repeat) with exact results."""
    rng = np.random.default_rng(7)
    C_, m = 4096, 37
    c = (rng.random(3 * C_ + 5) * 0.1).astype(np.float32)
    c[[500, 5000, 9000]] = [1.0, 0.6, 0.8]
    ref = _oracle_peaks_from_correlation(orc, c, m, C_, m + 4, 0.13, 0)
    got, mode = _peaks_from_correlation(am, native, c, m, C_, m + 4, 0.13, 0, 1)           # V = C + 5
    assert mode == 2 and sorted(got) == sorted(ref)
    c2 = c.copy()
    c2[7000] = -0.5                                                                          # chunk minimum far below -prom/2
    ref = _oracle_peaks_from_correlation(orc, c2, m, C_, m, 0.13, 0)
    got, mode = _peaks_from_correlation(am, native, c2, m, C_, m, 0.13, 0, 1, maxpk=4000)
    assert mode == 2 and sorted(got) == sorted(ref)
    # with a minimum distance that one kept peak per chunk satisfies, the same data stays in summary mode: the kept
    # peaks of height >= theta cover their chunks, whatever hides below theta cannot survive
    ref = _oracle_peaks_from_correlation(orc, c2, m, C_, m, 0.13, 2 * C_)
    got, mode = _peaks_from_correlation(am, native, c2, m, C_, m, 0.13, 2 * C_, 1)
    assert mode == 1 and sorted(got) == sorted(ref)


def test_edge_cases(am, orc, native):
    sr = 8000
    snip = orc.synth_pcm16(7, 0, 4000)
    conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13))
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    assert am.calc_chunks(sr, np.zeros(0, np.int16), algo, True, conf) == []                 # empty stream
    assert am.calc_chunks(sr, orc.synth_pcm16(8, 0, 3999), algo, True, conf) == []           # shorter than the snippet
    assert am.calc_chunks(sr, np.zeros(100000, np.int16), algo, True, conf) == []            # silence: flat, no peaks
    exact = snip.copy()                                                                      # stream == snippet: V = 1
    assert am.calc_chunks(sr, exact, algo, True, conf) == []
    assert algo.correlate_with_sample(np.zeros(10, np.float32), am.Mode.Valid).size == 0     # n < m
    one = algo.correlate_with_sample(orc.pcm16_to_f32(snip), am.Mode.Valid, True)            # autocorrelation == 1
    assert one.shape == (1,) and abs(one[0] - 1.0) < 1e-5
    # capacity errors are loud
    noisy = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(0.0, 0.0), max_peaks_per_chunk=16)
    with pytest.raises(native.NativeError) as e:
        am.calc_chunks(sr, orc.synth_pcm16(9, 0, 80000), algo, True, noisy)
    assert e.value.status == native.AM_ERR_CAPACITY
    with pytest.raises(native.NativeError):
        am.calc_chunks(sr, orc.synth_pcm16(9, 0, 80000), algo, True, am.Config(chunk_size=5.0, fft_log2=10))   # 2^10 < 2m
    with pytest.raises(ValueError):
        am.calc_chunks(sr + 1, orc.synth_pcm16(9, 0, 80000), algo, True, conf)               # SampleRateMismatch
    algo.close()


def test_stream_handle_is_shareable_between_threads(am, orc):
    """calc_chunks calls correlate_with_sample concurrently on one shared &algo (audio_matcher.rs:114-122)."""
    import threading
    rng = np.random.default_rng(0)
    s = rng.standard_normal(300).astype(np.float32)
    ws = [rng.standard_normal(5000).astype(np.float32) for _ in range(8)]
    algo = am.CudaConvolve(s, sr=1000)
    out = [None] * 8

    def work(i):
        out[i] = algo.correlate_with_sample(ws[i], am.Mode.Valid, False)
    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(8):
        assert _rel(out[i], orc.correlate(ws[i], s, orc.MODE_VALID, 0)) < 5e-6
    algo.close()


@pytest.fixture(scope="module")
def full_size(am, orc, native):
    """BASELINE.json configs[1]: one 10 s snippet vs a 24 h 48 kHz mono stream, generated on the device."""
    import torch
    sr, m, frames = 48000, 480000, 48000 * 86400
    L = native.lib()
    pcm = torch.empty(frames, dtype=torch.int16, device="cuda")
    native.check(L.am_synth_pcm16_device(orc.SEED_STREAM, 0, frames, pcm.data_ptr(), None))
    snip = orc.synth_pcm16(orc.SEED_SNIP, 0, m)
    sd = torch.from_numpy(snip).cuda()
    P, J, Cs = 600 * sr, 30 * sr, 60 * sr
    planted = {}
    k = 0
    while k * P + m <= frames:
        o = orc.plant_offset(k, P, J, Cs)
        native.check(L.am_synth_plant_device(pcm.data_ptr(), frames, 1, sd.data_ptr(), m, o, k % 4, None))
        planted[o] = k % 4
        k += 1
    torch.cuda.synchronize()
    conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(480.0, 0.13), fft_log2=22)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    yield dict(sr=sr, m=m, frames=frames, pcm=pcm, snip=snip, planted=planted, conf=conf, algo=algo)
    algo.close()


def test_full_size_properties(am, orc, full_size):
    fs = full_size
    algo, conf, sr = fs["algo"], fs["conf"], fs["sr"]
    whole = am.calc_chunks(sr, fs["pcm"], algo, True, conf)
    starts = [p.position.start for p in whole]
    assert len(starts) > 50 and starts == sorted(starts)
    assert all(s in fs["planted"] for s in starts)                         # only planted offsets, exact sample positions
    assert all(fs["planted"][s] <= 2 or p.prominence >= 0.13 for s, p in zip(starts, whole))
    for p in whole:                                                        # score ~ gain of the planted copy
        assert abs(p.height - 2.0 ** -fs["planted"][p.position.start]) < 0.02
    # shard + merge == whole (the multi-GPU decomposition, chunk ranges with halo)
    total = algo.num_chunks(fs["frames"])
    parts = []
    for r in range(3):
        a, b = r * total // 3, (r + 1) * total // 3
        parts += algo._calc(fs["pcm"], True, fs["frames"], 0, a, b - a, False, 1 << 16)
    merged = am.merge_peaks(parts, sr, conf.peak_config.distance)
    assert [p.position.start for p in merged] == starts                    # offsets exact; scores differ only by the
    for a, b in zip(merged, whole):                                        # rounding of a different block tiling
        assert abs(a.height - b.height) <= 1e-5 * abs(b.height) and abs(a.prominence - b.prominence) <= 1e-5 * abs(b.prominence)
    # idempotence / determinism
    again = am.calc_chunks(sr, fs["pcm"], algo, True, conf)
    assert [(p.position.start, p.height) for p in again] == [(p.position.start, p.height) for p in whole]


def test_full_size_chunks_vs_oracle(am, orc, full_size):
    """Two logical chunks of the 24 h workload (block length 2^22, real data) against the oracle."""
    fs = full_size
    sr, m, Cs = fs["sr"], fs["m"], 60 * fs["sr"]
    first, n = 9, 2                                                        # chunks 9-10 hold a planted copy (k = 1, 600 s)
    lo, hi = first * Cs, (first + n) * Cs + m
    host = fs["pcm"][lo:hi].cpu().numpy()
    x, s = orc.pcm16_to_f32(host), orc.pcm16_to_f32(fs["snip"])
    ref = orc.calc_chunks(x, s, sr, orc.make_config(60.0, m / sr, 480.0, 0.13), scale=True, precision=64, n_chunks=n,
                          final_filter=False)
    got = fs["algo"]._calc(fs["pcm"], True, fs["frames"], 0, first, n, False, 1 << 12)
    assert len(ref) >= 1
    assert sorted(p.position.start - lo for p in got) == sorted(p.start for p in ref)
    gh = {p.position.start - lo: p for p in got}
    for r in ref:
        assert abs(gh[r.start].height - r.height) <= REL_TOL * abs(r.height)
        assert abs(gh[r.start].prominence - r.prominence) <= REL_TOL * abs(r.prominence)


def test_host_memory_path_equals_device_path(am, orc, full_size):
    fs = full_size
    import torch
    frames = 48000 * 3600 * 2                                              # 2 h through the staged H2D path
    host = torch.empty(frames, dtype=torch.int16, pin_memory=True)
    host.copy_(fs["pcm"][:frames])
    torch.cuda.synchronize()
    a = am.calc_chunks(fs["sr"], host, fs["algo"], True, fs["conf"])
    st = fs["algo"].stats()
    b = am.calc_chunks(fs["sr"], fs["pcm"][:frames], fs["algo"], True, fs["conf"])
    # same offsets; heights to rounding only: host streams are cut into smaller segments (the unit of the double-buffered
    # upload), so the FFT blocks sit at other positions than on the resident path
    assert [p.position.start for p in a] == [p.position.start for p in b]
    for x, y in zip(a, b):
        assert abs(x.height - y.height) <= 1e-5 * abs(y.height) and abs(x.prominence - y.prominence) <= 1e-5 * abs(y.prominence)
    assert st["h2d_bytes"] >= frames * 2 and len(a) > 0
    # pageable host memory (numpy): goes through the library's pinned ring filled by host threads
    pageable = fs["pcm"][:frames].cpu().numpy()
    c = am.calc_chunks(fs["sr"], pageable, fs["algo"], True, fs["conf"])
    assert [(p.position.start, p.height, p.prominence) for p in c] == [(p.position.start, p.height, p.prominence) for p in a]


def test_device_memory_pressure_shrinks_the_buffers(am, orc, full_size):
    """The workspace (8 GiB) and segment (16 GiB) defaults assume a B200 to ourselves.  With ~3 GiB left on the device a
    new matcher must still run -- smaller launch groups and segments -- and find the same offsets."""
    import torch
    fs = full_size
    sr, frames = fs["sr"], fs["sr"] * 3600 * 6
    ref = am.calc_chunks(sr, fs["pcm"][:frames], fs["algo"], True, fs["conf"])
    launches_unconstrained = fs["algo"].stats()["kernel_launches"]
    algo2 = am.CudaConvolve(fs["snip"], sr=sr, config=fs["conf"])
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    filler = torch.empty(max(free - (3 << 30), 1), dtype=torch.uint8, device="cuda")
    try:
        got = am.calc_chunks(sr, fs["pcm"][:frames], algo2, True, fs["conf"])
        st = algo2.stats()
    finally:
        del filler
        torch.cuda.empty_cache()
        algo2.close()
    assert len(ref) > 10 and [p.position.start for p in got] == [p.position.start for p in ref]
    for a, b in zip(got, ref):                                             # (other segment / block tiling: rounding only)
        assert abs(a.height - b.height) <= 1e-5 * abs(b.height) and abs(a.prominence - b.prominence) <= 1e-5 * abs(b.prominence)
    assert st["summary_mode"] == 1 and st["chunks"] == 360
    assert st["kernel_launches"] > launches_unconstrained                  # more, smaller launch groups / segments


def _peak_tuple(p):
    return (p.position.start, p.position.stop, p.height, p.prominence, p.chunk)


@pytest.mark.parametrize("mem", ["host", "device", "pinned", "mixed"])
def test_calc_chunks_files_equals_one_call_per_file(am, orc, mem):
    """am_calc_chunks_files (the loop over args.within, src/matcher/mod.rs:42-99, as one call): every file's list must be
    exactly what am_calc_chunks returns for that file -- files of different lengths (several segments down to less than
    a chunk), one shorter than the snippet, an empty one, a loud coloured one that needs the dense repeat, and the
    oracle on one of them."""
    import torch
    sr, m = 8000, 4000
    rng = np.random.default_rng(77)
    snip = orc.synth_pcm16(3001, 0, m)
    lens = [sr * 33 + 17, sr * 4, m - 1, 0, sr * 61 + 5, sr * 12, sr * 20 + 3]
    files = []
    for i, n in enumerate(lens):
        pcm = orc.synth_pcm16(4000 + i, 0, n)
        for k, o in enumerate(range(5000 + 1000 * i, max(n - m, 0), 70001)):
            orc.synth_plant(pcm, 1, snip, o, k % 3)
        files.append(pcm)
    files[5] = _coloured_noise(rng, lens[5], 0.97, 9000.0)                 # scores far below theta - prominence
    for log2, chunk_s in ((0, 5.0), (20, 5.0), (15, 2.5)):
        conf = am.Config(chunk_size=chunk_s, peak_config=am.PeakConfig(1.0, 0.11), fft_log2=log2)
        algo = am.CudaConvolve(snip, sr=sr, config=conf)
        if mem == "host":
            arg = files
        elif mem == "device":
            arg = [torch.from_numpy(f).cuda() for f in files]
        elif mem == "mixed":                                                # pageable and pinned files in one call
            arg = [torch.from_numpy(f).pin_memory() if (i % 2 and f.size) else f for i, f in enumerate(files)]
        else:
            arg = [torch.from_numpy(f).pin_memory() if f.size else torch.empty(0, dtype=torch.int16) for f in files]
        single = [[_peak_tuple(p) for p in am.calc_chunks(sr, f, algo, True, conf)] for f in arg]
        multi = am.calc_chunks_files(sr, arg, algo, True, conf)
        st = algo.stats()
        assert [[_peak_tuple(p) for p in l] for l in multi] == single
        assert st["frames"] == sum(n for n in lens if n >= m) and st["kernel_launches"] > 0
        assert len(single[0]) > 0 and len(single[4]) > 0 and single[2] == [] and single[3] == []
        # a file that outgrows its share of the device peak list is repeated through the one-file path
        os.environ["AM_FILES_MIN_CAP"] = "1"
        try:
            again = am.calc_chunks_files(sr, arg, algo, True, conf)
        finally:
            os.environ.pop("AM_FILES_MIN_CAP")
        assert [[_peak_tuple(p) for p in l] for l in again] == single
        # capacity: counts are still reported
        buf_small = sum(len(x) for x in single) - 1
        with pytest.raises(Exception):
            am.calc_chunks_files(sr, arg, algo, True, conf, cap=buf_small)
        # a second call on the same handle (buffers reused) and a plain call afterwards
        assert [[_peak_tuple(p) for p in l] for l in am.calc_chunks_files(sr, arg[:2], algo, True, conf)] == single[:2]
        assert [_peak_tuple(p) for p in am.calc_chunks(sr, arg[0], algo, True, conf)] == single[0]
        algo.close()
    x, s = orc.pcm16_to_f32(files[4]), orc.pcm16_to_f32(snip)
    ref = orc.calc_chunks(x, s, sr, orc.make_config(2.5, m / sr, 1.0, 0.11), scale=True, precision=64, cap=1 << 18)
    _assert_peaks(multi[4], [[p.start, p.end, p.height, p.prominence, p.chunk] for p in ref])


def test_calc_chunks_files_with_a_batch_of_snippets(am, orc):
    """Several files x several snippets (am_matcher_create_batch): per file the snippet-major list of the one-file call."""
    sr, m = 8000, 3000
    snips = [orc.synth_pcm16(500 + i, 0, m) for i in range(3)]
    files = []
    for f, n in enumerate([sr * 41, sr * 7 + 3, sr * 23]):
        pcm = orc.synth_pcm16(80 + f, 0, n)
        for k, o in enumerate(range(9000, n - m, 45001)):
            orc.synth_plant(pcm, 1, snips[(k + f) % 3], o, k % 2)
        files.append(pcm)
    for log2 in (14, 20):
        conf = am.Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13), fft_log2=log2)
        batch = am.CudaConvolve(np.stack([orc.pcm16_to_f32(s) for s in snips]), sr=sr, config=conf, batch=True)
        key = lambda l: [(p.snippet_id,) + _peak_tuple(p) for p in l]
        single = [key(am.calc_chunks(sr, f, batch, True, conf)) for f in files]
        multi = am.calc_chunks_files(sr, files, batch, True, conf)
        assert [key(l) for l in multi] == single and sum(len(l) for l in single) >= 6
        assert all([p.snippet_id for p in l] == sorted(p.snippet_id for p in l) for l in multi)
        batch.close()


def test_calc_chunks_files_at_full_block_length(am, orc, full_size):
    """24 one-hour files (cfg 1-sized pieces of the 24 h stream, N = 2^22) in one call from pinned host memory == one
    call per file, and the offsets found are planted ones."""
    import torch
    fs = full_size
    sr, per = fs["sr"], fs["sr"] * 3600
    host = torch.empty(6 * per, dtype=torch.int16, pin_memory=True)
    host.copy_(fs["pcm"][:6 * per])
    torch.cuda.synchronize()
    files = [host[i * per:(i + 1) * per] for i in range(6)]
    multi = am.calc_chunks_files(sr, files, fs["algo"], True, fs["conf"])
    st = fs["algo"].stats()
    single = [am.calc_chunks(sr, f, fs["algo"], True, fs["conf"]) for f in files]
    assert [[_peak_tuple(p) for p in l] for l in multi] == [[_peak_tuple(p) for p in l] for l in single]
    found = [i * per + p.position.start for i, l in enumerate(multi) for p in l]
    assert len(found) >= 20 and all(o in fs["planted"] for o in found)
    # (a one-hour file is two upload segments; the second re-sends the overlap of the segment boundary)
    assert 6 * per * 2 <= st["h2d_bytes"] <= 6 * (per + 2 * fs["m"]) * 2 and st["chunks"] == 6 * 60


@pytest.mark.parametrize("env", [{"AM_COL_STREAM": "0"}, {"AM_ROW_STREAM": "1"}, {"AM_ROW_STREAM": "0"}, {"AM_BATCH_SNIPPETS": "3"}])
def test_alternate_kernel_paths(env):
    """The kernel choices are read once per process; the non-default ones (plain-grid forward column kernel that also
    serves windows TMA cannot describe, persistent fused row kernel, plain inverse-only row kernel of batch mode) get
    the correlation / golden / batch / full-size-vs-oracle tests in a child process; so does a batch whose snippet count
    is not a multiple of the snippets per inverse-row launch."""
    if os.environ.get("AM_ALT_PATH_CHILD"):
        pytest.skip("child run")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = dict(os.environ, AM_ALT_PATH_CHILD="1", **env)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu",
                          "-k", "correlate_vs_oracle or golden_cases or batch_equals or batch_of_8 or full_size_chunks_vs_oracle"],
                         cwd=root, env=child, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_c_abi_from_plain_c(am, native, tmp_path):
    """The boundary is a C ABI: compile tests/capi_smoke.c with gcc against include/ and the .so, run it."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "capi_smoke"
    libdir = os.path.join(root, "audio_matcher_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "capi_smoke.c"),
                    "-L", libdir, "-laudio_matcher_b200", "-lm", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "capi_smoke ok" in out.stdout, out.stderr


def test_two_gpu_sharded_c_abi(am, tmp_path):
    """am_comm_init + am_calc_chunks_sharded from two plain processes (one per GPU), the NCCL unique id handed over in a
    file: every rank must return the single-GPU result of the whole stream (tests/two_rank_sharded.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = [subprocess.Popen([sys.executable, os.path.join(root, "tests", "two_rank_sharded.py"), str(r), "2", str(tmp_path / "nccl_id")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "sharded ok" in o, f"rank {r}: {o[-2000:]}"


def test_two_gpu_sharded_nccl(am):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(root, "bench.py"), "--gpus", "2", "--hours", "1", "--steps", "1", "--warmup", "1",
           "--no-e2e"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["config"]["verified_offsets_are_planted"]
    assert line["config"]["verified_vs_oracle_chunks"]["all_ranks_ok"]
