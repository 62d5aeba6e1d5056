"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, host-only
entry points (merge, overshadow, lengths, defaults) agree with the oracle, compute entry points fail
loudly without a GPU, the Python mirror's helpers behave like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native):
    hdr = open(os.path.join(ROOT, "include", "audio_matcher.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(str(native.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in audio_matcher.h but not exported"
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    assert native.lib().am_abi_version() == native.ABI_VERSION


def test_struct_layouts(native):
    assert C.sizeof(native.AmPeak) == 40 and C.sizeof(native.AmConfig) == 40
    cfg = native.AmConfig()
    native.lib().am_config_default(C.byref(cfg))
    assert (cfg.chunk_size_s, cfg.distance_s, cfg.overlap_s) == (60.0, 480.0, -1.0)      # args.rs:70-76
    assert abs(cfg.prominence - 0.13) < 1e-7 and cfg.fft_log2 == 0


def test_out_len(native, orc):
    L = native.lib()
    for n, m in [(20, 3), (4000, 50), (3, 3), (2, 3), (0, 3), (5, 0)]:
        for mode in (0, 1, 2):
            assert L.am_out_len(n, m, mode) == orc.out_len(n, m, mode)
        assert L.am_valid_len(n, m) == orc.out_len(n, m, 2)


def test_no_gpu_fails_loudly(native):
    if native.lib().am_device_count() > 0:
        pytest.skip("a GPU is visible")
    import audio_matcher_b200 as am
    with pytest.raises(native.NativeError) as e:
        am.CudaConvolve(np.ones(8, np.float32), sr=8000)
    assert e.value.status == native.AM_ERR_CUDA and "no CPU fallback" in str(e.value)


def _mk(am, start, prom, chunk=0, height=1.0):
    return am.Peak(range(start, start + 1), height, prom, 0.1, 0.1, chunk)


def test_is_overshadowed_matches_reference_kats(native):
    import audio_matcher_b200 as am
    p1, p2, p3 = _mk(am, 3, 1.0), _mk(am, 5, 0.3), _mk(am, 1, 0.2)       # audio_matcher.rs:167-185
    assert am.is_overshadowed(p3, p1, 1, 3) and not am.is_overshadowed(p3, p1, 1, 2)   # :187-197
    assert am.is_overshadowed(p2, p1, 1, 3) and not am.is_overshadowed(p2, p1, 1, 2)
    assert not any(am.is_overshadowed(p, None, 1, 6) for p in (p1, p2, p3))            # :199-207
    assert not am.is_overshadowed(p1, p2, 1, 6) and not am.is_overshadowed(p1, p3, 1, 6)  # :209-218


def test_merge_peaks_matches_oracle(native, orc):
    import audio_matcher_b200 as am
    rng = np.random.default_rng(3)
    for sr, dist in [(48000, 480.0), (8000, 2.0), (44100, 0.0), (1, 3.0)]:
        starts = np.sort(rng.integers(0, sr * 3000 + 50, size=200))
        peaks = [_mk(am, int(s), float(np.float32(rng.random())), chunk=i // 3) for i, s in enumerate(starts)]
        peaks += [_mk(am, int(starts[5]), 0.5, chunk=99)]                # duplicate start from another chunk
        order = rng.permutation(len(peaks))
        got = am.merge_peaks([peaks[i] for i in order], sr, dist)
        ref = orc.merge_peaks([orc.Peak(p.position.start, p.position.stop, p.height, p.prominence, 0.1, 0.1, p.chunk)
                               for p in peaks], sr, dist)
        assert [(p.position.start, p.chunk) for p in got] == [(p.start, p.chunk) for p in ref]
    assert am.merge_peaks([], 48000, 480.0) == []


def test_merge_equal_distance_boundary(native, orc):
    """strict `<` on the distance (audio_matcher.rs:155) with Duration nanosecond truncation (mod.rs:127-129)."""
    import audio_matcher_b200 as am
    sr = 48000
    a, b = _mk(am, 1000, 0.9), _mk(am, 1000 + 480 * sr, 0.5)
    assert len(am.merge_peaks([a, b], sr, 480.0)) == 2                   # exactly 480 s apart: not overshadowed
    b2 = _mk(am, 1000 + 480 * sr - 1, 0.5)
    assert [p.position.start for p in am.merge_peaks([a, b2], sr, 480.0)] == [1000]


def test_config_from_args_defaults():
    import audio_matcher_b200 as am
    c = am.Config.from_args()
    assert (c.chunk_size, c.peak_config.distance) == (60.0, 480.0) and abs(c.peak_config.prominence - 0.13) < 1e-12
    assert am.Config.from_args(prominence_percent=15, distance=8, chunk_size=30).peak_config.prominence == 0.15
    assert list(am.test_data(range(-2, 2))) == [-2.0, -1.0, 0.0, 1.0]    # audio_matcher.rs:481-483
    assert [int(m) for m in am.Mode] == [0, 1, 2]


def test_shard_helpers():
    from audio_matcher_b200.matcher import Config, shard_chunks, shard_frames
    tot = 1441
    cover = []
    for r in range(8):
        f, n = shard_chunks(tot, 8, r)
        cover += list(range(f, f + n))
    assert cover == list(range(tot))
    conf = Config(chunk_size=60.0)
    lo, hi = shard_frames(10, 5, 48000 * 86400, 48000, conf, 480000)
    assert lo == 10 * 2880000 and hi == 15 * 2880000 + 480000            # own chunks + overlap halo
    assert shard_frames(1439, 1, 48000 * 86400, 48000, conf, 480000)[1] == 48000 * 86400


def test_labels():
    import audio_matcher_b200 as am
    from audio_matcher_b200.labels import offset_lines
    peaks = [_mk(am, 21 * 100, 0.9), _mk(am, 1003 * 100, 0.5), _mk(am, 4000 * 100, 0.7)]
    labels = am.timelabel_from_peaks(peaks, 100)                         # archive/data.rs:87-107
    assert [(l.start, l.end, l.name) for l in labels] == [(28.0, 1003.0, "Segment 1"), (1010.0, 4000.0, "Segment 2")]
    assert am.write_labels(labels, None, dry_run=True).splitlines()[0] == "28.000000\t1003.000000\tSegment 1"
    assert offset_lines(peaks, 100)[1] == "Offset 2: 00:16:43 with prominence 0.5"   # mod.rs:110-125
    import ctypes
    f32 = ctypes.c_float(0.98765432).value                                               # what am_peak.prominence hands back
    line = offset_lines([am.Peak(range(100, 101), 1.0, f32, 0.1, 0.1), am.Peak(range(5, 6), 1.0, 1.0, 0, 0), am.Peak(range(7, 8), 1.0, 2.5e-7, 0, 0)], 100)
    assert line[0].endswith("with prominence 0.9876543")                              # Rust `{}` on f32: shortest round-trip
    assert line[1].endswith("with prominence 1") and line[2].endswith("with prominence 0.00000025")
    assert offset_lines([], 100) == ["no offsets found"]


def test_integration_doc_matches_the_shim():
    """INTEGRATION.md shows the reference-side binding; it must be the shipped ffi/cuda_convolve.rs, not a variant, and
    the shim must bind only symbols the header declares, with the reference's calc_chunks parameter list."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = open(os.path.join(root, "ffi", "cuda_convolve.rs")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = doc.split("<!-- BEGIN ffi/cuda_convolve.rs -->")[1].split("<!-- END ffi/cuda_convolve.rs -->")[0]
    assert block.strip() == "```rust\n" + shim.strip() + "\n```"
    header = open(os.path.join(root, "include", "audio_matcher.h")).read()
    bound = re.findall(r"fn (am_\w+)\(", shim.split('extern "C" {')[1].split("}")[0])
    assert len(bound) >= 10 and all(re.search(rf"\b{name}\(", header) for name in bound)
    sig = re.search(r"pub fn calc_chunks<.*?>>\(\n(.*?)\n\) ->", shim, re.S).group(1)
    assert [a.split(":")[0].strip() for a in sig.split(",") if a.strip()] == ["sr", "m_samples", "algo_with_sample", "scale", "config"]
    assert "_config_from" not in shim


def test_calc_chunks_files_argument_checks():
    """Host side of am_calc_chunks_files: one sample format and one memory space per call, sample-rate check first
    (CliError::SampleRateMismatch, src/matcher/mod.rs:72); no device needed for the rejections."""
    import numpy as np
    import pytest
    import audio_matcher_b200 as am
    from audio_matcher_b200.matcher import CudaConvolve

    class Handle:                      # stands in for a matcher: the checks run before the native call
        sr = 8000
        _h = None

        def set_config(self, config):
            raise AssertionError("must not be reached")

    with pytest.raises(ValueError):
        am.calc_chunks_files(44100, [np.zeros(8, np.int16)], Handle(), True, am.Config())
    with pytest.raises(TypeError):
        CudaConvolve._calc_files_raw(Handle(), [np.zeros(8, np.int16), np.zeros(8, np.float32)], True, 16)
    buf, counts = CudaConvolve._calc_files_raw(Handle(), [], True, 16)
    assert counts == []
