"""Regenerates tests/golden/*.json.

The Rust reference cannot be built or imported in this image, so the golden vectors are
 (1) the known answers held by the reference's own tests (copied as data with file:line), and
 (2) outputs of the pinned CPU oracle on seeded synthetic inputs (regression vectors for the
     CUDA path; the inputs are regenerated from the seeds, only the expected peaks are stored).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import am_oracle as orc  # noqa: E402

reference_kats = {
    "correlate_valid": {   # src/matcher/audio_matcher.rs:489-517
        "cite": "src/matcher/audio_matcher.rs:489-517 (my_correlate_same_fftcorrelate)",
        "within_range": [-10, 10], "sample": [1.0, 2.0, 3.0], "mode": "Valid", "scale": False,
        "expect": list(range(-52, 51, 6)), "abs_tol": 1.2e-5},
    "find_peaks": {        # src/matcher/audio_matcher.rs:167-185
        "cite": "src/matcher/audio_matcher.rs:167-185 (overshadow_tests::test_data)",
        "y": [0.0, 0.7, 0.5, 1.0, 0.5, 0.8, 0.0], "min_prominence": 0.0,
        "expect_starts_in_order": [3, 5, 1], "expect_prominences": [1.0, 0.3, 0.2], "abs_tol": 1e-6},
    "is_overshadowed": {   # src/matcher/audio_matcher.rs:187-218
        "cite": "src/matcher/audio_matcher.rs:187-218",
        "sr": 1,
        "cases": [["p3", "p1", 3, True], ["p3", "p1", 2, False], ["p2", "p1", 3, True], ["p2", "p1", 2, False],
                  ["p1", None, 6, False], ["p2", None, 6, False], ["p3", None, 6, False],
                  ["p1", "p2", 6, False], ["p1", "p3", 6, False]]},
    "bench_shapes": {      # benches/my_benchmark.rs:29-79 (no expected values in the reference; numpy-derived)
        "cite": "benches/my_benchmark.rs:29-79", "sample_range": [100, 150], "within_range": [-2000, 2000],
        "valid_len": 3951, "first_value": -12287075.0, "sum_squares": 785425.0},
}

CASES = [  # (name, sr, stream_s, snippet_s, chunk_s, overlap_s or None, distance_s, prominence, channels)
    ("mono_8k", 8000, 60.0, 0.5, 5.0, None, 2.0, 0.13, 1),
    ("stereo_8k", 8000, 47.3, 0.5, 5.0, None, 8.0, 0.13, 2),
    ("gap_overlap_short", 8000, 40.0, 0.5, 5.0, 0.4, 2.0, 0.13, 1),     # ov < m - 1: untested offsets per boundary
    ("dup_overlap_long", 8000, 40.0, 0.5, 5.0, 0.75, 0.0, 0.2, 1),      # ov > m: duplicate peaks from both chunks
    ("mono_16k_1s", 16000, 90.0, 1.0, 10.0, None, 20.0, 0.13, 1),
]


def run_case(c):
    name, sr, stream_s, snip_s, chunk_s, ov_s, dist_s, prom, ch = c
    pcm, snip, planted = orc.synth_case(sr, stream_s, snip_s, channels=ch, chunk_s=chunk_s,
                                        plant_period_s=chunk_s * 2.5, plant_jitter_s=chunk_s / 2)
    x, s = orc.pcm16_to_f32(pcm, ch), orc.pcm16_to_f32(snip, 1)
    cfg = orc.make_config(chunk_s, len(s) / sr if ov_s is None else ov_s, dist_s, prom)
    peaks = orc.calc_chunks(x, s, sr, cfg, scale=True, precision=64)
    return {"name": name, "sr": sr, "stream_s": stream_s, "snippet_s": snip_s, "chunk_s": chunk_s, "overlap_s": ov_s,
            "distance_s": dist_s, "prominence": prom, "channels": ch, "planted": planted,
            "pcm_checksum": int(pcm.astype("int64").sum()),
            "peaks": [[p.start, p.end, p.height, p.prominence, p.chunk] for p in peaks]}


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(reference_kats, f, indent=1)
    with open(os.path.join(HERE, "oracle_cases.json"), "w") as f:
        json.dump([run_case(c) for c in CASES], f, indent=1)
    print("golden vectors written")
