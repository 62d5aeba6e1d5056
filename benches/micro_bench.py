#!/usr/bin/env python
"""Criterion-shaped micro benches mirroring /root/reference/benches/my_benchmark.rs with synthetic inputs.

  correlate_vs_bib      (:29-53)  snippet 100..150 vs stream -2000..2000, Mode::Valid, unscaled
  correlate_vs_conj     (:55-79)  same shapes; here: direct-sum kernel vs block-FFT kernel
  correlate_chunk       (audio_matcher.rs:120-122) correlate_with_sample on one logical chunk window, host buffers
  compare_chunk_sizes   (:81-108) full calc_chunks with --distance 8/20/60/120 s on the 1 h 44.1 kHz workload
                                  (the reference's res/local/*.mp3 are not shipped; BASELINE.json configs[0] shapes)

One JSON line per bench on stdout: {"group", "name", "mean_us", "iters"}.  Needs a GPU.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import audio_matcher_b200 as am  # noqa: E402
from oracle import am_oracle as orc  # noqa: E402  (CPU comparison arm only)


def bench(fn, min_time=0.5, max_iters=2000):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_time or n >= max_iters:
            return 1e6 * dt / n, n


def emit(group, name, fn, **kw):
    us, n = bench(fn, **kw)
    print(json.dumps({"group": group, "name": name, "mean_us": round(us, 2), "iters": n}), flush=True)


def main():
    data1 = am.test_data(range(100, 150))          # my_benchmark.rs:31
    data2 = am.test_data(range(-2000, 2000))       # my_benchmark.rs:32
    algo = am.CudaConvolve(data1, sr=1)
    emit("correlate_vs_bib", "correlate cuda func (host buffers)", lambda: algo.correlate_with_sample(data2, am.Mode.Valid, False))
    emit("correlate_vs_bib", "correlate old func (CPU oracle port, f32 exact-length FFT)",
         lambda: orc.correlate(data2, data1, orc.MODE_VALID, 32))
    os.environ["AM_NO_DIRECT"] = "1"
    emit("correlate_vs_conj", "block-FFT kernel", lambda: algo.correlate_with_sample(data2, am.Mode.Valid, False))
    os.environ.pop("AM_NO_DIRECT")
    emit("correlate_vs_conj", "direct-sum kernel", lambda: algo.correlate_with_sample(data2, am.Mode.Valid, False))
    algo.close()

    sr, snip_s = 44100, 10.0
    pcm, snip, planted = orc.synth_case(sr, 3600.0, snip_s)
    import torch
    dev = torch.from_numpy(pcm).cuda()
    for distance in (8, 20, 60, 120):              # my_benchmark.rs:95
        conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(float(distance), 0.13), fft_log2=22)
        algo = am.CudaConvolve(snip, sr=sr, config=conf)
        found = len(am.calc_chunks(sr, dev, algo, True, conf))
        emit("compare_chunk_sizes", f"peaks in 1 h synthetic stream/{distance} (device PCM, {found} peaks)",
             lambda: am.calc_chunks(sr, dev, algo, True, conf), min_time=1.0, max_iters=200)
        emit("compare_chunk_sizes", f"peaks in 1 h synthetic stream/{distance} (host PCM, H2D included)",
             lambda: am.calc_chunks(sr, pcm, algo, True, conf), min_time=1.0, max_iters=200)
        algo.close()

    # the fine seam alone (CorrelateAlgo::correlate_with_sample, audio_matcher.rs:67-72, what calc_chunks calls once per
    # logical chunk, :120-122): one 60 s + overlap window of f32 samples in host memory in, the scaled Valid correlation in
    # host memory out -- what a maintainer gets by swapping only LibConvolve::new for CudaConvolve::new
    window = orc.pcm16_to_f32(pcm[:60 * sr + len(snip)])
    algo = am.CudaConvolve(snip, sr=sr, config=am.Config(fft_log2=22))
    emit("correlate_chunk", f"correlate_with_sample, one {len(window)}-sample window vs {len(snip)}-sample snippet (host f32 in / out)",
         lambda: algo.correlate_with_sample(window, am.Mode.Valid, True), min_time=1.0, max_iters=200)
    import time as _t
    t0 = _t.perf_counter()
    orc.correlate(window, orc.pcm16_to_f32(snip), orc.MODE_VALID, 32)
    print(json.dumps({"group": "correlate_chunk", "name": "the same window through the CPU oracle port (f32 exact-length FFT, 1 thread)",
                      "mean_us": round(1e6 * (_t.perf_counter() - t0), 2), "iters": 1}), flush=True)
    algo.close()


if __name__ == "__main__":
    main()
