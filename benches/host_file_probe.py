#!/usr/bin/env python
"""One file of `hours` (default 1) of 44.1 kHz mono PCM (cfg 1) through calc_chunks from device, pinned and pageable memory:
wall-clock ms per call.  python benches/host_file_probe.py [hours]; AM_STAGE_* / AM_SEGMENT_MB vary the staging.  Needs a GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_matcher_b200 as am
from oracle import am_oracle as orc
sr, m = 44100, 441000
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frames = int(hours * 3600 * sr)
conf = am.Config(chunk_size=60.0, peak_config=am.PeakConfig(480.0, 0.13), fft_log2=22)
algo = am.CudaConvolve(orc.synth_pcm16(orc.SEED_SNIP, 0, m), sr=sr, config=conf)
pcm = orc.synth_pcm16(orc.SEED_STREAM, 0, frames)
dev = torch.from_numpy(pcm).cuda()
pin = torch.from_numpy(pcm).pin_memory()
def t(x, n=20):
    am.calc_chunks(sr, x, algo, True, conf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): am.calc_chunks(sr, x, algo, True, conf)
    return (time.perf_counter() - t0) / n * 1e3
print(f"{hours} h: device {t(dev):.2f} ms  pinned {t(pin):.2f} ms  pageable {t(pcm):.2f} ms  bytes {frames*2/1e6:.0f} MB  threads {os.environ.get('AM_STAGE_THREADS')} seg {os.environ.get('AM_SEGMENT_MB')}")
