# experiment helper: per-kernel times of the cfg2 path for the currently installed library build,
# tolerant of deliberately broken ablation builds (results are not checked)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_matcher_b200 as am
from audio_matcher_b200 import _native as N
from oracle import am_oracle as orc
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 12.0
sr, m = 48000, 480000
frames = int(hours * 3600 * sr)
conf = am.Config(chunk_size=60.0, overlap_length=-1.0, peak_config=am.PeakConfig(480.0, 1e30), fft_log2=22)
stream = torch.cuda.current_stream()
algo = am.CudaConvolve(orc.synth_pcm16(orc.SEED_SNIP, 0, m), sr=sr, config=conf, stream=stream.cuda_stream)
L = N.lib()
pcm = torch.empty(frames, dtype=torch.int16, device="cuda")
N.check(L.am_synth_pcm16_device(orc.SEED_STREAM, 0, frames, pcm.data_ptr(), stream.cuda_stream))
torch.cuda.synchronize()
nc = algo.num_chunks(frames)
def step():
    try:
        algo._calc(pcm, True, frames, 0, 0, nc, True, 1 << 16)
    except Exception as e:
        step.err = str(e)[:80]
step.err = None
for _ in range(3): step()
algo.set_profiling(True)
torch.cuda.synchronize()
for _ in range(5): step()
torch.cuda.synchronize()
kt = algo.kernel_times()
print({k: round(v["total_ms"] / 5, 3) for k, v in kt.items()}, step.err)
