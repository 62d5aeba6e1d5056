"""Ad-hoc GPU bring-up script (not a pytest): prints parity diagnostics for several sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import audio_matcher_b200 as am
from audio_matcher_b200 import Mode
from oracle import am_oracle as orc

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def corr_case(n, m, mode, fft_log2=0, seed=0, scale=False):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(n).astype(np.float32); s = rng.standard_normal(m).astype(np.float32)
    algo = am.CudaConvolve(s, sr=1000, config=am.Config(fft_log2=fft_log2))
    got = algo.correlate_with_sample(w, mode, scale)
    st = algo.stats()
    from scipy import signal
    ref = signal.correlate(w.astype(np.float64), s.astype(np.float64), mode=["full", "same", "valid"][int(mode)], method="fft" if n*m > 1e6 else "direct")
    if scale: ref = ref / float((s.astype(np.float64)**2).sum())
    ok = got.shape == ref.shape and rel(got, ref) < 2e-5
    print(f"corr n={n} m={m} mode={mode.name} fft_log2={st['fft_log2']} split={st['log2_n1']}x{st['log2_n2']} launches={st['kernel_launches']} relerr={rel(got, ref) if got.shape==ref.shape else 'shape'} {'OK' if ok else 'FAIL'}", flush=True)
    algo.close()
    return ok

ok = True
# KAT 1
algo = am.CudaConvolve(np.array([1, 2, 3], dtype=np.float32), sr=1)
got = algo.correlate_with_sample(am.test_data(range(-10, 10)), Mode.Valid, False)
print("KAT1", got, "maxerr", np.abs(got - np.arange(-52, 51, 6)).max())
ok &= bool(np.abs(got - np.arange(-52, 51, 6)).max() < 1.2e-5)
print("inv_ac", algo.inverse_sample_auto_correlation(), 1 / 14)
algo.close()
for mode in (Mode.Valid, Mode.Full, Mode.Same):
    ok &= corr_case(4000, 50, mode)
ok &= corr_case(4000, 50, Mode.Valid, scale=True)
for l in range(7, 14):
    ok &= corr_case(20000, 50, Mode.Valid, fft_log2=l, seed=l)
for l in range(14, 25):
    ok &= corr_case(3 * (1 << l) // 2 + 1234, min((1 << l) // 4, 480000), Mode.Valid, fft_log2=l, seed=l)
ok &= corr_case(3_087_000, 441_000, Mode.Valid)
print("CORRELATE", "ALL OK" if ok else "FAILURES", flush=True)

# calc_chunks parity vs oracle
def chunks_case(sr, stream_s, snip_s, chunk_s, dist_s, prom=0.13, fmt="i16", channels=1, fft_log2=0, device=False, maxpk=0):
    pcm, snip, planted = orc.synth_case(sr, stream_s, snip_s, channels=channels, chunk_s=chunk_s, plant_period_s=chunk_s * 2.5, plant_jitter_s=chunk_s / 2)
    x = orc.pcm16_to_f32(pcm, channels); s = orc.pcm16_to_f32(snip, 1)
    cfg = orc.make_config(chunk_s, len(s) / sr, dist_s, prom)
    t = time.time(); ref = orc.calc_chunks(x, s, sr, cfg, scale=True, precision=64); t_or = time.time() - t
    conf = am.Config(chunk_size=chunk_s, overlap_length=-1.0, peak_config=am.PeakConfig(dist_s, prom), fft_log2=fft_log2, max_peaks_per_chunk=maxpk)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    stream = pcm.reshape(-1, 2) if channels == 2 else pcm
    if device:
        import torch
        stream = torch.from_numpy(stream).cuda()
    t = time.time(); got = am.calc_chunks(sr, stream, algo, True, conf); t_gpu = time.time() - t
    st = algo.stats(); algo.close()
    same = [p.position.start for p in got] == [p.start for p in ref] and [p.position.stop for p in got] == [p.end for p in ref]
    herr = max([abs(a.height - b.height) / abs(b.height) for a, b in zip(got, ref)], default=0) if same else -1
    perr = max([abs(a.prominence - b.prominence) / abs(b.prominence) for a, b in zip(got, ref)], default=0) if same else -1
    good = same and herr < 1e-4 and perr < 1e-4
    print(f"chunks sr={sr} {stream_s}s snip={snip_s}s chunk={chunk_s}s dist={dist_s} prom={prom} ch={channels} dev={device}: got {len(got)} ref {len(ref)} planted {len(planted)} offsets_equal={same} herr={herr:.2e} perr={perr:.2e} oracle {t_or:.2f}s gpu {t_gpu:.3f}s fft=2^{st['fft_log2']} launches={st['kernel_launches']} {'OK' if good else 'FAIL'}", flush=True)
    if not same:
        print("  got", [(p.position.start, round(p.height, 4), round(p.prominence, 4)) for p in got][:12])
        print("  ref", [(p.start, round(p.height, 4), round(p.prominence, 4)) for p in ref][:12])
    return good

ok2 = True
ok2 &= chunks_case(8000, 60.0, 0.5, 5.0, 2.0)
ok2 &= chunks_case(8000, 60.0, 0.5, 5.0, 0.0, prom=0.09, maxpk=8000)
ok2 &= chunks_case(8000, 60.0, 0.5, 5.0, 1.0, prom=0.09, maxpk=8000)
ok2 &= chunks_case(8000, 60.0, 0.5, 5.0, 8.0, channels=2)
ok2 &= chunks_case(8000, 61.3, 0.5, 5.0, 2.0, device=True)
ok2 &= chunks_case(8000, 60.0, 2.0, 5.0, 2.0, fft_log2=15)
ok2 &= chunks_case(44100, 600.0, 10.0, 60.0, 480.0)
print("CALC_CHUNKS", "ALL OK" if ok2 else "FAILURES", flush=True)
sys.exit(0 if (ok and ok2) else 1)
