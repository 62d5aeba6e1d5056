"""Randomised parity soak (run by hand on a GPU box: python tests/fuzz_gpu.py [seconds] [seed]).
Random stream / snippet material (white, coloured, tonal, loud, quiet), chunk geometry (incl. ov != m), block length,
minimum distance, prominence and candidate cap; calc_chunks through the C ABI (one-shot, push session, 3-way shard +
merge, several files per call) must equal the CPU oracle: offsets bit-exact, heights / prominences within 1e-4 relative."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audio_matcher_b200 as am  # noqa: E402
from audio_matcher_b200 import _native as N  # noqa: E402
from oracle import am_oracle as orc  # noqa: E402  (checker)


def material(rng, n, kind, rms):
    if kind == "white":
        x = rng.standard_normal(n)
    elif kind == "coloured":
        from scipy import signal
        x = signal.lfilter([1.0], [1.0, -float(rng.uniform(0.8, 0.99))], rng.standard_normal(n))
    else:
        t = np.arange(n)
        f = rng.uniform(50, 2000)
        x = np.sin(2 * np.pi * f / 8000 * t + rng.uniform(0, 6)) + 0.2 * rng.standard_normal(n)
    x *= rms / np.sqrt(np.mean(x * x))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


counts_ties = [0]


class Borderline(Exception):
    pass


def one_case(rng, case):
    sr = 8000
    m = int(rng.integers(300, 6000))
    chunk_s = float(rng.choice([2.0, 2.5, 5.0, 7.3]))
    n = int(rng.integers(m + 10, sr * 60))
    kind = str(rng.choice(["white", "coloured", "tonal"]))
    snip = material(rng, m, kind, float(rng.uniform(800, 4000)))
    pcm = material(rng, n, kind, float(rng.uniform(500, 9000)))
    for k in range(int(rng.integers(0, 6))):
        o = int(rng.integers(0, max(1, n - m)))
        seg = pcm[o:o + m].astype(np.int32) // 2 + (snip.astype(np.int32) >> int(rng.integers(0, 3)))
        pcm[o:o + m] = np.clip(seg, -32768, 32767).astype(np.int16)
    ov_mode = int(rng.integers(0, 4))
    overlap = [-1.0, (m - 1) / sr, (m + int(rng.integers(1, 40))) / sr, max(0.0, (m - int(rng.integers(2, 200))) / sr)][ov_mode]
    dist = float(rng.choice([0.0, 1.0, 2.0, 5.0, 480.0]))
    prom = float(rng.choice([0.05, 0.13, 0.13, 0.3, 0.6]))
    log2 = int(rng.choice([0, 0, 14, 15, 17, 20]))
    while log2 and (1 << log2) < 2 * m:
        log2 += 1
    cap = int(rng.choice([0, 0, 16, 64, 300]))
    x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
    ov_s = len(s) / sr if overlap < 0 else overlap
    try:
        ref = orc.calc_chunks(x, s, sr, orc.make_config(chunk_s, ov_s, dist, prom), scale=True, precision=64, cap=1 << 20)
    except OverflowError:
        return "skip"
    if len(ref) > 20000:
        return "skip"
    conf = am.Config(chunk_size=chunk_s, overlap_length=overlap, peak_config=am.PeakConfig(dist, prom), fft_log2=log2, max_peaks_per_chunk=cap)
    algo = am.CudaConvolve(snip, sr=sr, config=conf)
    desc = dict(case=case, m=m, n=n, kind=kind, chunk_s=chunk_s, overlap=overlap, dist=dist, prom=prom, log2=log2, cap=cap, ref=len(ref))

    def same(got, what):
        a = [(p.position.start, p.position.stop, p.chunk) for p in got]
        b = [(p.start, p.end, p.chunk) for p in ref]
        if a != b:
            # the only accepted difference: peaks whose prominence sits on the threshold to fp32 rounding (an f64 oracle and
            # an f32 correlation may land on either side); with a minimum distance such a peak can suppress others, so the
            # whole case is then set aside
            pa = {k: p.prominence for k, p in zip(a, got)}
            pb = {k: p.prominence for k, p in zip(b, ref)}
            only_a, only_b = set(a) - set(b), set(b) - set(a)
            # ... and plateaus: two adjacent outputs that are equal to fp32 rounding in one computation and one ulp apart
            # in the other give the same peak with a start one sample off
            shifted = {k for k in only_a for q in only_b if k[2] == q[2] and abs(k[0] - q[0]) <= 2 and abs(k[1] - q[1]) <= 2
                       and abs(pa[k] - pb[q]) <= 1e-5 * max(1.0, abs(pb[q]))}
            if shifted and len(shifted) == len(only_a) == len(only_b):
                counts_ties[0] += len(shifted)
                raise Borderline()
            # ... and twins: two local maxima of one chunk whose heights agree to fp32 rounding.  The walk of the (by a
            # rounding) lower one stops at the other, so it keeps only the dip between them as prominence and falls
            # under the threshold; which of the two survives may differ between an f64 and an f32 correlation.
            ha = {k: p.height for k, p in zip(a, got)}
            hb = {k: p.height for k, p in zip(b, ref)}
            twins = {k for k in only_a for q in only_b if k[2] == q[2] and abs(ha[k] - hb[q]) <= 2e-6 * max(1.0, abs(hb[q]))
                     and abs(pa[k] - pb[q]) <= 1e-5 * max(1.0, abs(pb[q]))}
            if twins and len(twins) == len(only_a) == len(only_b):
                counts_ties[0] += len(twins)
                raise Borderline()
            odd = [pa.get(k, pb.get(k)) for k in set(a) ^ set(b)]
            assert odd and (dist > 0 or all(abs(q - prom) <= 1e-4 * prom for q in odd)), (what, desc, sorted(set(a) ^ set(b))[:6], odd[:6], len(a), len(b))
            assert any(abs(q - prom) <= 1e-4 * prom for q in list(pa.values()) + list(pb.values())), (what, desc, len(a), len(b))
            raise Borderline()
        for p, r in zip(got, ref):
            assert abs(p.height - r.height) <= 1e-4 * abs(r.height) + 2e-6, (what, desc, p, r)   # (absolute floor: heights near zero, scores are O(1))
            if abs(p.prominence - r.prominence) > 1e-4 * abs(r.prominence) + 2e-6:
                # A prominence walk stops at the first sample STRICTLY higher than the peak: two peaks of one chunk whose
                # heights agree to fp32 rounding (periodic / coloured material) can legitimately swap roles between an
                # f64 and an f32 correlation.  Accept exactly that case, nothing else.
                ties = [q for q in ref if q.chunk == r.chunk and q.start != r.start and abs(q.height - r.height) <= 2e-6 * max(1.0, abs(r.height))]
                assert ties, (what, desc, p, r)
                counts_ties[0] += 1
    try:
        same(am.calc_chunks(sr, pcm, algo, True, conf, cap=1 << 20), "one-shot")
        piece = int(rng.integers(500, 200000))
        same(am.calc_chunks_streamed(sr, (pcm[i:i + piece] for i in range(0, n, piece)), n + int(rng.integers(0, 1000)), algo, True, conf,
                                     cap=1 << 20), "push session")
        total = algo.num_chunks(n)
        parts = []
        for r in range(3):
            a, b = r * total // 3, (r + 1) * total // 3
            lo, hi = algo.shard_frames(a, b - a, n)
            if b > a:
                parts += algo._calc(np.ascontiguousarray(pcm[lo:hi]), True, n, lo, a, b - a, False, 1 << 20)
        same(am.merge_peaks(parts, sr, dist), "shards + merge")
        # several files in one call (am_calc_chunks_files): the stream, its first half, the stream again
        half = np.ascontiguousarray(pcm[:n // 2])
        multi = am.calc_chunks_files(sr, [pcm, half, pcm], algo, True, conf, cap=1 << 20)
        same(multi[0], "files[0]")
        same(multi[2], "files[2]")
        key = lambda l: [(p.position.start, p.position.stop, p.height, p.prominence, p.chunk) for p in l]
        assert key(multi[1]) == key(am.calc_chunks(sr, half, algo, True, conf, cap=1 << 20)), ("files[1]", desc)
    except Borderline:
        return "threshold-borderline"
    except N.NativeError as e:
        if e.status == N.AM_ERR_CAPACITY and (cap or dist == 0.0):
            return "capacity"                      # a cap of 16..300 kept peaks per chunk can legitimately be exceeded
        raise AssertionError((desc, str(e)))
    finally:
        algo.close()
    return "ok"


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t0, counts, case = time.time(), {}, 0
    while time.time() - t0 < budget:
        r = one_case(rng, case)
        counts[r] = counts.get(r, 0) + 1
        case += 1
    print("fuzz done:", counts, f"in {time.time() - t0:.0f} s;", counts_ties[0], "prominences differed through an fp32 height tie")


if __name__ == "__main__":
    main()
