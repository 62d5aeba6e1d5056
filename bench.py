#!/usr/bin/env python
"""bench.py -- audio-hours matched per second (snippet vs stream), BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N ...            the reference algorithm's CPU path (oracle port)

A step is one calc_chunks pass over the whole workload (headline: BASELINE.json configs[1], one 10 s snippet vs a
24 h 48 kHz mono int16 stream per GPU, overlap-save block 2^22).  With N > 1 every rank matches its own 24 h shard
of an N*24 h stream (weak scaling; chunk ranges + halo, no data-path collective) and only the peak candidates are
all-gathered over NCCL, inside the C ABI (am_calc_chunks_sharded).  `value` is timed with CUDA events with the PCM
already in HBM; `e2e` is the same call with the PCM in pinned host memory (H2D inside the timed region) and the peak
list read back; `e2e.variants` adds what a Vec<i16> / Vec<f32> caller gets (pageable memory).  The other BASELINE
configs (cfg 1 with its --distance sweep, cfg 3, cfg 4, cfg 5) are measured in the same run and reported in
`other_configs`.  One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sr, channels, snippet_s, stream_hours per GPU, fft_log2, n_snippets)   -- BASELINE.json configs
    "cfg1": (44100, 1, 10.0, 1.0, 22, 1),
    "cfg2": (48000, 1, 10.0, 24.0, 22, 1),
    "cfg3": (48000, 1, 10.0, 24.0, 22, 64),
    "cfg4": (44100, 1, 30.0, 125.0, 23, 1),
    "cfg5": (96000, 2, 2.0, 12.5, 20, 1),
}
METRIC = "audio-hours matched/sec (snippet vs stream)"
UNIT = "audio-hours/s"
CHUNK_S, DIST_S, PROM = 60.0, 480.0, 0.13
PLANT_PERIOD_S, PLANT_JITTER_S = 600.0, 30.0
# dram__bytes_read.sum + dram__bytes_write.sum per block pair at N = 2^22 from one `ncu --set full` capture of a
# 64-pair launch group, keyed by the hash of the kernel sources the capture was taken with: a stale table reports
# traffic = null instead of a wrong number.  (profiles/r02_ncu_full.csv)
NCU_TRAFFIC = {"kernel_src_sha16": "3254771927fa0c71", "file": "profiles/r02_ncu_full.csv",
               "bytes_per_pair": {"k_row": 67.0e6, "k_col_fwd": 48.15e6, "k_col_inv": 40.9e6}}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def kernel_src_sha16() -> str:
    h = hashlib.sha256()
    for f in ("am_fft.cuh", "am_kernels.cuh"):                 # the transform kernels the traffic table is about
        with open(os.path.join(ROOT, "audio_matcher_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores() -> int:
    """Threads the CPU arm may use: the CPUs this process is allowed on.  OMP_NUM_THREADS is deliberately ignored
    (torchrun exports OMP_NUM_THREADS=1, which is not a statement about the host)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_near_gpu(index: int):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the PCI device) so that the pinned host
    buffer of the e2e leg is allocated on that NUMA node; returns the CPU count or None if the topology is unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def plant_plan(sr, snippet_s, total_frames):
    """(offset, shift) of every planted occurrence in the global stream."""
    from oracle import am_oracle as orc
    m, P, J, C = int(round(snippet_s * sr)), int(round(PLANT_PERIOD_S * sr)), int(round(PLANT_JITTER_S * sr)), int(round(CHUNK_S * sr))
    out, k = [], 0
    while k * P + m <= total_frames or k in (3, 4):
        o = orc.plant_offset(k, P, J, C)
        if o + m <= total_frames:
            out.append((o, k % 4))
        k += 1
        if k > 10_000_000:
            break
    return out


def workload_name(name, wl):
    sr, ch, snip_s, hours, fft_log2, n_snip = wl
    return (f"{name}: {'one' if n_snip == 1 else n_snip} {snip_s:g} s snippet{'s' if n_snip > 1 else ''} vs {hours:g} h {sr} Hz {'stereo' if ch == 2 else 'mono'} int16 "
            f"stream per GPU, chunk 60 s, distance 480 s, prominence 0.13")


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm's own CPU path through the oracle port (the Rust crate cannot be built
# in this image: no cargo/rustc, four private git dependencies).
# ---------------------------------------------------------------------------------------------------------
def cpu_sample(wl, chunks, cores, precision=32, distance_s=DIST_S):
    """One pass of the oracle over `chunks` logical chunks of the workload with `cores` threads -> (seconds, peaks)."""
    from oracle import am_oracle as orc
    sr, ch, snip_s, hours, fft_log2, n_snip = wl
    if not hasattr(cpu_sample, "_cache") or cpu_sample._cache[0] != (wl, chunks):
        pcm, snip, _ = orc.synth_case(sr, chunks * CHUNK_S + snip_s, snip_s, channels=ch, chunk_s=CHUNK_S,
                                      plant_period_s=PLANT_PERIOD_S, plant_jitter_s=PLANT_JITTER_S)
        cpu_sample._cache = ((wl, chunks), orc.pcm16_to_f32(pcm, ch), orc.pcm16_to_f32(snip, 1))
    _, x, s = cpu_sample._cache
    cfg = orc.make_config(CHUNK_S, len(s) / sr, distance_s, PROM)
    t = time.perf_counter()
    peaks = orc.calc_chunks(x, s, sr, cfg, scale=True, precision=precision, threads=cores, n_chunks=chunks)
    return time.perf_counter() - t, peaks


def run_reference(args, name, wl, out_stream):
    """Each step is a bounded sample of the workload: one logical chunk per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sr, ch, snip_s, hours, fft_log2, n_snip = wl
    cores = host_cores()
    chunks = cores if args.sample_chunks <= 0 else args.sample_chunks
    stream_s = chunks * CHUNK_S
    times = []
    for i in range(args.warmup + args.steps):
        dt, peaks = cpu_sample(wl, chunks, cores)
        if i >= args.warmup:
            times.append(dt)
        log(f"[reference] step {i}: {dt:.2f}s, {len(peaks)} peaks, {cores} threads")
    total = sum(times)
    value = (stream_s / 3600.0) * len(times) / total / n_snip     # a batch is n_snip independent reference runs
    sample = (f"{chunks} logical chunks ({stream_s:.0f} s of {sr} Hz audio) of the {name} workload per step, {cores} threads "
              f"(sched_getaffinity; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')} ignored)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(name, wl),
                   "note": "oracle port of the reference CPU path (exact-length scalar Bluestein/Stockham FFTs per chunk, snippet FFT "
                           "recomputed per chunk); not rustfft-class -- see cpu_baseline.optimised_port in the b200 arm's line"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=out_stream, flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class Shard:
    """One rank's share of a workload: matcher handle, synthetic PCM resident in HBM, expected offsets."""

    # coloured programme (the loud / coloured bench leg): stream and snippets are boxcar-32 low-pass noise, the stream
    # at 4 x the snippets' RMS (8000 vs 2000), so that scores are noisy (about +-0.16) and every chunk is full of local
    # maxima -- the case summary mode has to survive without a dense repeat of the whole call
    COLOURED = {"taps": 32, "stream_mul": 306, "snip_mul": 76, "shift": 10}

    def __init__(self, name, wl, world, rank, stream, distance_s=DIST_S, total_hours=None, coloured=False):
        import numpy as np
        import torch
        import audio_matcher_b200 as am
        from audio_matcher_b200 import _native as N
        from audio_matcher_b200.matcher import shard_chunks
        from oracle import am_oracle as orc   # workload constants (seeds, plant offsets) + checker only
        self.name, self.wl, self.world, self.rank, self.stream = name, wl, world, rank, stream
        sr, ch, snip_s, hours, fft_log2, n_snip = wl
        self.sr, self.ch, self.n_snip = sr, ch, n_snip
        self.m = m = int(round(snip_s * sr))
        # weak scaling: `hours` per GPU; total_hours fixes the whole job instead (strong scaling)
        self.total_frames = int(round((total_hours if total_hours else hours * world) * 3600 * sr))
        self.conf = am.Config(chunk_size=CHUNK_S, overlap_length=-1.0, peak_config=am.PeakConfig(distance_s, PROM), fft_log2=fft_log2)
        L = N.lib()
        self.coloured = coloured
        if coloured:
            cp = self.COLOURED
            self.snips_np = []
            for i in range(n_snip):
                t = torch.empty(m, dtype=torch.int16, device="cuda")
                N.check(L.am_synth_coloured_pcm16_device(orc.SEED_SNIP + i, 1 << 40, m, cp["taps"], cp["snip_mul"], cp["shift"], t.data_ptr(), stream.cuda_stream))
                torch.cuda.synchronize()
                self.snips_np.append(t.cpu().numpy())
        else:
            self.snips_np = [orc.synth_pcm16(orc.SEED_SNIP + i, 0, m) for i in range(n_snip)]
        if n_snip == 1:
            self.algo = am.CudaConvolve(self.snips_np[0], sr=sr, config=self.conf, stream=stream.cuda_stream)
        else:
            self.algo = am.CudaConvolve(np.stack([orc.pcm16_to_f32(x) for x in self.snips_np]), sr=sr, config=self.conf,
                                        stream=stream.cuda_stream, batch=True)
        total_chunks = self.algo.num_chunks(self.total_frames)
        self.c0, self.nc = shard_chunks(total_chunks, world, rank)
        self.lo, self.hi = self.algo.shard_frames(self.c0, self.nc, self.total_frames)
        n = self.hi - self.lo
        self.pcm = torch.empty((n, ch) if ch == 2 else (n,), dtype=torch.int16, device="cuda")
        if coloured:
            cp = self.COLOURED
            N.check(L.am_synth_coloured_pcm16_device(orc.SEED_STREAM, (1 << 41) + self.lo * ch, n * ch, cp["taps"], cp["stream_mul"], cp["shift"],
                                                     self.pcm.data_ptr(), stream.cuda_stream))
        else:
            N.check(L.am_synth_pcm16_device(orc.SEED_STREAM, self.lo * ch, n * ch, self.pcm.data_ptr(), stream.cuda_stream))
        snips_dev = [torch.from_numpy(x).cuda() for x in self.snips_np]
        self.plan = plant_plan(sr, snip_s, self.total_frames)
        for k, (o, shift) in enumerate(self.plan):             # occurrence k carries snippet k mod n_snip
            if o + m <= self.lo or o >= self.hi:
                continue
            skip = max(0, self.lo - o)
            N.check(L.am_synth_plant_device(self.pcm.data_ptr(), n, ch, snips_dev[k % n_snip].data_ptr() + 2 * skip, m - skip,
                                            o + skip - self.lo, shift, stream.cuda_stream))
        self.expected = {(o, k % n_snip) for k, (o, _) in enumerate(self.plan)}
        torch.cuda.synchronize()

    def step(self, samples):
        """One C-ABI call (am_calc_chunks_range / am_calc_chunks_sharded); the am_peak array it fills is turned into
        Python objects once, after the timed loop."""
        if self.world == 1:
            return self.algo._calc_raw(samples, True, self.total_frames, self.lo, self.c0, self.nc, True, 1 << 14)
        return self.algo.calc_chunks_sharded_raw(samples, True, total_frames=self.total_frames, buf_first_frame=self.lo,
                                                 first_chunk=self.c0, num_chunks=self.nc, cap=1 << 14)

    def timed(self, samples, warmup, steps, profile=False):
        import torch
        import torch.distributed as dist
        for _ in range(warmup):
            peaks = self.step(samples)
        if profile:
            self.algo.set_profiling(True)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            peaks = self.step(samples)
        e1.record(self.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        from audio_matcher_b200.matcher import Peak
        buf, n = peaks
        return ms, [Peak._from_native(buf[i]) for i in range(n)]

    def hours_total(self):
        return self.total_frames / self.sr / 3600.0

    def verify_vs_oracle(self, n_chunks=2):
        """found == oracle on `n_chunks` consecutive logical chunks of this rank's shard (the first ones that hold a
        planted occurrence, else the shard's first): the sub-stream is matched by the GPU path and by the CPU oracle
        as a stream of its own, lists compared before the global filter.  -> (ok, chunks checked, first chunk)"""
        import numpy as np
        from oracle import am_oracle as orc
        C_ = int(round(CHUNK_S * self.sr))
        ov = self.m
        first = self.c0
        for o, _ in self.plan:
            ci = o // C_
            if self.c0 <= ci and ci + n_chunks <= self.c0 + self.nc:
                first = ci
                break
        n_chunks = min(n_chunks, self.nc)
        f0 = C_ * first - self.lo
        f1 = min(self.hi - self.lo, f0 + C_ * n_chunks + ov)
        sub = self.pcm[f0:f1]
        got = self.algo._calc(sub, True, None, 0, 0, n_chunks, False, 1 << 14)
        host = sub.cpu().numpy()
        x = orc.pcm16_to_f32(host.reshape(-1), self.ch)
        ok = True
        for sn in range(min(self.n_snip, 2)):
            s = orc.pcm16_to_f32(self.snips_np[sn], 1)
            ref = orc.calc_chunks(x, s, self.sr, orc.make_config(CHUNK_S, len(s) / self.sr, self.conf.peak_config.distance, PROM),
                                  scale=True, precision=32, threads=min(host_cores(), n_chunks), n_chunks=n_chunks, final_filter=False)
            mine = sorted((p.chunk, p.position.start) for p in got if p.snippet_id == sn)
            theirs = sorted((p.chunk, p.start) for p in ref)
            ok = ok and mine == theirs
            for a, b in zip(sorted((p for p in got if p.snippet_id == sn), key=lambda p: (p.chunk, p.position.start)),
                            sorted(ref, key=lambda p: (p.chunk, p.start))):
                ok = ok and abs(a.height - b.height) <= 1e-4 * abs(b.height) and abs(a.prominence - b.prominence) <= 1e-4 * abs(b.prominence)
        return ok, n_chunks, first

    def close(self):
        import torch
        self.algo.close()
        self.pcm = None
        torch.cuda.empty_cache()


def model_bytes_per_step(frames, b_in, n_fft, m, n_snip):
    """frames * (B_in + 8 + S (8 + 4 sigma)) * N / V_N.  sigma = 1 iff the S spectra (S * 8N bytes)
    exceed what stays L2-resident (order 100 MB)."""
    sigma = 1 if n_snip * 8 * n_fft > 100e6 else 0
    vn = n_fft - m + 1
    return frames * (b_in + 8 + n_snip * (8 + 4 * sigma)) * n_fft / vn, sigma


def measure_config(name, wl, world, rank, stream, steps, warmup, distance_s=DIST_S, total_hours=None, verify_chunks=0, coloured=False):
    """Compact record for `other_configs`: value, step time, model fraction, peak check."""
    sh = Shard(name, wl, world, rank, stream, distance_s=distance_s, total_hours=total_hours, coloured=coloured)
    ms, peaks = sh.timed(sh.pcm, warmup, steps, profile=True)
    ktimes = sh.algo.kernel_times()
    sh.algo.set_profiling(False)
    stats = sh.algo.stats()
    ms_per_step = ms / steps
    value = sh.hours_total() / (ms_per_step / 1000.0)
    starts = [(p.position.start, p.snippet_id) for p in peaks]
    n_fft = 1 << stats["fft_log2"] if stats["fft_log2"] else 0
    peak_gbs, _ = measured_hbm_peak()
    mb, sigma = model_bytes_per_step(sh.hi - sh.lo, 2 * sh.ch, n_fft, sh.m, sh.n_snip) if n_fft else (0, 0)
    label = workload_name(name, wl) if distance_s == DIST_S else f"{name}, --distance {distance_s:g} s"
    if coloured:
        label = (f"{name}, loud coloured programme: stream and snippet are boxcar-32 low-pass noise, stream RMS 4 x the snippet's "
                 "(scores about +-0.16: noise peaks pass the 0.13 prominence, so found > planted is the reference's answer too)")
    rec = {"workload": label,
           "n_gpus": world, "hours_total": sh.hours_total(), "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps,
           "fft_log2": stats["fft_log2"], "four_step": f"{1 << stats['log2_n1']}x{1 << stats['log2_n2']}",
           "summary_mode": stats["summary_mode"], "dense_chunks": stats.get("dense_chunks", 0), "n_snippets": sh.n_snip,
           "snippet_hours_per_s": value * sh.n_snip,
           "model_frac_of_hbm": mb / (ms_per_step / 1000.0) / 1e9 / peak_gbs if mb else None, "model_sigma": sigma,
           "peaks_found": len(starts), "planted": len(sh.plan),
           "verified_offsets_are_planted": None if coloured else (len(starts) > 0 and all(s in sh.expected for s in starts)),
           "kernel_ms_per_step": {k: round(v["total_ms"] / steps, 4) for k, v in ktimes.items()}}
    if verify_chunks:
        ok, nchk, first = sh.verify_vs_oracle(verify_chunks)
        rec["verified_vs_oracle_chunks"] = nchk if ok else 0
        rec["verified_vs_oracle_first_chunk"] = first
    sh.close()
    return rec


def measure_files(name, wl, stream, n_files, steps, warmup):
    """The reference's run over many files (`for main_file in &args.within`, src/matcher/mod.rs:42-99) on its own
    CPU-runnable config: `n_files` files of the workload's length, resident in HBM, matched with ONE
    am_calc_chunks_files call vs one am_calc_chunks call per file."""
    import torch
    sr, ch, snip_s, hours, fft_log2, n_snip = wl
    sh = Shard(name, wl, 1, 0, stream, total_hours=hours * n_files)
    per = int(round(hours * 3600 * sr))
    files = [sh.pcm[i * per:(i + 1) * per] for i in range(n_files)]

    def t_events(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            r = fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, r

    loop_ms, _ = t_events(lambda: [sh.algo._calc_raw(f, True, None, 0, 0, None, True, 1 << 10) for f in files])
    files_ms, (buf, counts) = t_events(lambda: sh.algo._calc_files_raw(files, True, 1 << 14))
    stats = sh.algo.stats()
    found, k = [], 0
    for i, c in enumerate(counts):
        found += [(i * per + buf[k + j].start, 0) for j in range(c)]
        k += c
    total_h = n_files * hours
    rec = {"workload": f"{name} x {n_files} files: {workload_name(name, wl)}; one am_calc_chunks_files call over all files",
           "n_gpus": 1, "hours_total": total_h, "value": total_h / (files_ms / 1e3), "unit": UNIT, "ms_per_step": files_ms, "steps": steps,
           "one_call_per_file": {"value": total_h / (loop_ms / 1e3), "unit": UNIT, "ms_per_step": loop_ms},
           "fft_log2": stats["fft_log2"], "four_step": f"{1 << stats['log2_n1']}x{1 << stats['log2_n2']}",
           "summary_mode": stats["summary_mode"], "dense_chunks": stats.get("dense_chunks", 0), "n_snippets": 1,
           "peaks_found": len(found), "planted": len(sh.plan),
           "verified_offsets_are_planted": len(found) > 0 and all(o in sh.expected for o in found), "kernel_ms_per_step": {}}
    sh.close()
    return rec


def main():
    # keep stdout clean for the single JSON line: libraries (NCCL's version banner, torchrun notices)
    # write to fd 1 too, so everything else goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out_stream):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--hours", type=float, default=0.0, help="override stream hours per GPU")
    ap.add_argument("--fft-log2", type=int, default=0)
    ap.add_argument("--sample-chunks", type=int, default=0, help="CPU baseline sample size in logical chunks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (other_configs)")
    ap.add_argument("--no-verify", action="store_true", help="skip the sampled-chunk comparison with the oracle")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.hours > 0:
        wl[3] = args.hours
    if args.fft_log2:
        wl[4] = args.fft_log2
    wl = tuple(wl)
    if args.impl == "reference":
        run_reference(args, args.workload, wl, out_stream)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import audio_matcher_b200 as am   # noqa: F401  (fails loudly if the CUDA library is missing)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from audio_matcher_b200.matcher import comm_init_from_torch
        comm = comm_init_from_torch()                    # NCCL communicator of the C ABI (am_comm_init)
    stream = torch.cuda.current_stream()
    sr, ch, snip_s, hours, fft_log2, n_snip = wl

    sh = Shard(args.workload, wl, world, rank, stream)
    # the snippet spectrum is computed once per matcher, on its first call: excluded from the timed steps but reported
    # (SURVEY.md 8d).  The first call also allocates the matcher's buffers; both are outside every later call.
    sh.algo.set_profiling(True)
    t_first = time.perf_counter()
    sh.step(sh.pcm)
    torch.cuda.synchronize()
    first_call_ms = 1e3 * (time.perf_counter() - t_first)
    spectrum_ms = sh.algo.kernel_times().get("spectrum", {}).get("total_ms")
    sh.algo.set_profiling(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, peaks = sh.timed(sh.pcm, args.warmup, args.steps, profile=True)
    clocks = sampler.stop()
    ktimes = sh.algo.kernel_times()
    sh.algo.set_profiling(False)
    stats = sh.algo.stats()
    ms_per_step = ms / args.steps
    value = sh.hours_total() / (ms_per_step / 1000.0)
    # sanity: every reported offset is a planted one; the oracle comparison proper is verify_vs_oracle + tests/
    starts = [(p.position.start, p.snippet_id) for p in peaks]
    verified = len(starts) > 0 and all(s in sh.expected for s in starts)
    vs_oracle = None
    if not args.no_verify:
        ok, nchk, first = sh.verify_vs_oracle(2)
        vs_oracle = {"chunks": nchk if ok else 0, "first_chunk": first, "ok": ok}
        if world > 1:
            t = torch.tensor([1.0 if ok else 0.0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            vs_oracle["all_ranks_ok"] = bool(t.item() > 0.5)

    # ---- roofline of the dominant kernel (algorithmic bytes, SURVEY.md 8d / DESIGN.md)
    n_fft = 1 << stats["fft_log2"] if stats["fft_log2"] else 0
    b_in = 2 * ch
    pairs = (stats["fft_blocks"] + 1) // 2
    step_model_bytes, sigma = model_bytes_per_step(sh.hi - sh.lo, b_in, n_fft, sh.m, n_snip) if n_fft else (0, 0)
    model_bytes = {
        "k_col_fwd": pairs * (2 * n_fft * b_in + 8 * n_fft),
        "k_row": pairs * n_fft * (16 if n_snip == 1 else 16 + n_snip * (16 + 8 * sigma)),
        "k_col_inv": pairs * 8 * n_fft * n_snip,
        "k_small": pairs * 2 * n_fft * b_in * n_snip,
    }
    peak_gbs, peak_src = measured_hbm_peak()
    dom = max((k for k in ktimes if k in model_bytes), key=lambda k: ktimes[k]["total_ms"], default=None)
    roofline = None
    if dom:
        per_step_ms = ktimes[dom]["total_ms"] / args.steps
        achieved = model_bytes[dom] / (per_step_ms / 1000.0) / 1e9
        traffic, sha = None, kernel_src_sha16()
        tsrc = f"stale: kernel sources {sha} differ from the ncu capture's ({NCU_TRAFFIC['kernel_src_sha16'] or 'none'})"
        if stats["fft_log2"] == 22 and n_snip == 1 and dom in NCU_TRAFFIC["bytes_per_pair"] and NCU_TRAFFIC["kernel_src_sha16"] == sha:
            traffic = NCU_TRAFFIC["bytes_per_pair"][dom] * pairs / (ktimes[dom]["launches"] / args.steps)   # per launch
            tsrc = f"ncu dram bytes per block pair ({NCU_TRAFFIC['file']}) x pairs per launch"
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": tsrc,
                    "algorithmic_bytes_per_launch": model_bytes[dom] / (ktimes[dom]["launches"] / args.steps),
                    "peak_source": peak_src, "launches_per_step": ktimes[dom]["launches"] / args.steps,
                    "ms_per_step": per_step_ms, "algorithmic_bytes_per_step": model_bytes[dom]}
    kernel_share = {k: round(v["total_ms"] / args.steps, 4) for k, v in ktimes.items()}

    # ---- end to end: PCM in host memory, H2D inside the timed region, peak list read back
    e2e = None
    if not args.no_e2e:
        near = bind_near_gpu(local_rank) if world > 1 else None      # several ranks share the host: keep uploads NUMA-local
        host = torch.empty(sh.pcm.shape, dtype=torch.int16, pin_memory=True)
        host.copy_(sh.pcm)
        torch.cuda.synchronize()
        e_ms, e_peaks = sh.timed(host, min(args.warmup, 2), args.steps)
        est = sh.algo.stats()
        e_step_s = e_ms / args.steps / 1000.0
        verified = verified and [(p.position.start, p.snippet_id) for p in e_peaks] == starts
        h2d_gbs = est["h2d_bytes"] / e_step_s / 1e9
        agg = h2d_gbs
        if world > 1:
            t = torch.tensor([h2d_gbs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            agg = float(t.item())
        e2e = {"value": sh.hours_total() / e_step_s, "unit": UNIT, "h2d_bytes_per_step": est["h2d_bytes"],
               "d2h_bytes_per_step": est["d2h_bytes"], "ms_per_step": e_ms / args.steps, "host_memory": "pinned int16",
               "host_cpus_near_gpu": near, "h2d_gbs_per_gpu": h2d_gbs, "h2d_gbs_aggregate": agg,
               "limiter": "host-to-device copies (PCIe per GPU; with N > 1 the ranks share one host's memory fabric)"}
        # what the drop-in shim's callers hold: pageable Vec<i16> (a decoder's output) and pageable Vec<f32>
        # (ffi/cuda_convolve.rs collects the reference's f32 iterator); fewer steps, they are slow
        variants = []
        if world == 1:
            v_steps = max(1, min(3, args.steps))
            host_np = host.numpy()
            for label, arr in (("pageable int16 (Vec<i16>)", np.array(host_np, copy=True)),
                               ("pageable f32 (Vec<f32>, what ffi/cuda_convolve.rs passes)", None)):
                if arr is None:
                    if ch != 1:
                        continue
                    arr = host_np.astype(np.float32) * np.float32(1.0 / 65535.0)
                v_ms, v_peaks = sh.timed(arr, 1, v_steps)
                vst = sh.algo.stats()
                v_s = v_ms / v_steps / 1000.0
                variants.append({"host_memory": label, "value": sh.hours_total() / v_s, "unit": UNIT, "ms_per_step": v_ms / v_steps,
                                 "h2d_bytes_per_step": vst["h2d_bytes"], "h2d_gbs": vst["h2d_bytes"] / v_s / 1e9, "steps": v_steps,
                                 "offsets_equal": [(p.position.start, p.snippet_id) for p in v_peaks] == starts})
                if arr.dtype == np.float32:
                    # exactly what ffi/cuda_convolve.rs does with the reference's sample iterator: a push session fed with
                    # 4 Mi-sample f32 blocks (am_stream_begin / am_stream_push / am_stream_finish), wall clock
                    from audio_matcher_b200.matcher import StreamSession
                    from audio_matcher_b200 import _native as N
                    t0 = time.perf_counter()
                    st = StreamSession(sh.algo, len(arr), N.FMT_F32_MONO, True, 1 << 14)
                    for i in range(0, len(arr), 1 << 22):
                        st.push(arr[i:i + (1 << 22)])
                    s_peaks = st.finish()
                    dt = time.perf_counter() - t0
                    sst = sh.algo.stats()
                    variants.append({"host_memory": "push session, pageable f32 blocks of 2^22 samples (the Rust shim's calc_chunks)",
                                     "value": sh.hours_total() / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": sst["h2d_bytes"],
                                     "h2d_gbs": sst["h2d_bytes"] / dt / 1e9, "steps": 1, "timing": "wall clock",
                                     "offsets_equal": [(p.position.start, p.snippet_id) for p in s_peaks] == starts})
                del arr
            # many files, one snippet (the loop over args.within, src/matcher/mod.rs:42-99): the same PCM as one-hour files,
            # one am_calc_chunks per file vs ONE am_calc_chunks_files call (uploads overlap matching across files)
            per = sr * 3600
            if ch == 1 and n_snip == 1 and host.shape[0] >= 2 * per:
                files = [host[i * per:(i + 1) * per] for i in range(host.shape[0] // per)]
                f_hours = len(files) * per / sr / 3600.0

                def t_events(fn, reps):
                    fn()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    for _ in range(reps):
                        r = fn()
                    b.record(stream)
                    torch.cuda.synchronize()
                    return a.elapsed_time(b) / reps, r

                loop_ms, _ = t_events(lambda: [sh.algo._calc_raw(f, True, None, 0, 0, None, True, 1 << 10) for f in files], v_steps)
                files_ms, (fbuf, fcounts) = t_events(lambda: sh.algo._calc_files_raw(files, True, 1 << 14), v_steps)
                fst = sh.algo.stats()
                found, k = [], 0
                for i, c in enumerate(fcounts):
                    found += [(i * per + fbuf[k + j].start, 0) for j in range(c)]
                    k += c
                variants.append({"host_memory": f"{len(files)} one-hour files, pinned int16, ONE am_calc_chunks_files call",
                                 "value": f_hours / (files_ms / 1e3), "unit": UNIT, "ms_per_step": files_ms, "h2d_bytes_per_step": fst["h2d_bytes"],
                                 "h2d_gbs": fst["h2d_bytes"] / (files_ms / 1e3) / 1e9, "steps": v_steps,
                                 "one_call_per_file": {"value": f_hours / (loop_ms / 1e3), "unit": UNIT, "ms_per_step": loop_ms},
                                 "peaks_found": len(found), "offsets_are_planted": len(found) > 0 and all(o in sh.expected for o in found)})
            e2e["variants"] = variants
        del host

    frames_here = sh.hi - sh.lo
    peaks_found, planted = len(starts), len(sh.plan)
    sh.close()

    # ---- the other BASELINE configs, bounded: driver-visible in the same line
    others = []
    if not args.no_others and args.workload == "cfg2" and args.hours == 0:
        def add(fn):
            try:
                others.append(fn())
            except Exception as e:                                   # one config failing must not lose the headline line
                others.append({"error": f"{type(e).__name__}: {e}"})
                log("other config failed:", e)
        if world == 1:
            add(lambda: measure_config("cfg1", WORKLOADS["cfg1"], 1, 0, stream, 10, 3, verify_chunks=0 if args.no_verify else 2))
            for d in (8.0, 20.0, 60.0, 120.0):                      # benches/my_benchmark.rs:95, --distance sweep on cfg 1
                add(lambda d=d: measure_config("cfg1", WORKLOADS["cfg1"], 1, 0, stream, 10, 3, distance_s=d))
            add(lambda: measure_files("cfg1", WORKLOADS["cfg1"], stream, 24, 5, 2))
            # summary-mode robustness: the headline workload on loud coloured programme material (ratio to the white-noise step)
            def loud():
                r = measure_config("cfg2", WORKLOADS["cfg2"], 1, 0, stream, 5, 2, verify_chunks=0 if args.no_verify else 2, coloured=True)
                r["step_time_vs_white_noise"] = r["ms_per_step"] / ms_per_step
                return r
            add(loud)
            add(lambda: measure_config("cfg3", WORKLOADS["cfg3"], 1, 0, stream, 2, 1))
            add(lambda: measure_config("cfg4", WORKLOADS["cfg4"], 1, 0, stream, 3, 2, verify_chunks=0 if args.no_verify else 2))
            add(lambda: measure_config("cfg5", WORKLOADS["cfg5"], 1, 0, stream, 5, 2, verify_chunks=0 if args.no_verify else 2))
        else:
            # cfg 4: 125 h per GPU = the BASELINE-named 1000 h archive at 8 GPUs; cfg 5: the 100 h stream split N ways
            add(lambda: measure_config("cfg4", WORKLOADS["cfg4"], world, rank, stream, 3, 2))
            add(lambda: measure_config("cfg5", WORKLOADS["cfg5"], world, rank, stream, 5, 2, total_hours=100.0))

    if rank != 0:
        if world > 1:
            if comm is not None:
                comm.close()
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        chunks = cores if args.sample_chunks <= 0 else args.sample_chunks
        dt, _ = cpu_sample(wl, chunks, cores, precision=32)
        cpu_baseline = {"value": (chunks * CHUNK_S / 3600.0) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{chunks} logical chunks ({chunks * CHUNK_S:.0f} s of audio) of this workload, "
                                  f"exact-length f32 FFTs per chunk like the reference, {dt:.1f} s wall",
                        "note": "scalar any-length FFT port (Bluestein for the prime length 3,839,999), memory-bound beyond ~16 threads; "
                                "not rustfft-class -- optimised_port is the fairer CPU line"}
        # second CPU line: same semantics with power-of-two transforms and a cached snippet spectrum,
        # so the comparison is not inflated by the reference's exact-length transforms
        dt2, _ = cpu_sample(wl, chunks, cores, precision=33)
        cpu_baseline["optimised_port"] = {"value": (chunks * CHUNK_S / 3600.0) / dt2, "unit": UNIT, "cores": cores,
                                          "note": f"power-of-two f32 FFTs + cached snippet spectrum, same sample, {dt2:.1f} s wall"}
        if not args.no_others and args.workload == "cfg2":
            # the CPU baseline's own config is cfg 1 (44.1 kHz, exact length 3,527,999 = 241 x 14639)
            c1 = WORKLOADS["cfg1"]
            n1 = min(chunks, 60)
            sweep = {}
            for d in (DIST_S, 8.0, 20.0, 60.0, 120.0):
                dtd, pk = cpu_sample(c1, n1, cores, precision=32, distance_s=d)
                sweep[f"{d:g}"] = {"value": (n1 * CHUNK_S / 3600.0) / dtd, "wall_s": round(dtd, 2), "peaks": len(pk)}
            cpu_baseline["cfg1"] = {"unit": UNIT, "cores": cores, "sample": f"{n1} of cfg 1's 60 logical chunks", "by_distance_s": sweep}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.workload, wl), "fft_log2": stats["fft_log2"],
                   "four_step": f"{1 << stats['log2_n1']}x{1 << stats['log2_n2']}", "frames_per_gpu": frames_here,
                   "peak_pass": {0: "dense correlation", 1: "run summaries", 2: "run summaries, some chunks repeated densely"}[stats["summary_mode"]],
                   "dense_chunks": stats.get("dense_chunks", 0),
                   "l2_policy": "inputs larger than L2 (PCM per GPU %.1f GB)" % (frames_here * b_in / 1e9),
                   "n_snippets": n_snip, "snippet_hours_per_s": value * n_snip,
                   "peaks_found": peaks_found, "planted": planted, "verified_offsets_are_planted": verified,
                   "verified_vs_oracle_chunks": vs_oracle,
                   "multi_gpu_merge": "ncclAllGather inside am_calc_chunks_sharded (C ABI)" if world > 1 else None,
                   "model_bytes_per_step": step_model_bytes,
                   "model_gbs": step_model_bytes / (ms_per_step / 1000.0) / 1e9,
                   "model_frac_of_hbm": step_model_bytes / (ms_per_step / 1000.0) / 1e9 / peak_gbs,
                   "kernel_ms_per_step": kernel_share, "kernel_src_sha16": kernel_src_sha16(),
                   "snippet_spectrum_precompute": {"kernel_ms": spectrum_ms, "first_call_ms": first_call_ms,
                                                   "note": "once per matcher (fp64 transform of the zero-padded snippet on the device), outside "
                                                           "the timed steps; first_call_ms also holds the one-time buffer allocation"}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": stats["kernel_launches"] * args.steps,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "other_configs": others,
    }
    print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
