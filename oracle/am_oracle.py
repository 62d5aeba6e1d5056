"""ctypes binding of the CPU oracle (oracle/am_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package.
Parity status: "port" oracle, pinned against the reference's own known-answer
tests (see the header of am_oracle.c); the behaviours listed there as UNPINNED
are defined by this restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libam_oracle.so"

MODE_FULL, MODE_SAME, MODE_VALID = 0, 1, 2


def build(force: bool = False) -> Path:
    src_mtime = max((_HERE / f).stat().st_mtime for f in ("am_oracle.c", "orc_fft.inc"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libam_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrcPeak(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("height", C.c_float),
                ("prominence", C.c_float), ("left_diff", C.c_float), ("right_diff", C.c_float),
                ("chunk", C.c_uint32), ("_pad", C.c_uint32)]


class OrcConfig(C.Structure):
    _fields_ = [("chunk_size_s", C.c_double), ("overlap_s", C.c_double), ("distance_s", C.c_double),
                ("prominence", C.c_float), ("_pad", C.c_float)]


@dataclass
class Peak:
    start: int
    end: int
    height: float
    prominence: float
    left_diff: float
    right_diff: float
    chunk: int = 0


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        f32p, f64p, i16p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int16)
        L.orc_pcm16_to_f32.argtypes = [i16p, C.c_size_t, C.c_int, f32p]
        L.orc_out_len.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        L.orc_out_len.restype = C.c_size_t
        L.orc_correlate_f.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t, C.c_int, f32p]
        L.orc_correlate_d.argtypes = [f64p, C.c_size_t, f64p, C.c_size_t, C.c_int, f64p]
        L.orc_correlate_direct.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t, C.c_int, f64p]
        L.orc_inv_autocorr_f32.argtypes = [f32p, C.c_size_t]
        L.orc_inv_autocorr_f32.restype = C.c_float
        L.orc_inv_autocorr_exact.argtypes = [f32p, C.c_size_t]
        L.orc_inv_autocorr_exact.restype = C.c_double
        L.orc_find_peaks.argtypes = [f32p, C.c_size_t, C.c_int, C.c_float, C.c_size_t,
                                     C.POINTER(OrcPeak), C.c_size_t]
        L.orc_find_peaks.restype = C.c_size_t
        L.orc_is_overshadowed.argtypes = [C.POINTER(OrcPeak), C.POINTER(OrcPeak), C.c_uint32, C.c_double]
        L.orc_is_overshadowed.restype = C.c_int
        L.orc_num_chunks.argtypes = [C.c_size_t, C.c_uint32, C.POINTER(OrcConfig)]
        L.orc_num_chunks.restype = C.c_size_t
        L.orc_calc_chunks_range.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t, C.c_uint32,
                                            C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_int,
                                            C.c_size_t, C.c_size_t, C.c_int, C.POINTER(OrcPeak), C.c_size_t]
        L.orc_calc_chunks_range.restype = C.c_size_t
        L.orc_merge_peaks.argtypes = [C.POINTER(OrcPeak), C.c_size_t, C.c_uint32, C.c_double,
                                      C.POINTER(OrcPeak), C.c_size_t]
        L.orc_merge_peaks.restype = C.c_size_t
        L.orc_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hash64.restype = C.c_uint64
        L.orc_synth_pcm16.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, i16p]
        L.orc_synth_plant.argtypes = [i16p, C.c_size_t, C.c_int, i16p, C.c_size_t, C.c_uint64, C.c_int]
        L.orc_plant_offset.argtypes = [C.c_uint64] * 5
        L.orc_plant_offset.restype = C.c_uint64
        L.orc_start_ns.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_start_ns.restype = C.c_uint64
        L.orc_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def pcm16_to_f32(pcm: np.ndarray, channels: int = 1) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
    frames = pcm.size // channels
    out = np.empty(frames, dtype=np.float32)
    lib().orc_pcm16_to_f32(_p(pcm, C.c_int16), frames, channels, _p(out, C.c_float))
    return out


def out_len(n: int, m: int, mode: int = MODE_VALID) -> int:
    return lib().orc_out_len(n, m, mode)


def correlate(within, sample, mode: int = MODE_VALID, precision: int = 32) -> np.ndarray:
    """precision 32: f32 exact-length FFT (the reference's arithmetic); 64: f64 FFT;
    0: direct O(n*m) sums in double."""
    w, s = _f32(within), _f32(sample)
    olen = out_len(w.size, s.size, mode)
    if precision == 32:
        out = np.empty(olen, dtype=np.float32)
        rc = lib().orc_correlate_f(_p(w, C.c_float), w.size, _p(s, C.c_float), s.size, mode, _p(out, C.c_float))
    elif precision == 64:
        wd, sd = w.astype(np.float64), s.astype(np.float64)
        out = np.empty(olen, dtype=np.float64)
        rc = lib().orc_correlate_d(_p(wd, C.c_double), wd.size, _p(sd, C.c_double), sd.size, mode,
                                   _p(out, C.c_double))
    else:
        out = np.empty(olen, dtype=np.float64)
        rc = lib().orc_correlate_direct(_p(w, C.c_float), w.size, _p(s, C.c_float), s.size, mode,
                                        _p(out, C.c_double))
    if rc != 0:
        raise MemoryError("oracle correlate failed")
    return out


def inv_autocorr(sample, exact: bool = False) -> float:
    s = _f32(sample)
    if exact:
        return float(lib().orc_inv_autocorr_exact(_p(s, C.c_float), s.size))
    return float(lib().orc_inv_autocorr_f32(_p(s, C.c_float), s.size))


def _peaks(buf, n):
    return [Peak(p.start, p.end, p.height, p.prominence, p.left_diff, p.right_diff, p.chunk) for p in buf[:n]]


def find_peaks(y, min_prominence: float = 0.0, min_distance: int = 0, use_prominence: bool = True):
    y = _f32(y)
    cap = y.size // 2 + 1
    buf = (OrcPeak * cap)()
    n = lib().orc_find_peaks(_p(y, C.c_float), y.size, int(use_prominence), min_prominence, min_distance, buf, cap)
    return _peaks(buf, n)


def make_config(chunk_size_s=60.0, overlap_s=0.0, distance_s=480.0, prominence=0.13) -> OrcConfig:
    return OrcConfig(chunk_size_s, overlap_s, distance_s, prominence, 0.0)


def is_overshadowed(element: Peak, other: Peak | None, sr: int, max_distance_s: float) -> bool:
    def conv(p):
        return OrcPeak(p.start, p.end, p.height, p.prominence, p.left_diff, p.right_diff, p.chunk, 0)
    e = conv(element)
    o = C.byref(conv(other)) if other is not None else None
    return bool(lib().orc_is_overshadowed(C.byref(e), o, sr, max_distance_s))


def num_chunks(L: int, sr: int, cfg: OrcConfig) -> int:
    return lib().orc_num_chunks(L, sr, C.byref(cfg))


def calc_chunks(stream, snippet, sr: int, cfg: OrcConfig, scale: bool = True, precision: int = 64,
                threads: int = 0, first_chunk: int = 0, n_chunks: int | None = None,
                final_filter: bool = True, cap: int = 1 << 16):
    w, s = _f32(stream), _f32(snippet)
    buf = (OrcPeak * cap)()
    nc = (1 << 62) if n_chunks is None else n_chunks
    n = lib().orc_calc_chunks_range(_p(w, C.c_float), w.size, _p(s, C.c_float), s.size, sr, C.byref(cfg),
                                    int(scale), precision, threads, first_chunk, nc, int(final_filter), buf, cap)
    if n > cap:
        raise OverflowError(f"{n} peaks > cap {cap}")
    return _peaks(buf, n)


def merge_peaks(peaks, sr: int, distance_s: float):
    n = len(peaks)
    src = (OrcPeak * max(n, 1))()
    for i, p in enumerate(peaks):
        src[i] = OrcPeak(p.start, p.end, p.height, p.prominence, p.left_diff, p.right_diff, p.chunk, i)
    dst = (OrcPeak * max(n, 1))()
    k = lib().orc_merge_peaks(src, n, sr, distance_s, dst, n)
    return _peaks(dst, k)


# ---- synthetic inputs ----------------------------------------
SEED_STREAM, SEED_SNIP, SEED_PLANT = 0x5EED0001, 0x5EED1000, 0x5EED2000


def synth_pcm16(seed: int, first: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.int16)
    lib().orc_synth_pcm16(seed, first, count, _p(out, C.c_int16))
    return out


def synth_plant(pcm: np.ndarray, channels: int, snip: np.ndarray, offset: int, shift: int) -> None:
    assert pcm.dtype == np.int16 and pcm.flags.c_contiguous and snip.dtype == np.int16
    lib().orc_synth_plant(_p(pcm, C.c_int16), pcm.size // channels, channels, _p(snip, C.c_int16), snip.size,
                          offset, shift)


def plant_offset(k: int, P: int, J: int, Cs: int, seed: int = SEED_PLANT) -> int:
    return lib().orc_plant_offset(seed, k, P, J, Cs)


def synth_case(sr: int, stream_s: float, snippet_s: float, channels: int = 1, chunk_s: float = 60.0,
               plant_period_s: float = 600.0, plant_jitter_s: float = 30.0, snippet_id: int = 0):
    """The bench/test workload of SURVEY.md 8d: returns (pcm, snip_pcm, planted[(offset, shift)])."""
    frames, m = int(round(stream_s * sr)), int(round(snippet_s * sr))
    pcm = synth_pcm16(SEED_STREAM, 0, frames * channels)
    snip = synth_pcm16(SEED_SNIP + snippet_id, 0, m)
    P, J, Cs = int(round(plant_period_s * sr)), int(round(plant_jitter_s * sr)), int(round(chunk_s * sr))
    planted = []
    k = 0
    while True:
        o = plant_offset(k, P, J, Cs)
        if k not in (3, 4) and k * P + m > frames:
            break
        if o + m <= frames:
            synth_plant(pcm, channels, snip, o, k % 4)
            planted.append((o, k % 4))
        k += 1
        if k > 100000:
            break
    return pcm, snip, planted


def threads() -> int:
    return lib().orc_threads()
