"""ctypes binding of libaudio_matcher_b200.so (include/audio_matcher.h).

This is the same binding a Rust `extern "C"` block would declare.  The
library is built in-tree by `build_native()` (called from __graft_entry__.build()); there
is no fallback: if the shared object is missing or no CUDA device is visible the compute
entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libaudio_matcher_b200.so"
SOURCES = [_PKG / "csrc" / f for f in ("am_capi.cu", "am_kernels.cuh", "am_peaks.cuh", "am_fft.cuh")] + [
    _PKG.parent / "include" / "audio_matcher.h"]

AM_OK, AM_ERR_INVALID, AM_ERR_CUDA, AM_ERR_CAPACITY, AM_ERR_NOMEM, AM_ERR_UNSUPPORTED = range(6)
MODE_FULL, MODE_SAME, MODE_VALID = 0, 1, 2
FMT_F32_MONO, FMT_I16_MONO, FMT_I16_STEREO = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1

NVCC_FLAGS = ["-std=c++17", "-O3", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "128", "-ldl"]
PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t)
COMM_ID_BYTES = 128
ABI_VERSION = 3


class AmConfig(C.Structure):
    _fields_ = [("chunk_size_s", C.c_double), ("overlap_s", C.c_double), ("distance_s", C.c_double),
                ("prominence", C.c_float), ("fft_log2", C.c_uint32), ("max_peaks_per_chunk", C.c_uint32),
                ("reserved", C.c_uint32)]


class AmPeak(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("height", C.c_float), ("prominence", C.c_float),
                ("left_diff", C.c_float), ("right_diff", C.c_float), ("snippet_id", C.c_uint32),
                ("chunk", C.c_uint32)]


class AmStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("fft_blocks", C.c_uint64), ("frames", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("fft_log2", C.c_uint32),
                ("log2_n1", C.c_uint32), ("log2_n2", C.c_uint32), ("chunks", C.c_uint32),
                ("summary_mode", C.c_uint32), ("dense_chunks", C.c_uint32)]


class AmKernelTime(C.Structure):
    _fields_ = [("kernel_class", C.c_int), ("launches", C.c_uint64), ("total_ms", C.c_double), ("name", C.c_char * 24)]


# every symbol include/audio_matcher.h declares: name -> (restype, argtypes)
_VP, _SZ = C.c_void_p, C.c_size_t
SYMBOLS = {
    "am_last_error": (C.c_char_p, []),
    "am_abi_version": (C.c_int, []),
    "am_device_count": (C.c_int, []),
    "am_config_default": (None, [C.POINTER(AmConfig)]),
    "am_matcher_create": (C.c_int, [_VP, _SZ, C.c_uint32, C.POINTER(AmConfig), C.POINTER(_VP)]),
    "am_matcher_create_pcm16": (C.c_int, [_VP, _SZ, C.c_int, C.c_uint32, C.POINTER(AmConfig), C.POINTER(_VP)]),
    "am_matcher_create_batch": (C.c_int, [_VP, _SZ, _SZ, C.c_uint32, C.POINTER(AmConfig), C.POINTER(_VP)]),
    "am_matcher_snippet_count": (_SZ, [_VP]),
    "am_matcher_select_snippet": (C.c_int, [_VP, _SZ]),
    "am_matcher_destroy": (None, [_VP]),
    "am_matcher_set_stream": (C.c_int, [_VP, _VP]),
    "am_matcher_set_config": (C.c_int, [_VP, C.POINTER(AmConfig)]),
    "am_matcher_get_stats": (C.c_int, [_VP, C.POINTER(AmStats)]),
    "am_matcher_set_profiling": (C.c_int, [_VP, C.c_int]),
    "am_matcher_get_kernel_times": (C.c_int, [_VP, C.POINTER(AmKernelTime), _SZ, C.POINTER(_SZ)]),
    "am_matcher_set_progress": (C.c_int, [_VP, _VP, _VP]),
    "am_inverse_sample_auto_correlation": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "am_out_len": (_SZ, [_SZ, _SZ, C.c_int]),
    "am_valid_len": (_SZ, [_SZ, _SZ]),
    "am_correlate": (C.c_int, [_VP, _VP, _SZ, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _SZ, C.c_int,
                               C.POINTER(_SZ)]),
    "am_num_chunks": (_SZ, [_VP, _SZ]),
    "am_chunk_geometry": (C.c_int, [_VP, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "am_shard_frames": (C.c_int, [_VP, _SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "am_stream_begin": (C.c_int, [_VP, _SZ, C.c_int, C.c_int, C.POINTER(_VP)]),
    "am_stream_push": (C.c_int, [_VP, _VP, _SZ]),
    "am_stream_finish": (C.c_int, [_VP, C.POINTER(AmPeak), _SZ, C.POINTER(_SZ)]),
    "am_stream_abort": (None, [_VP]),
    "am_comm_get_unique_id": (C.c_int, [_VP]),
    "am_comm_init": (C.c_int, [C.c_int, C.c_int, _VP, C.POINTER(_VP)]),
    "am_comm_destroy": (None, [_VP]),
    "am_comm_rank": (C.c_int, [_VP]),
    "am_comm_size": (C.c_int, [_VP]),
    "am_calc_chunks_sharded": (C.c_int, [_VP, _VP, _VP, _SZ, _SZ, _SZ, C.c_int, C.c_int, C.c_int, _SZ, _SZ,
                                         C.POINTER(AmPeak), _SZ, C.POINTER(_SZ)]),
    "am_calc_chunks": (C.c_int, [_VP, _VP, _SZ, C.c_int, C.c_int, C.c_int, C.POINTER(AmPeak), _SZ, C.POINTER(_SZ)]),
    "am_calc_chunks_files": (C.c_int, [_VP, _SZ, C.POINTER(_VP), C.POINTER(_SZ), C.c_int, C.c_int, C.c_int, C.POINTER(AmPeak), _SZ,
                                       C.POINTER(_SZ)]),
    "am_calc_chunks_range": (C.c_int, [_VP, _VP, _SZ, _SZ, _SZ, C.c_int, C.c_int, C.c_int, _SZ, _SZ, C.c_int,
                                       C.POINTER(AmPeak), _SZ, C.POINTER(_SZ)]),
    "am_merge_peaks": (C.c_int, [C.POINTER(AmPeak), _SZ, C.c_uint32, C.c_double, C.POINTER(AmPeak), _SZ,
                                 C.POINTER(_SZ)]),
    "am_is_overshadowed": (C.c_int, [C.POINTER(AmPeak), C.POINTER(AmPeak), C.c_uint32, C.c_double]),
    "am_debug_peaks_from_correlation": (C.c_int, [_VP, _VP, _SZ, C.c_int, C.POINTER(AmPeak), _SZ, C.POINTER(_SZ),
                                                  C.POINTER(C.c_uint32)]),
    "am_synth_pcm16_device": (C.c_int, [C.c_uint64, C.c_uint64, _SZ, _VP, _VP]),
    "am_synth_coloured_pcm16_device": (C.c_int, [C.c_uint64, C.c_uint64, _SZ, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "am_synth_plant_device": (C.c_int, [_VP, _SZ, C.c_int, _VP, _SZ, C.c_uint64, C.c_int, _VP]),
}


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(s.exists() and s.stat().st_mtime > t for s in SOURCES)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libaudio_matcher_b200.so (in-tree)."""
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(SOURCES[0])]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return LIB_PATH


class NativeError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"audio_matcher_b200 status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a) first; "
                              "audio_matcher_b200 has no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != AM_OK:
        raise NativeError(status, lib().am_last_error().decode("utf-8", "replace"))
