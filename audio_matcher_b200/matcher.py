"""Host-side mirror of the reference's matcher interface over the CUDA library.

Same names, argument meaning and error behaviour as src/matcher/audio_matcher.rs:
  Mode (:54-59), Config / PeakConfig (:24-53), trait CorrelateAlgo (:65-76) implemented by
  CudaConvolve (the drop-in for LibConvolve :282-344), calc_chunks (:88-141),
  is_overshadowed (:143-160), test_data (:481-483).
All numeric work happens in libaudio_matcher_b200.so; nothing here computes a correlation.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Iterable, Sequence

import numpy as np

from . import _native as N


class Mode(IntEnum):                      # audio_matcher.rs:54-59
    Full = N.MODE_FULL
    Same = N.MODE_SAME
    Valid = N.MODE_VALID


@dataclass
class PeakConfig:                         # audio_matcher.rs:31-35
    distance: float = 8 * 60.0            # seconds (args.rs:75)
    prominence: float = 0.13              # already /100 (audio_matcher.rs:44)


@dataclass
class Config:                             # audio_matcher.rs:24-30 (the progress-bar arrow is UI only)
    chunk_size: float = 60.0              # seconds (args.rs:71)
    overlap_length: float = -1.0          # seconds; < 0 => snippet duration m / sr (audio_matcher.rs:41)
    peak_config: PeakConfig = field(default_factory=PeakConfig)
    fft_log2: int = 0                     # extension: overlap-save block length override
    max_peaks_per_chunk: int = 0

    @classmethod
    def from_args(cls, prominence_percent: float = 13.0, distance: float | None = None,
                  chunk_size: float | None = None, s_duration: float = -1.0) -> "Config":
        """Config::from_args (audio_matcher.rs:38-52) with the CLI defaults of args.rs:70-76."""
        return cls(chunk_size=60.0 if chunk_size is None else chunk_size, overlap_length=s_duration,
                   peak_config=PeakConfig(distance=480.0 if distance is None else distance,
                                          prominence=prominence_percent / 100.0))

    def _native(self) -> N.AmConfig:
        return N.AmConfig(self.chunk_size, self.overlap_length, self.peak_config.distance,
                          self.peak_config.prominence, self.fft_log2, self.max_peaks_per_chunk, 0)


@dataclass
class Peak:                               # find_peaks::Peak<f32>
    position: range
    height: float
    prominence: float
    left_diff: float
    right_diff: float
    chunk: int = 0
    snippet_id: int = 0

    @property
    def start(self) -> int:
        return self.position.start

    def _native(self) -> N.AmPeak:
        return N.AmPeak(self.position.start, self.position.stop, self.height, self.prominence, self.left_diff,
                        self.right_diff, self.snippet_id, self.chunk)

    @classmethod
    def _from_native(cls, p: N.AmPeak) -> "Peak":
        return cls(range(p.start, p.end), p.height, p.prominence, p.left_diff, p.right_diff, p.chunk, p.snippet_id)


def test_data(rng: Iterable[int]) -> np.ndarray:
    """audio_matcher.rs:481-483: integer range -> Vec<f32>."""
    return np.fromiter((float(i) for i in rng), dtype=np.float32)


def _describe(samples):
    """-> (pointer, frames, fmt, mem, keepalive).  Accepts numpy f32 / int16 arrays (mono 1-D or
    stereo (frames, 2)) in host memory and CUDA tensors of the same dtypes/shapes in device memory."""
    if hasattr(samples, "is_cuda") and hasattr(samples, "data_ptr"):           # torch tensor
        t = samples if samples.is_contiguous() else samples.contiguous()
        name = str(t.dtype)
        if name.endswith("float32"):
            kind = "f32"
        elif name.endswith("int16"):
            kind = "i16"
        else:
            raise TypeError(f"unsupported tensor dtype {t.dtype}")
        shape = tuple(t.shape)
        mem = N.MEM_DEVICE if t.is_cuda else N.MEM_HOST
        ptr = t.data_ptr()
        keep = t
    else:
        a = np.asarray(samples)
        if a.dtype == np.int16:
            kind = "i16"
        else:
            a = np.ascontiguousarray(a, dtype=np.float32)
            kind = "f32"
        a = np.ascontiguousarray(a)
        shape, mem, ptr, keep = a.shape, N.MEM_HOST, a.ctypes.data, a
    if len(shape) == 1:
        fmt = N.FMT_F32_MONO if kind == "f32" else N.FMT_I16_MONO
    elif len(shape) == 2 and shape[1] == 2 and kind == "i16":
        fmt = N.FMT_I16_STEREO                                                  # "can only handle stereo", mp3_reader.rs:26
    elif len(shape) == 2 and shape[1] == 1:
        fmt = N.FMT_F32_MONO if kind == "f32" else N.FMT_I16_MONO
    else:
        raise TypeError(f"unsupported sample layout {shape} / {kind}")
    return ptr, int(shape[0]), fmt, mem, keep


class CudaConvolve:
    """CorrelateAlgo<f32> (audio_matcher.rs:65-76) backed by the CUDA library: the drop-in for
    LibConvolve::new(sample_data) (audio_matcher.rs:289)."""

    def __init__(self, sample_data, sr: int = 48000, config: Config | None = None, stream: int | None = None,
                 batch: bool = False):
        """`batch=True`: sample_data is a float32 array [n_snippets][m] matched as n independent snippets
        that share the stream-side transforms (extension; the reference takes one snippet per run)."""
        self.n_snippets = 1
        if batch:
            a = np.ascontiguousarray(sample_data, dtype=np.float32)
            if a.ndim != 2 or a.shape[0] < 1 or a.shape[1] < 1:
                raise ValueError("batch expects [n_snippets][m]")
            self.sr, self.m, self.n_snippets = int(sr), int(a.shape[1]), int(a.shape[0])
            self._config = config or Config()
            cfg = self._config._native()
            h = C.c_void_p()
            N.check(N.lib().am_matcher_create_batch(a.ctypes.data, self.m, self.n_snippets, self.sr, C.byref(cfg), C.byref(h)))
            self._h = h
            if stream is not None:
                self.set_stream(stream)
            return
        ptr, frames, fmt, mem, keep = _describe(sample_data)
        if mem != N.MEM_HOST:
            raise TypeError("the snippet is taken from host memory (LibConvolve::new owns a copy)")
        if frames == 0:
            raise ValueError("empty snippet")
        self.sr = int(sr)
        self.m = frames
        self._config = config or Config()
        cfg = self._config._native()
        h = C.c_void_p()
        if fmt == N.FMT_F32_MONO:
            N.check(N.lib().am_matcher_create(ptr, frames, self.sr, C.byref(cfg), C.byref(h)))
        else:
            N.check(N.lib().am_matcher_create_pcm16(ptr, frames, 2 if fmt == N.FMT_I16_STEREO else 1, self.sr,
                                                    C.byref(cfg), C.byref(h)))
        self._h = h
        if stream is not None:
            self.set_stream(stream)

    def close(self) -> None:
        if getattr(self, "_h", None):
            N.lib().am_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int) -> None:
        N.check(N.lib().am_matcher_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_config(self, config: Config) -> None:
        self._config = config
        cfg = config._native()
        N.check(N.lib().am_matcher_set_config(self._h, C.byref(cfg)))

    def stats(self) -> dict:
        s = N.AmStats()
        N.check(N.lib().am_matcher_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in N.AmStats._fields_}

    def select_snippet(self, snippet_id: int) -> None:
        N.check(N.lib().am_matcher_select_snippet(self._h, snippet_id))

    def set_profiling(self, on: bool) -> None:
        N.check(N.lib().am_matcher_set_profiling(self._h, int(on)))

    def kernel_times(self) -> dict:
        buf = (N.AmKernelTime * 16)()
        got = C.c_size_t()
        N.check(N.lib().am_matcher_get_kernel_times(self._h, buf, 16, C.byref(got)))
        return {buf[i].name.decode(): {"launches": buf[i].launches, "total_ms": buf[i].total_ms} for i in range(got.value)}

    # --- trait CorrelateAlgo -------------------------------------------------------------
    def inverse_sample_auto_correlation(self) -> float:
        out = C.c_float()
        N.check(N.lib().am_inverse_sample_auto_correlation(self._h, C.byref(out)))
        return out.value

    def correlate_with_sample(self, within, mode: Mode = Mode.Valid, scale: bool = False) -> np.ndarray:
        ptr, n, fmt, mem, keep = _describe(within)
        olen = N.lib().am_out_len(n, self.m, int(mode))
        out = np.empty(olen, dtype=np.float32)
        got = C.c_size_t()
        N.check(N.lib().am_correlate(self._h, ptr, n, fmt, mem, int(mode), int(bool(scale)), out.ctypes.data, olen,
                                     N.MEM_HOST, C.byref(got)))
        return out[:got.value]

    def scale(self, data: np.ndarray) -> None:
        """CorrelateAlgo::scale (audio_matcher.rs:73-75): in-place multiply by the inverse autocorrelation."""
        data *= np.float32(self.inverse_sample_auto_correlation())

    def set_progress(self, fn) -> None:
        """fn(phase, first_chunk, n_chunks) or None: the progress callbacks of audio_matcher.rs:102-117,129
        (phase 0: a segment of logical chunks was submitted to the GPU; phase 1: the call's peaks are final)."""
        self._progress = None if fn is None else N.PROGRESS_FN(lambda user, phase, c0, nc: fn(phase, c0, nc))
        N.check(N.lib().am_matcher_set_progress(self._h, C.cast(self._progress, C.c_void_p) if self._progress else None, None))

    # --- calc_chunks plumbing ---------------------------------------------------------------
    def num_chunks(self, frames: int) -> int:
        return N.lib().am_num_chunks(self._h, frames)

    def chunk_geometry(self) -> tuple[int, int]:
        """(chunk, overlap) in samples as calc_chunks rounds them (audio_matcher.rs:99-100)."""
        c, ov = C.c_size_t(), C.c_size_t()
        N.check(N.lib().am_chunk_geometry(self._h, C.byref(c), C.byref(ov)))
        return c.value, ov.value

    def shard_frames(self, first_chunk: int, num_chunks: int, total_frames: int) -> tuple[int, int]:
        """Frames [lo, hi) a rank must hold for logical chunks [first_chunk, first_chunk + num_chunks): the library's
        own geometry, one source of truth for shard buffers."""
        lo, hi = C.c_size_t(), C.c_size_t()
        N.check(N.lib().am_shard_frames(self._h, total_frames, first_chunk, num_chunks, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def calc_chunks_sharded(self, shard_samples, scale: bool, *, total_frames: int, buf_first_frame: int,
                            first_chunk: int, num_chunks: int, comm: "Comm | None" = None, cap: int = 1 << 16) -> list[Peak]:
        """am_calc_chunks_sharded: this rank's chunk range, one ncclAllGather of the device-resident peak records
        inside the library, global sort + neighbour filter; every rank returns the complete list."""
        buf, n = self.calc_chunks_sharded_raw(shard_samples, scale, total_frames=total_frames, buf_first_frame=buf_first_frame,
                                              first_chunk=first_chunk, num_chunks=num_chunks, comm=comm, cap=cap)
        return [Peak._from_native(buf[i]) for i in range(n)]

    def calc_chunks_sharded_raw(self, shard_samples, scale: bool, *, total_frames: int, buf_first_frame: int,
                                first_chunk: int, num_chunks: int, comm: "Comm | None" = None, cap: int = 1 << 16):
        """-> (ctypes array of am_peak, count): the C-ABI call itself, no Python objects."""
        comm = comm or Comm.current
        if comm is None:
            raise RuntimeError("no communicator: call comm_init_from_torch() / Comm(...) first")
        ptr, frames, fmt, mem, keep = _describe(shard_samples)
        buf = (N.AmPeak * cap)()
        got = C.c_size_t()
        N.check(N.lib().am_calc_chunks_sharded(self._h, comm._c, ptr, buf_first_frame, frames, int(total_frames), fmt, mem,
                                               int(bool(scale)), first_chunk, int(num_chunks), buf, cap, C.byref(got)))
        return buf, got.value

    def _calc(self, samples, scale: bool, total_frames: int | None, buf_first_frame: int, first_chunk: int,
              num_chunks: int | None, final_filter: bool, cap: int) -> list[Peak]:
        buf, n = self._calc_raw(samples, scale, total_frames, buf_first_frame, first_chunk, num_chunks, final_filter, cap)
        return [Peak._from_native(buf[i]) for i in range(n)]

    def _calc_raw(self, samples, scale: bool, total_frames: int | None, buf_first_frame: int, first_chunk: int,
                  num_chunks: int | None, final_filter: bool, cap: int):
        """-> (ctypes array of am_peak, count): the C-ABI call without building Python objects."""
        ptr, frames, fmt, mem, keep = _describe(samples)
        total = frames if total_frames is None else int(total_frames)
        buf = (N.AmPeak * cap)()
        got = C.c_size_t()
        nc = (1 << 62) if num_chunks is None else int(num_chunks)
        N.check(N.lib().am_calc_chunks_range(self._h, ptr, buf_first_frame, frames, total, fmt, mem, int(bool(scale)),
                                             first_chunk, nc, int(final_filter), buf, cap, C.byref(got)))
        return buf, got.value

    def _calc_files_raw(self, files, scale: bool, cap: int):
        """-> (ctypes array of am_peak, per-file counts): one am_calc_chunks_files call over `files` (all in the same
        sample format and memory space)."""
        desc = [_describe(f) for f in files]
        if not desc:
            return (N.AmPeak * 1)(), []
        fmt, mem = desc[0][2], desc[0][3]
        if any(d[2] != fmt or d[3] != mem for d in desc):
            raise TypeError("all files of one call must share the sample format and the memory space")
        n = len(desc)
        ptrs = (C.c_void_p * n)(*[d[0] for d in desc])
        frames = (C.c_size_t * n)(*[d[1] for d in desc])
        counts = (C.c_size_t * n)()
        buf = (N.AmPeak * cap)()
        N.check(N.lib().am_calc_chunks_files(self._h, n, ptrs, frames, fmt, mem, int(bool(scale)), buf, cap, counts))
        return buf, list(counts)


class StreamSession:
    """Push session (am_stream_begin / push / finish): calc_chunks for a decoder that yields the stream in pieces, the
    shape of the reference's lazy sample iterator (mp3_reader.rs:13-66, matcher/mod.rs:71-83).

        with StreamSession(algo, max_frames=claimed_samples) as st:
            for block in decoder:            # numpy int16 (mono or [frames, 2]) or float32 blocks of any size
                st.push(block)
        peaks = st.peaks
    """

    def __init__(self, algo: CudaConvolve, max_frames: int, fmt: int = N.FMT_I16_MONO, scale: bool = True, cap: int = 1 << 16):
        self._algo, self._cap, self.peaks, self._fmt = algo, cap, None, fmt
        s = C.c_void_p()
        N.check(N.lib().am_stream_begin(algo._h, int(max_frames), fmt, int(bool(scale)), C.byref(s)))
        self._s = s

    def push(self, block) -> None:
        ptr, frames, fmt, mem, keep = _describe(block)
        if mem != N.MEM_HOST or fmt != self._fmt:
            raise TypeError("push() takes host blocks in the session's sample format")
        N.check(N.lib().am_stream_push(self._s, ptr, frames))

    def finish(self) -> list[Peak]:
        buf = (N.AmPeak * self._cap)()
        got = C.c_size_t()
        s, self._s = self._s, None
        N.check(N.lib().am_stream_finish(s, buf, self._cap, C.byref(got)))
        self.peaks = [Peak._from_native(buf[i]) for i in range(got.value)]
        return self.peaks

    def abort(self) -> None:
        if self._s:
            N.lib().am_stream_abort(self._s)
            self._s = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None and self._s:
            self.finish()
        else:
            self.abort()
        return False


def calc_chunks_streamed(sr: int, blocks, max_frames: int, algo_with_sample: CudaConvolve, scale: bool, config: Config,
                         fmt: int = N.FMT_I16_MONO, cap: int = 1 << 16) -> list[Peak]:
    """calc_chunks (audio_matcher.rs:88-141) over an iterable of decoded blocks with a claimed length (`with_size`,
    mod.rs:83): matching overlaps the producer."""
    if int(sr) != algo_with_sample.sr:
        raise ValueError(f"sample rate mismatch {algo_with_sample.sr} != {sr}")
    algo_with_sample.set_config(config)
    with StreamSession(algo_with_sample, max_frames, fmt, scale, cap) as st:
        for b in blocks:
            st.push(b)
    return st.peaks


def calc_chunks(sr: int, m_samples, algo_with_sample: CudaConvolve, scale: bool, config: Config,
                cap: int = 1 << 16) -> list[Peak]:
    """calc_chunks (audio_matcher.rs:88-141).  `m_samples` is the decoded stream (numpy array in host
    memory or CUDA tensor in device memory; f32 mono, int16 mono, or int16 stereo (frames, 2))."""
    if int(sr) != algo_with_sample.sr:
        raise ValueError(f"sample rate mismatch {algo_with_sample.sr} != {sr}")       # CliError::SampleRateMismatch
    algo_with_sample.set_config(config)
    return algo_with_sample._calc(m_samples, scale, None, 0, 0, None, True, cap)


def calc_chunks_files(sr: int, files, algo_with_sample: CudaConvolve, scale: bool, config: Config,
                      cap: int = 1 << 16) -> list[list[Peak]]:
    """The loop over args.within of matcher::run (src/matcher/mod.rs:42-99: one calc_chunks per file against the same
    snippet) as ONE library call: the work of all files is queued before the first result is read, so uploads
    overlap matching across files.  -> one peak list per file, each equal to calc_chunks on that file."""
    if int(sr) != algo_with_sample.sr:
        raise ValueError(f"sample rate mismatch {algo_with_sample.sr} != {sr}")       # CliError::SampleRateMismatch
    algo_with_sample.set_config(config)
    buf, counts = algo_with_sample._calc_files_raw(list(files), scale, cap)
    out, k = [], 0
    for c in counts:
        out.append([Peak._from_native(buf[k + i]) for i in range(c)])
        k += c
    return out


def is_overshadowed(element: Peak, other: Peak | None, sr: int, max_distance: float) -> bool:
    """audio_matcher.rs:143-160."""
    if other is None:
        return False
    e, o = element._native(), other._native()
    return bool(N.lib().am_is_overshadowed(C.byref(e), C.byref(o), sr, max_distance))


def merge_peaks(peaks: Sequence[Peak], sr: int, distance: float) -> list[Peak]:
    """Global stable sort by start + filter_surrounding (audio_matcher.rs:135-139) over peaks gathered
    from several shards."""
    n = len(peaks)
    src = (N.AmPeak * max(n, 1))(*[p._native() for p in peaks])
    dst = (N.AmPeak * max(n, 1))()
    got = C.c_size_t()
    N.check(N.lib().am_merge_peaks(src, n, sr, distance, dst, n, C.byref(got)))
    return [Peak._from_native(dst[i]) for i in range(got.value)]


def shard_chunks(total_chunks: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk range [first, first + count) of rank `rank`."""
    first = rank * total_chunks // world_size
    last = (rank + 1) * total_chunks // world_size
    return first, last - first


def shard_frames(first_chunk: int, num_chunks: int, total_frames: int, sr: int, config: Config, m: int) -> tuple[int, int]:
    """Frames [lo, hi) a rank must hold for its chunk range: its chunks plus the overlap halo."""
    # f64::round / llround: half away from zero (Python's round() is half to even); CudaConvolve.shard_frames asks
    # the library itself and is what callers with a handle should use
    C_ = int(math.floor(config.chunk_size * sr + 0.5))
    ov = int(math.floor((config.overlap_length if config.overlap_length >= 0 else m / sr) * sr + 0.5))
    lo = C_ * first_chunk
    hi = min(total_frames, C_ * (first_chunk + num_chunks - 1) + C_ + ov) if num_chunks > 0 else lo
    return lo, hi


def _preload_bundled_nccl() -> None:
    """The library binds NCCL with dlopen("libnccl.so.2") and reuses a copy that is already mapped.  In a Python
    process that will also import torch, the copy torch bundles (nvidia/nccl/lib) must be the one that gets mapped:
    a system libnccl.so.2 loaded first would satisfy torch's own DT_NEEDED by soname and can be too old for it."""
    import importlib.util
    import os
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass                                    # no bundled copy: the library falls back to the system's libnccl.so.2


class Comm:
    """am_comm: the NCCL communicator of the C ABI (one process per GPU).  `unique_id` is the 128-byte id rank 0
    got from Comm.unique_id(); how it reaches the other ranks is the caller's business."""

    current: "Comm | None" = None

    def __init__(self, nranks: int, rank: int, unique_id: bytes):
        if len(unique_id) != N.COMM_ID_BYTES:
            raise ValueError("unique_id must be 128 bytes")
        _preload_bundled_nccl()
        c = C.c_void_p()
        idbuf = C.create_string_buffer(bytes(unique_id), N.COMM_ID_BYTES)
        N.check(N.lib().am_comm_init(nranks, rank, idbuf, C.byref(c)))
        self._c, self.nranks, self.rank = c, nranks, rank
        Comm.current = self

    @staticmethod
    def unique_id() -> bytes:
        _preload_bundled_nccl()
        buf = C.create_string_buffer(N.COMM_ID_BYTES)
        N.check(N.lib().am_comm_get_unique_id(buf))
        return buf.raw

    def close(self) -> None:
        if getattr(self, "_c", None):
            N.lib().am_comm_destroy(self._c)
            self._c = None
            if Comm.current is self:
                Comm.current = None


def comm_init_from_torch(group=None) -> Comm:
    """Bootstrap the C-ABI communicator from an initialised torch.distributed group: rank 0's unique id is broadcast
    as 128 bytes (torch is only the messenger; the data path collective is the library's own ncclAllGather)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(N.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=0, group=group)
    return Comm(world, rank, bytes(t.cpu().numpy().tobytes()))


def _pack(peaks: Sequence[Peak]):
    arr = (N.AmPeak * max(len(peaks), 1))(*[p._native() for p in peaks])
    return arr, len(peaks)


def gather_raw(arr, count: int, group=None, cap: int = 4096):
    """All-gather of per-rank am_peak arrays (KB-sized) through torch.distributed: NCCL between GPUs,
    gloo in the CPU tests.  One collective: every rank contributes a fixed-size record
    [count:int64][cap * sizeof(am_peak) bytes]; a second round only if some rank overflows `cap`."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arr, count
    ws = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    psz = C.sizeof(N.AmPeak)
    while True:
        rec = np.zeros(8 + cap * psz, dtype=np.uint8)
        rec[:8] = np.frombuffer(np.int64(count).tobytes(), dtype=np.uint8)
        n_here = min(count, cap)
        if n_here:
            rec[8:8 + n_here * psz] = np.frombuffer(arr, dtype=np.uint8, count=n_here * psz)
        mine = torch.from_numpy(rec).to(dev)
        out = torch.empty(ws * rec.size, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, mine, group=group)
        host = out.cpu().numpy().reshape(ws, rec.size)
        counts = [int(np.frombuffer(host[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(ws)]
        if max(counts) <= cap:
            break
        cap = max(counts)                                   # rare: somebody had more candidates than the record holds
    total = sum(counts)
    merged = (N.AmPeak * max(total, 1))()
    off = 0
    for r in range(ws):
        nb = counts[r] * psz
        if nb:
            C.memmove(C.addressof(merged) + off, host[r, 8:8 + nb].ctypes.data, nb)
            off += nb
    return merged, total


def gather_peaks(local: Sequence[Peak], group=None) -> list[Peak]:
    arr, n = _pack(local)
    arr, n = gather_raw(arr, n, group)
    return [Peak._from_native(arr[i]) for i in range(n)]


def _merge_raw(arr, n: int, sr: int, distance: float) -> list[Peak]:
    dst = (N.AmPeak * max(n, 1))()
    got = C.c_size_t()
    N.check(N.lib().am_merge_peaks(arr, n, sr, distance, dst, n, C.byref(got)))
    return [Peak._from_native(dst[i]) for i in range(got.value)]


def calc_chunks_sharded(sr: int, shard_samples, algo_with_sample: CudaConvolve, scale: bool, config: Config, *,
                        total_frames: int, buf_first_frame: int, first_chunk: int, num_chunks: int, group=None,
                        cap: int = 1 << 16, set_config: bool = True) -> list[Peak]:
    """Multi-GPU calc_chunks: this rank runs logical chunks [first_chunk, first_chunk + num_chunks) on
    the frames it holds (its range plus the overlap halo, no halo exchange), the per-rank candidates
    are all-gathered, and every rank applies the global sort + neighbour filter."""
    if set_config:
        algo_with_sample.set_config(config)
    if Comm.current is not None and group is None:
        return algo_with_sample.calc_chunks_sharded(shard_samples, scale, total_frames=total_frames, buf_first_frame=buf_first_frame,
                                                    first_chunk=first_chunk, num_chunks=num_chunks, cap=cap)
    # no C-ABI communicator (CPU tests over gloo): gather the records through torch.distributed
    arr, n = algo_with_sample._calc_raw(shard_samples, scale, total_frames, buf_first_frame, first_chunk, num_chunks,
                                        False, cap)
    arr, n = gather_raw(arr, n, group)
    return _merge_raw(arr, n, sr, config.peak_config.distance)
