//! Drop-in `CorrelateAlgo` / `calc_chunks` for NilsJochem/audio-matcher backed by libaudio_matcher_b200.so.
//!
//! Where it goes: `src/matcher/audio_matcher/cuda_convolve.rs`, declared as `pub mod cuda_convolve;` INSIDE
//! `src/matcher/audio_matcher.rs` -- a child module may read the private fields of `Config` / `PeakConfig`
//! (audio_matcher.rs:24-35), so the reference's types need no new getters.  Link with `-l audio_matcher_b200` and
//! swap two names in `matcher::run` (src/matcher/mod.rs:34,81), see INTEGRATION.md.
//! It binds entry points declared in `include/audio_matcher.h` and nothing else.
//!
//! UNTESTED SOURCE: the build image has no Rust toolchain (no cargo / rustc) and the crate's four private git
//! dependencies are unreachable, so this file has never been compiled.  The same ABI calls, in the same order, are
//! exercised from C (`tests/capi_smoke.c`), C++ (`include/audio_matcher.hpp`) and Python (`audio_matcher_b200`).
use std::{ffi::CStr, os::raw::{c_char, c_int, c_void}, ptr};

use super::{Config, CorrelateAlgo, Mode};
use crate::matcher::mp3_reader::SampleType;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct AmConfig {
    pub chunk_size_s: f64,
    pub overlap_s: f64,
    pub distance_s: f64,
    pub prominence: f32,
    pub fft_log2: u32,
    pub max_peaks_per_chunk: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct AmPeak {
    pub start: u64,
    pub end: u64,
    pub height: f32,
    pub prominence: f32,
    pub left_diff: f32,
    pub right_diff: f32,
    pub snippet_id: u32,
    pub chunk: u32,
}

#[repr(C)]
pub struct AmMatcher {
    _private: [u8; 0],
}
#[repr(C)]
pub struct AmStreamSession {
    _private: [u8; 0],
}

pub const AM_FMT_F32_MONO: c_int = 0;
pub const AM_MEM_HOST: c_int = 0;

#[link(name = "audio_matcher_b200")]
extern "C" {
    fn am_last_error() -> *const c_char;
    fn am_matcher_create(snippet: *const f32, m: usize, sr: u32, cfg: *const AmConfig, out: *mut *mut AmMatcher) -> c_int;
    fn am_matcher_destroy(h: *mut AmMatcher);
    fn am_matcher_set_config(h: *mut AmMatcher, cfg: *const AmConfig) -> c_int;
    fn am_inverse_sample_auto_correlation(h: *mut AmMatcher, out: *mut f32) -> c_int;
    fn am_out_len(n: usize, m: usize, mode: c_int) -> usize;
    fn am_correlate(h: *mut AmMatcher, within: *const c_void, n: usize, fmt: c_int, within_mem: c_int, mode: c_int,
                    scale: c_int, out: *mut f32, cap: usize, out_mem: c_int, out_len: *mut usize) -> c_int;
    fn am_calc_chunks_files(h: *mut AmMatcher, n_files: usize, streams: *const *const c_void, frames: *const usize, fmt: c_int,
                            mem: c_int, scale: c_int, out: *mut AmPeak, cap: usize, n_out: *mut usize) -> c_int;
    fn am_stream_begin(h: *mut AmMatcher, max_frames: usize, fmt: c_int, scale: c_int, out: *mut *mut AmStreamSession) -> c_int;
    fn am_stream_push(s: *mut AmStreamSession, pcm: *const c_void, frames: usize) -> c_int;
    fn am_stream_finish(s: *mut AmStreamSession, out: *mut AmPeak, cap: usize, n_out: *mut usize) -> c_int;
    fn am_stream_abort(s: *mut AmStreamSession);
}

fn last_error() -> Box<dyn std::error::Error> {
    unsafe { CStr::from_ptr(am_last_error()) }.to_string_lossy().into_owned().into()
}

/// Replaces `LibConvolve` (src/matcher/audio_matcher.rs:282-344).
pub struct CudaConvolve {
    h: *mut AmMatcher,
    m: usize,
    sr: u16,
}
// Every entry point of the library takes the handle's mutex (am_capi.cu), so a shared `&CudaConvolve` is safe to
// use from the rayon workers of the reference's generic calc_chunks (audio_matcher.rs:114-122); calls serialise.
unsafe impl Send for CudaConvolve {}
unsafe impl Sync for CudaConvolve {}

impl CudaConvolve {
    /// `LibConvolve::new(sample_data)` (audio_matcher.rs:289) plus the sample rate the ABI needs.
    pub fn new(sample_data: Box<[SampleType]>, sr: u16) -> Result<Self, Box<dyn std::error::Error>> {
        let mut h = ptr::null_mut();
        let rc = unsafe { am_matcher_create(sample_data.as_ptr(), sample_data.len(), u32::from(sr), ptr::null(), &mut h) };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(Self { h, m: sample_data.len(), sr })
    }
}

impl Drop for CudaConvolve {
    fn drop(&mut self) {
        unsafe { am_matcher_destroy(self.h) }
    }
}

impl CorrelateAlgo<SampleType> for CudaConvolve {
    fn inverse_sample_auto_correlation(&self) -> SampleType {
        let mut v = 0f32;
        unsafe { am_inverse_sample_auto_correlation(self.h, &mut v) };
        v
    }

    fn correlate_with_sample(&self, within: &[SampleType], mode: Mode, scale: bool)
        -> Result<Vec<SampleType>, Box<dyn std::error::Error>> {
        let mode = match mode { Mode::Full => 0, Mode::Same => 1, Mode::Valid => 2 };
        let mut out = vec![0f32; unsafe { am_out_len(within.len(), self.m, mode) }];
        let mut n = 0usize;
        let rc = unsafe {
            am_correlate(self.h, within.as_ptr().cast(), within.len(), AM_FMT_F32_MONO, AM_MEM_HOST, mode,
                         c_int::from(scale), out.as_mut_ptr(), out.len(), AM_MEM_HOST, &mut n)
        };
        if rc != 0 {
            return Err(last_error());
        }
        out.truncate(n);
        Ok(out)
    }
}

/// `calc_chunks` (audio_matcher.rs:88-141) with the same parameter list, the algo fixed to `CudaConvolve`: peaks
/// sorted by start, neighbours within `distance` with a larger prominence removed.  The iterator is not collected:
/// its frames are pushed block by block (`am_stream_push`), so the GPU matches while the decoder behind the
/// iterator (mp3_reader.rs:13-66) is still producing; `m_samples.len()` is the claimed length of `with_size`
/// (mod.rs:78,83) and only has to be an upper bound.
pub fn calc_chunks<Iter: ExactSizeIterator<Item = SampleType>>(
    sr: u16,
    m_samples: Iter,
    algo_with_sample: &CudaConvolve,
    scale: bool,
    config: Config,
) -> Vec<find_peaks::Peak<SampleType>> {
    assert_eq!(sr, algo_with_sample.sr, "sample rate of the stream differs from the snippet's"); // CliError::SampleRateMismatch, mod.rs:72
    let cfg = AmConfig {
        chunk_size_s: config.chunk_size.as_secs_f64(),
        overlap_s: config.overlap_length.as_secs_f64(),
        distance_s: config.peak_config.distance.as_secs_f64(),
        prominence: config.peak_config.prominence,
        fft_log2: 0,
        max_peaks_per_chunk: 0,
        reserved: 0,
    };
    const BLOCK: usize = 1 << 22; // 16 MB of f32 per push: the library copies it into its pinned ring with a pool of threads
    let mut block: Vec<SampleType> = Vec::with_capacity(BLOCK);
    let mut peaks = vec![AmPeak::default(); 1 << 16];
    let mut n = 0usize;
    unsafe {
        assert_eq!(am_matcher_set_config(algo_with_sample.h, &cfg), 0, "{}", last_error());
        let mut session = ptr::null_mut();
        let claimed = m_samples.len().max(1);
        assert_eq!(am_stream_begin(algo_with_sample.h, claimed, AM_FMT_F32_MONO, c_int::from(scale), &mut session), 0, "{}", last_error());
        let mut it = m_samples;
        loop {
            block.clear();
            block.extend(it.by_ref().take(BLOCK));
            if block.is_empty() {
                break;
            }
            if am_stream_push(session, block.as_ptr().cast(), block.len()) != 0 {
                am_stream_abort(session);
                panic!("{}", last_error()); // the reference unwraps as well (audio_matcher.rs:122)
            }
        }
        assert_eq!(am_stream_finish(session, peaks.as_mut_ptr(), peaks.len(), &mut n), 0, "{}", last_error());
    }
    // find_peaks 0.1: `pub struct Peak<T> { pub position: Range<usize>, pub left_diff: T, pub right_diff: T,
    // pub height: Option<T>, pub prominence: Option<T> }` (crate source is not in the reference tree; downstream code
    // reads position.start and prominence.unwrap(): audio_matcher.rs:135,155, mod.rs:122,128, archive/data.rs:94)
    peaks[..n].iter().map(|p| find_peaks::Peak {
        position: p.start as usize..p.end as usize,
        left_diff: p.left_diff,
        right_diff: p.right_diff,
        height: Some(p.height),
        prominence: Some(p.prominence),
    }).collect()
}

/// The loop `for main_file in &args.within` of `matcher::run` (src/matcher/mod.rs:42-99) as ONE library call, for callers
/// that hold the decoded files (e.g. decoded on one thread per file): one peak list per file, each equal to
/// `calc_chunks` on that file.  The work of all files is queued before the first result is read back, so the upload of
/// file k+1 overlaps the matching of file k and short files keep the GPU busy.
pub fn calc_chunks_files(
    sr: u16,
    files: &[Vec<SampleType>],
    algo_with_sample: &CudaConvolve,
    scale: bool,
    config: Config,
) -> Vec<Vec<find_peaks::Peak<SampleType>>> {
    assert_eq!(sr, algo_with_sample.sr, "sample rate of the stream differs from the snippet's");
    let cfg = AmConfig {
        chunk_size_s: config.chunk_size.as_secs_f64(),
        overlap_s: config.overlap_length.as_secs_f64(),
        distance_s: config.peak_config.distance.as_secs_f64(),
        prominence: config.peak_config.prominence,
        fft_log2: 0,
        max_peaks_per_chunk: 0,
        reserved: 0,
    };
    let ptrs: Vec<*const c_void> = files.iter().map(|f| f.as_ptr().cast()).collect();
    let frames: Vec<usize> = files.iter().map(Vec::len).collect();
    let mut counts = vec![0usize; files.len()];
    let mut peaks = vec![AmPeak::default(); 1 << 16];
    unsafe {
        assert_eq!(am_matcher_set_config(algo_with_sample.h, &cfg), 0, "{}", last_error());
        assert_eq!(
            am_calc_chunks_files(algo_with_sample.h, files.len(), ptrs.as_ptr(), frames.as_ptr(), AM_FMT_F32_MONO, AM_MEM_HOST,
                                 c_int::from(scale), peaks.as_mut_ptr(), peaks.len(), counts.as_mut_ptr()),
            0, "{}", last_error()
        );
    }
    let mut rest = &peaks[..];
    counts.iter().map(|&n| {
        let (mine, tail) = rest.split_at(n);
        rest = tail;
        mine.iter().map(|p| find_peaks::Peak {
            position: p.start as usize..p.end as usize,
            left_diff: p.left_diff,
            right_diff: p.right_diff,
            height: Some(p.height),
            prominence: Some(p.prominence),
        }).collect()
    }).collect()
}
