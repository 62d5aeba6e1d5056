//! Drop-in `CorrelateAlgo` / `calc_chunks` for NilsJochem/audio-matcher backed by libaudio_matcher_b200.so.
//!
//! Put this file at `src/matcher/cuda_convolve.rs` of the reference crate, add `pub mod cuda_convolve;` to
//! `src/matcher/mod.rs`, link with `-l audio_matcher_b200`, and swap the two names in `matcher::run`
//! (`src/matcher/mod.rs:34,81`).  It binds exactly the entry points declared in `include/audio_matcher.h`.
//! Shipped as source: the build image has no Rust toolchain.
use std::{ffi::CStr, os::raw::{c_char, c_int, c_void}, ptr};

use crate::matcher::audio_matcher::{Config, CorrelateAlgo, Mode};
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

pub const AM_FMT_F32_MONO: c_int = 0;
pub const AM_FMT_I16_MONO: c_int = 1;
pub const AM_FMT_I16_STEREO: c_int = 2;
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
    fn am_calc_chunks(h: *mut AmMatcher, stream: *const c_void, frames: usize, fmt: c_int, mem: c_int, scale: c_int,
                      out: *mut AmPeak, cap: usize, n_out: *mut usize) -> c_int;
}

fn last_error() -> Box<dyn std::error::Error> {
    unsafe { CStr::from_ptr(am_last_error()) }.to_string_lossy().into_owned().into()
}

/// Replaces `LibConvolve` (src/matcher/audio_matcher.rs:282-344).
pub struct CudaConvolve {
    h: *mut AmMatcher,
    m: usize,
}
// The handle is internally locked; calc_chunks shares `&algo` across rayon workers (audio_matcher.rs:114-122).
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
        Ok(Self { h, m: sample_data.len() })
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

/// Same contract as `calc_chunks` (audio_matcher.rs:88-141): peaks sorted by start, neighbours within
/// `distance` with a larger prominence removed.  `cfg` carries the four values of `Config`/`PeakConfig`
/// (their fields are private in the reference: add getters or build `AmConfig` in `Config::from_args`).
pub fn calc_chunks_cuda(m_samples: impl ExactSizeIterator<Item = SampleType>, algo: &CudaConvolve, scale: bool,
                        cfg: AmConfig) -> Vec<find_peaks::Peak<SampleType>> {
    let stream: Vec<f32> = m_samples.collect();
    let mut peaks = vec![AmPeak::default(); 1 << 16];
    let mut n = 0usize;
    unsafe {
        assert_eq!(am_matcher_set_config(algo.h, &cfg), 0, "{}", last_error());
        let rc = am_calc_chunks(algo.h, stream.as_ptr().cast(), stream.len(), AM_FMT_F32_MONO, AM_MEM_HOST,
                                c_int::from(scale), peaks.as_mut_ptr(), peaks.len(), &mut n);
        assert_eq!(rc, 0, "{}", last_error()); // the reference unwraps as well (audio_matcher.rs:122)
    }
    peaks[..n].iter().map(|p| find_peaks::Peak {
        position: p.start as usize..p.end as usize,
        left_diff: p.left_diff,
        right_diff: p.right_diff,
        height: Some(p.height),
        prominence: Some(p.prominence),
    }).collect()
}

#[allow(dead_code)]
fn _config_from(config: &Config, chunk_size_s: f64, overlap_s: f64, distance_s: f64, prominence: f32) -> AmConfig {
    let _ = config;
    AmConfig { chunk_size_s, overlap_s, distance_s, prominence, fft_log2: 0, max_peaks_per_chunk: 0, reserved: 0 }
}
