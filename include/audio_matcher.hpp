// audio_matcher.hpp -- C++17 host-side mirror of the reference's matcher interface over the C ABI
// (include/audio_matcher.h).  Header-only; link against libaudio_matcher_b200.so.
//
// The reference is a Rust crate; a maintainer binds the C ABI from Rust (ffi/cuda_convolve.rs).
// This header gives the same surface to C++ callers and lets the parity tests read like the reference's own:
//
//   reference (src/matcher/audio_matcher.rs)                      here
//   ------------------------------------------------------------  ------------------------------------------
//   enum Mode { Full, Same, Valid }                        :54-59  audio_matcher::Mode
//   struct PeakConfig / Config                             :24-53  audio_matcher::PeakConfig / Config
//   trait CorrelateAlgo<R>                                 :65-76  audio_matcher::CorrelateAlgo (abstract)
//     fn inverse_sample_auto_correlation(&self) -> R                 float inverse_sample_auto_correlation() const
//     fn correlate_with_sample(&self, within, mode, scale)           std::vector<float> correlate_with_sample(...)
//     fn scale(&self, data: &mut [R])                                void scale(std::vector<float>&) const
//   LibConvolve::new(Box<[f32]>)                           :289    audio_matcher::CudaConvolve(snippet, sr, config)
//   fn calc_chunks(sr, m_samples, algo, scale, config)     :88-141 audio_matcher::calc_chunks(sr, samples, algo, scale, config)
//   fn is_overshadowed(element, other, sr, max_distance)   :143-160 audio_matcher::is_overshadowed(...)
//   fn test_data(range)                                    :481-483 audio_matcher::test_data(from, to)
//   print_offsets / start_as_duration      src/matcher/mod.rs:110-129 audio_matcher::offset_lines / start_as_duration
//   timelabel_from_peaks                 src/archive/data.rs:87-107 audio_matcher::timelabel_from_peaks, TimeLabel
//
// Error behaviour: the reference returns Result<_, Box<dyn Error>> from correlate_with_sample and panics (unwrap,
// :122) inside calc_chunks; here every failing C-ABI call throws audio_matcher::Error carrying am_last_error().
// There is no CPU fallback: without a CUDA device construction throws.
#pragma once

#include <charconv>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "audio_matcher.h"

namespace audio_matcher {

struct Error : std::runtime_error {
    am_status status;
    Error(am_status st, const std::string &what) : std::runtime_error(what), status(st) {}
};
inline void check(am_status st) {
    if (st != AM_OK) throw Error(st, am_last_error());
}

enum class Mode { Full = AM_MODE_FULL, Same = AM_MODE_SAME, Valid = AM_MODE_VALID };

struct PeakConfig {
    double distance_s = 480.0;   // args.rs:75
    float prominence = 0.13f;    // args.rs:76 / 100 (audio_matcher.rs:44)
};
struct Config {
    double chunk_size_s = 60.0;  // args.rs:71
    double overlap_length_s = -1.0;   // < 0: the snippet duration m / sr (audio_matcher.rs:41)
    PeakConfig peak_config;
    uint32_t fft_log2 = 0;
    am_config raw() const {
        am_config c;
        am_config_default(&c);
        c.chunk_size_s = chunk_size_s;
        c.overlap_s = overlap_length_s;
        c.distance_s = peak_config.distance_s;
        c.prominence = peak_config.prominence;
        c.fft_log2 = fft_log2;
        return c;
    }
};

// find_peaks::Peak<f32> as the reference's downstream code reads it (position.start, prominence; mod.rs:122,128)
struct Peak {
    std::size_t start = 0, end = 0;       // position: Range<usize>
    float height = 0.f;
    std::optional<float> prominence;
    float left_diff = 0.f, right_diff = 0.f;
    uint32_t chunk = 0;
    am_peak raw() const {
        am_peak p{};
        p.start = start; p.end = end; p.height = height; p.prominence = prominence.value_or(0.f);
        p.left_diff = left_diff; p.right_diff = right_diff; p.chunk = chunk;
        return p;
    }
    static Peak from(const am_peak &p) {
        Peak q;
        q.start = (std::size_t)p.start; q.end = (std::size_t)p.end; q.height = p.height; q.prominence = p.prominence;
        q.left_diff = p.left_diff; q.right_diff = p.right_diff; q.chunk = p.chunk;
        return q;
    }
};

// trait CorrelateAlgo<f32>, audio_matcher.rs:65-76
class CorrelateAlgo {
public:
    virtual ~CorrelateAlgo() = default;
    virtual float inverse_sample_auto_correlation() const = 0;
    virtual std::vector<float> correlate_with_sample(const std::vector<float> &within, Mode mode, bool scale) const = 0;
    // default method of the trait (:73-75)
    virtual void scale(std::vector<float> &data) const {
        const float f = inverse_sample_auto_correlation();
        for (float &x : data) x *= f;
    }
};

// The GPU implementation behind the C ABI (the reference's LibConvolve seam, :282-344)
class CudaConvolve final : public CorrelateAlgo {
public:
    CudaConvolve(const std::vector<float> &snippet, uint32_t sr, const Config &config = Config()) : sr_(sr), m_(snippet.size()) {
        if (am_abi_version() != AM_ABI_VERSION) throw Error(AM_ERR_INVALID, "audio_matcher ABI version mismatch");
        const am_config c = config.raw();
        check(am_matcher_create(snippet.data(), snippet.size(), sr, &c, &h_));
    }
    CudaConvolve(const CudaConvolve &) = delete;
    CudaConvolve &operator=(const CudaConvolve &) = delete;
    ~CudaConvolve() override { am_matcher_destroy(h_); }

    float inverse_sample_auto_correlation() const override {
        float v = 0.f;
        check(am_inverse_sample_auto_correlation(h_, &v));
        return v;
    }
    std::vector<float> correlate_with_sample(const std::vector<float> &within, Mode mode, bool scale) const override {
        std::vector<float> out(am_out_len(within.size(), m_, (am_mode)mode));
        std::size_t n = 0;
        check(am_correlate(h_, within.data(), within.size(), AM_FMT_F32_MONO, AM_MEM_HOST, (am_mode)mode, scale ? 1 : 0,
                           out.data(), out.size(), AM_MEM_HOST, &n));
        out.resize(n);
        return out;
    }
    am_matcher *handle() const { return h_; }
    uint32_t sample_rate() const { return sr_; }
    std::size_t snippet_len() const { return m_; }

private:
    am_matcher *h_ = nullptr;
    uint32_t sr_;
    std::size_t m_;
};

// calc_chunks, audio_matcher.rs:88-141: peaks of the whole stream, sorted by start, neighbours filtered.
// `config` replaces the one the matcher was created with (the reference passes it per call).
inline std::vector<Peak> calc_chunks(uint32_t sr, const std::vector<float> &m_samples, const CudaConvolve &algo, bool scale,
                                     const Config &config) {
    if (sr != algo.sample_rate()) throw Error(AM_ERR_INVALID, "sample rate differs from the matcher's");
    const am_config c = config.raw();
    check(am_matcher_set_config(algo.handle(), &c));
    const std::size_t chunks = am_num_chunks(algo.handle(), m_samples.size());
    std::vector<am_peak> raw(chunks * 64 + 64);
    std::size_t n = 0;
    check(am_calc_chunks(algo.handle(), m_samples.data(), m_samples.size(), AM_FMT_F32_MONO, AM_MEM_HOST, scale ? 1 : 0,
                         raw.data(), raw.size(), &n));
    std::vector<Peak> out;
    out.reserve(n);
    for (std::size_t i = 0; i < n; ++i) out.push_back(Peak::from(raw[i]));
    return out;
}

// The loop over args.within of matcher::run (src/matcher/mod.rs:42-99) as one library call: one peak list per file, each
// equal to calc_chunks on that file; uploads overlap matching across files.
inline std::vector<std::vector<Peak>> calc_chunks_files(uint32_t sr, const std::vector<std::vector<float>> &files,
                                                        const CudaConvolve &algo, bool scale, const Config &config) {
    if (sr != algo.sample_rate()) throw Error(AM_ERR_INVALID, "sample rate differs from the matcher's");
    const am_config c = config.raw();
    check(am_matcher_set_config(algo.handle(), &c));
    std::vector<const void *> ptrs;
    std::vector<std::size_t> frames, counts(files.size(), 0);
    std::size_t cap = 64;
    for (const auto &f : files) {
        ptrs.push_back(f.data());
        frames.push_back(f.size());
        cap += am_num_chunks(algo.handle(), f.size()) * 64;
    }
    std::vector<am_peak> raw(cap);
    check(am_calc_chunks_files(algo.handle(), files.size(), ptrs.data(), frames.data(), AM_FMT_F32_MONO, AM_MEM_HOST, scale ? 1 : 0,
                               raw.data(), raw.size(), counts.data()));
    std::vector<std::vector<Peak>> out(files.size());
    std::size_t k = 0;
    for (std::size_t f = 0; f < files.size(); ++f)
        for (std::size_t i = 0; i < counts[f]; ++i) out[f].push_back(Peak::from(raw[k++]));
    return out;
}

// is_overshadowed, audio_matcher.rs:143-160 (`other` may be absent)
inline bool is_overshadowed(const Peak &element, const Peak *other, uint32_t sr, double max_distance_s) {
    if (!other) return false;
    const am_peak e = element.raw(), o = other->raw();
    return am_is_overshadowed(&e, &o, sr, max_distance_s) != 0;
}

// ---- output side of the path: what matcher::run does with the peaks (src/matcher/mod.rs:85-99) ----------------

// start_as_duration, src/matcher/mod.rs:127-129 (seconds)
inline double start_as_duration(const Peak &peak, uint32_t sr) { return (double)peak.start / (double)sr; }

// the log lines of print_offsets, src/matcher/mod.rs:110-125
inline std::vector<std::string> offset_lines(const std::vector<Peak> &peaks, uint32_t sr) {
    std::vector<std::string> out;
    if (peaks.empty()) out.push_back("no offsets found");
    for (std::size_t i = 0; i < peaks.size(); ++i) {
        const unsigned long long secs = (unsigned long long)start_as_duration(peaks[i], sr);
        // Rust's `{}` on an f32 prints the shortest decimal that round-trips, never in exponent form
        char prom[64];
        const auto r = std::to_chars(prom, prom + sizeof prom - 1, peaks[i].prominence.value_or(0.f), std::chars_format::fixed);
        *r.ptr = 0;
        char buf[224];
        std::snprintf(buf, sizeof buf, "Offset %zu: %02llu:%02llu:%02llu with prominence %s", i + 1, secs / 3600, (secs / 60) % 60,
                      secs % 60, prom);
        out.push_back(buf);
    }
    return out;
}

// audacity TimeLabel as matcher::run writes it: start<TAB>end<TAB>name
struct TimeLabel {
    double start = 0.0, end = 0.0;
    std::string name;
    std::string line() const {
        char buf[96];
        std::snprintf(buf, sizeof buf, "%.6f\t%.6f\t", start, end);
        return std::string(buf) + name;
    }
};

// timelabel_from_peaks, src/archive/data.rs:87-107: consecutive peak pairs -> a label from delay_start after a
// peak to the next peak, numbered from 1 ('#' in the pattern is replaced by the number)
inline std::vector<TimeLabel> timelabel_from_peaks(const std::vector<Peak> &peaks, uint32_t sr, double delay_start_s = 7.0,
                                                   const std::string &name_pattern = "Segment #") {
    std::vector<TimeLabel> out;
    for (std::size_t i = 0; i + 1 < peaks.size(); ++i) {
        TimeLabel l;
        l.start = start_as_duration(peaks[i], sr) + delay_start_s;
        l.end = start_as_duration(peaks[i + 1], sr);
        l.name = name_pattern;
        for (std::size_t pos = 0; (pos = l.name.find('#', pos)) != std::string::npos;) {
            const std::string num = std::to_string(i + 1);
            l.name.replace(pos, 1, num);
            pos += num.size();
        }
        out.push_back(l);
    }
    return out;
}

// test_data, audio_matcher.rs:481-483
inline std::vector<float> test_data(long from, long to) {
    std::vector<float> v;
    for (long i = from; i < to; ++i) v.push_back((float)i);
    return v;
}

}  // namespace audio_matcher
