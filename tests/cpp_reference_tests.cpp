// cpp_reference_tests.cpp -- the reference's own unit tests (src/matcher/audio_matcher.rs:162-218, :485-517) written
// against the C++ mirror in include/audio_matcher.hpp.  Built and run by tests/test_cpp_mirror.py:
//   g++ -std=c++17 -I include tests/cpp_reference_tests.cpp -L audio_matcher_b200 -laudio_matcher_b200
// Without a CUDA device only the host-side tests run (and constructing the algorithm must fail loudly).
#include <cmath>
#include <cstdio>
#include <string>

#include "audio_matcher.hpp"

using namespace audio_matcher;

static int failures = 0;
#define ASSERT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) { std::fprintf(stderr, "%s:%d: assertion failed: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static Peak peak(std::size_t start, float prominence) {
    Peak p;
    p.start = start; p.end = start + 1; p.height = prominence; p.prominence = prominence;
    return p;
}

// overshadow_tests::test_data (:167-185): without a GPU the three peaks are written down, with one they come from
// find_peaks on [0, 0.7, 0.5, 1.0, 0.5, 0.8, 0] with min prominence 0
struct Three { Peak p1, p2, p3; };
static Three test_peaks(bool gpu) {
    if (!gpu) return {peak(3, 1.0f), peak(5, 0.3f), peak(1, 0.2f)};
    Config conf;
    conf.chunk_size_s = 7.0; conf.overlap_length_s = 0.0; conf.peak_config = {0.0, 0.0f};
    CudaConvolve one_tap({1.0f}, 1, conf);                 // correlation with a 1-tap snippet is the stream itself
    auto peaks = calc_chunks(1, {0.f, 0.7f, 0.5f, 1.0f, 0.5f, 0.8f, 0.f}, one_tap, true, conf);
    ASSERT(peaks.size() == 3);
    if (peaks.size() != 3) return {peak(3, 1.0f), peak(5, 0.3f), peak(1, 0.2f)};
    // the same stream as two "files" of one run (the loop over args.within, src/matcher/mod.rs:42) plus an empty one
    const std::vector<float> y = {0.f, 0.7f, 0.5f, 1.0f, 0.5f, 0.8f, 0.f};
    const auto per_file = calc_chunks_files(1, {y, {}, y}, one_tap, true, conf);
    ASSERT(per_file.size() == 3 && per_file[0].size() == 3 && per_file[1].empty() && per_file[2].size() == 3);
    if (per_file.size() == 3 && per_file[2].size() == 3)
        for (int i = 0; i < 3; ++i) ASSERT(per_file[2][i].start == peaks[i].start && *per_file[2][i].prominence == *peaks[i].prominence);
    // calc_chunks sorts by start: 1, 3, 5
    const Peak p3 = peaks[0], p1 = peaks[1], p2 = peaks[2];
    ASSERT(p3.start == 1 && std::fabs(*p3.prominence - 0.2f) < 1e-6f);
    ASSERT(p2.start == 5 && std::fabs(*p2.prominence - 0.3f) < 1e-6f);
    ASSERT(p1.start == 3 && std::fabs(*p1.prominence - 1.0f) < 1e-6f);
    return {p1, p2, p3};
}

static void distance_dropoff(const Three &t) {                 // :187-197
    ASSERT(is_overshadowed(t.p3, &t.p1, 1, 3.0));
    ASSERT(!is_overshadowed(t.p3, &t.p1, 1, 2.0));
    ASSERT(is_overshadowed(t.p2, &t.p1, 1, 3.0));
    ASSERT(!is_overshadowed(t.p2, &t.p1, 1, 2.0));
}
static void not_overshadowed_by_none(const Three &t) {         // :199-207
    ASSERT(!is_overshadowed(t.p1, nullptr, 1, 6.0));
    ASSERT(!is_overshadowed(t.p2, nullptr, 1, 6.0));
    ASSERT(!is_overshadowed(t.p3, nullptr, 1, 6.0));
}
static void true_peak_not_overshadowed(const Three &t) {       // :209-217
    ASSERT(!is_overshadowed(t.p1, &t.p2, 1, 6.0));
    ASSERT(!is_overshadowed(t.p1, &t.p3, 1, 6.0));
}

// my_correlate_same_fftcorrelate (:489-517): Valid, unscaled, stream -10..9, snippet [1, 2, 3]
static void correlate_kat() {
    const std::vector<float> sample = {1.f, 2.f, 3.f};
    CudaConvolve algo(sample, 1);
    const auto out = algo.correlate_with_sample(test_data(-10, 10), Mode::Valid, false);
    ASSERT(out.size() == 18);
    for (std::size_t k = 0; k < out.size() && k < 18; ++k) ASSERT(std::fabs(out[k] - (float)(-52 + 6 * (int)k)) < 1.2e-5f);
    ASSERT(std::fabs(algo.inverse_sample_auto_correlation() - 1.0f / 14.0f) < 1e-8f);
    auto scaled = algo.correlate_with_sample(test_data(-10, 10), Mode::Valid, true);
    auto by_hand = out;
    algo.scale(by_hand);                                       // the trait's default method (:73-75)
    for (std::size_t k = 0; k < out.size(); ++k) ASSERT(std::fabs(scaled[k] - by_hand[k]) < 1e-6f);
    ASSERT(algo.correlate_with_sample(test_data(-10, 10), Mode::Full, false).size() == 22);
    ASSERT(algo.correlate_with_sample(test_data(-10, 10), Mode::Same, false).size() == 20);
}

// output side (src/matcher/mod.rs:110-129, src/archive/data.rs:87-107); same expectations as tests/test_host_logic.py
static void labels_and_offsets() {
    std::vector<Peak> peaks = {peak(21 * 100, 0.9f), peak(1003 * 100, 0.5f), peak(4000 * 100, 0.7f)};
    const auto labels = timelabel_from_peaks(peaks, 100);
    ASSERT(labels.size() == 2);
    ASSERT(labels[0].start == 28.0 && labels[0].end == 1003.0 && labels[0].name == "Segment 1");
    ASSERT(labels[1].start == 1010.0 && labels[1].end == 4000.0 && labels[1].name == "Segment 2");
    ASSERT(labels[0].line() == "28.000000\t1003.000000\tSegment 1");
    ASSERT(offset_lines(peaks, 100)[1] == "Offset 2: 00:16:43 with prominence 0.5");
    ASSERT(offset_lines({peak(100, 0.98765432f)}, 100)[0] == "Offset 1: 00:00:01 with prominence 0.9876543");   // Rust `{}` on f32
    ASSERT(offset_lines({peak(100, 1.0f)}, 100)[0] == "Offset 1: 00:00:01 with prominence 1");
    ASSERT(offset_lines({}, 100).size() == 1 && offset_lines({}, 100)[0] == "no offsets found");
    ASSERT(timelabel_from_peaks({peak(5, 1.f)}, 1).empty());
}

int main() {
    const bool gpu = am_device_count() > 0;
    labels_and_offsets();
    if (!gpu) {
        bool threw = false;
        try {
            CudaConvolve algo({1.f, 2.f, 3.f}, 1);
        } catch (const Error &e) {
            threw = e.status != AM_OK && std::string(e.what()).size() > 0;
        }
        ASSERT(threw);                                         // no CPU fallback: fails loudly
    } else {
        correlate_kat();
    }
    const Three t = test_peaks(gpu);
    distance_dropoff(t);
    not_overshadowed_by_none(t);
    true_peak_not_overshadowed(t);
    if (failures) return 1;
    std::printf(gpu ? "cpp reference tests ok\n" : "cpp reference tests ok (host only)\n");
    return 0;
}
