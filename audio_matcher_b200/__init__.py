"""audio_matcher_b200 -- B200-native (sm_100a) drop-in for audio-matcher's hot path.

Host-side mirror of the reference's matcher interface (src/matcher/audio_matcher.rs) over
the C ABI in include/audio_matcher.h.  Only the snippet-vs-stream correlation + peak picking
path lives here; decoding, tagging and the CLI stay with the reference.
"""
from .matcher import (Comm, Config, CudaConvolve, Mode, Peak, PeakConfig, StreamSession, calc_chunks,  # noqa: F401
                      calc_chunks_files, calc_chunks_sharded, calc_chunks_streamed, comm_init_from_torch, is_overshadowed, merge_peaks, test_data)
from .labels import TimeLabel, print_offsets, timelabel_from_peaks, write_labels  # noqa: F401
from .mp3_duration import claimed_samples, mp3_duration  # noqa: F401

__all__ = ["Comm", "comm_init_from_torch", "StreamSession", "calc_chunks_streamed", "Config", "CudaConvolve", "Mode", "Peak", "PeakConfig", "calc_chunks", "calc_chunks_files", "calc_chunks_sharded",
           "is_overshadowed", "merge_peaks", "test_data", "TimeLabel", "print_offsets", "timelabel_from_peaks",
           "write_labels", "mp3_duration", "claimed_samples"]
