"""Output side of the path: what matcher::run does with the peaks calc_chunks returns.

  print_offsets         src/matcher/mod.rs:110-125
  timelabel_from_peaks  src/archive/data.rs:87-107  (start + delay .. next start, "Segment i")
  write_labels          the Audacity label track matcher::run writes (src/matcher/mod.rs:92-99);
                        the private `audacity` crate's TimeLabel::write is not in the reference tree,
                        so the file format is Audacity's documented one: start<TAB>end<TAB>name
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
from pathlib import Path
from typing import Iterable, Sequence


@dataclass
class TimeLabel:
    start: float          # seconds
    end: float
    name: str | None

    def line(self) -> str:
        return f"{self.start:.6f}\t{self.end:.6f}\t{self.name or ''}"


def start_as_duration(peak, sr: int) -> float:
    """src/matcher/mod.rs:127-129"""
    return peak.position.start / float(sr)


def timelabel_from_peaks(peaks: Iterable, sr: int, delay_start: float = 7.0, name_pattern: str = "Segment #"):
    """src/archive/data.rs:87-107: consecutive peak pairs -> labels starting delay_start after a peak
    and ending at the next one, numbered from 1 ('#' in the pattern is replaced)."""
    starts = [start_as_duration(p, sr) for p in peaks]
    return [TimeLabel(a + delay_start, b, name_pattern.replace("#", str(i)))
            for i, (a, b) in enumerate(zip(starts, starts[1:]), start=1)]


def offset_lines(peaks: Sequence, sr: int) -> list[str]:
    """The log lines of print_offsets (src/matcher/mod.rs:110-125)."""
    if not peaks:
        return ["no offsets found"]
    out = []
    for i, p in enumerate(peaks, start=1):
        secs = int(start_as_duration(p, sr))
        # Rust's `{}` on an f32: the shortest decimal that round-trips, never in exponent form (0.98765433, not the
        # 0.9876543283462524 of the widened double).  hours()/minutes()/seconds() come from the private `common`
        # crate: read here as hh of the total, mm and ss within the hour / minute (unpinned).
        prom = np.format_float_positional(np.float32(p.prominence), unique=True, trim="-")
        out.append(f"Offset {i}: {secs // 3600:0>2}:{(secs // 60) % 60:0>2}:{secs % 60:0>2} with prominence {prom}")
    return out


def print_offsets(peaks: Sequence, sr: int, log=print) -> None:
    for line in offset_lines(peaks, sr):
        log(line)


def write_labels(labels: Iterable[TimeLabel], path, dry_run: bool = False) -> str:
    text = "".join(l.line() + "\n" for l in labels)
    if not dry_run:
        Path(path).write_text(text)
    return text
