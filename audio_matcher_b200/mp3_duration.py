"""mp3_duration -- the length claim of the reference's matcher and its tag cache.

Host-only mirror of `mp3_duration` (src/matcher/mp3_reader.rs:68-108).  The reference asks, in this order,
  1. the ID3 `TLEN` frame, which IT reads and writes as WHOLE SECONDS (src/worker/tagger.rs:176-178, :193:
     `Duration::from_secs(tag.duration())`, `set_duration(value.as_secs() as u32)`),
  2. the `mp3-duration` crate (a walk over the MPEG frame headers),
  3. a full decode, summing `frame.data.len() / (channels * sample_rate)` per frame, after which it stores the
     result -- truncated to seconds -- in `TLEN` and saves the tag (`:100-106`).
The value feeds `overlap_length = s_duration` for the snippet (`audio_matcher.rs:41`) and the `with_size` length
claim `samples = (m_duration * sr) as usize` for the stream (`mod.rs:77-83`).  Because step 3 caches whole seconds,
the SECOND run on the same snippet sees e.g. 7 s for a 7.34 s file: `ov = 308,700 < m - 1`, the per-boundary gap
that `calc_chunks` reproduces (tests/test_gpu_parity.py, oracle golden cases).

Steps 2 and 3 are one frame-header walk here: every MPEG audio frame carries a fixed number of samples per
channel (384 / 1152 / 576 by layer and version), so `sum(samples_per_frame / sample_rate)` is what the decode
sums (an `Info`/`Xing` header frame is counted like any other frame).

Pinned on the reference's own fixture `res/id3test.mp3` (tests/golden/mp3_duration.json, made by
tests/golden/make_mp3_golden.py): the tag reading is the 7 s that `tagger.rs:791` asserts, and the frame walk finds
281 frames = 323,712 samples = 7.34 s, i.e. the `as_secs() == 7` of `short_mp3_duration` (mp3_reader.rs:112-121; that
test names `res/local/Interlude.mp3`, which is absent, but SURVEY.md 8a gives the same 323,712 samples for it).  What
stays unpinned is the `mp3-duration` crate's treatment of VBR headers (step 2), which no reference test exercises.
"""
from __future__ import annotations

import os
import struct
from typing import List, Optional, Tuple


class NoFile(FileNotFoundError):
    """CliError::NoFile (src/matcher/errors.rs)."""


class NoMp3(ValueError):
    """CliError::NoMp3: no MPEG audio frame found."""


class ID3Error(ValueError):
    """CliError::ID3: the tag could not be rewritten."""


# ---------------------------------------------------------------------------------------------- MPEG frames
_BITRATES = {  # kbit/s by (version_is_1, layer)
    (True, 1): [0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448],
    (True, 2): [0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384],
    (True, 3): [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320],
    (False, 1): [0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256],
    (False, 2): [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160],
    (False, 3): [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160],
}
_RATES = {3: [44100, 48000, 32000], 2: [22050, 24000, 16000], 0: [11025, 12000, 8000]}  # by version bits


def parse_frame_header(b: bytes) -> Optional[Tuple[int, int, int, int]]:
    """(frame_bytes, samples_per_channel, sample_rate, channels) of the MPEG audio frame header in b[:4], or None."""
    if len(b) < 4 or b[0] != 0xFF or (b[1] & 0xE0) != 0xE0:
        return None
    version = (b[1] >> 3) & 3           # 3 = MPEG1, 2 = MPEG2, 0 = MPEG2.5, 1 = reserved
    layer = 4 - ((b[1] >> 1) & 3)       # 1, 2, 3 (4 = reserved)
    br_idx, sr_idx, pad = b[2] >> 4, (b[2] >> 2) & 3, (b[2] >> 1) & 1
    if version == 1 or layer == 4 or br_idx in (0, 15) or sr_idx == 3:
        return None                     # reserved values and free-format streams are not walked
    v1 = version == 3
    bitrate = _BITRATES[(v1, layer)][br_idx] * 1000
    rate = _RATES[version][sr_idx]
    if layer == 1:
        size, samples = (12 * bitrate // rate + pad) * 4, 384
    elif layer == 2 or v1:
        size, samples = 144 * bitrate // rate + pad, 1152
    else:
        size, samples = 72 * bitrate // rate + pad, 576
    channels = 1 if (b[3] >> 6) == 3 else 2
    return size, samples, rate, channels


def _synchsafe(b: bytes) -> int:
    return (b[0] & 0x7F) << 21 | (b[1] & 0x7F) << 14 | (b[2] & 0x7F) << 7 | (b[3] & 0x7F)


def _to_synchsafe(n: int) -> bytes:
    return bytes([(n >> 21) & 0x7F, (n >> 14) & 0x7F, (n >> 7) & 0x7F, n & 0x7F])


def _id3v2_extent(data: bytes) -> int:
    """Bytes occupied by a leading ID3v2 tag (header + body + footer), 0 if there is none."""
    if len(data) < 10 or data[:3] != b"ID3" or data[3] == 0xFF or data[4] == 0xFF or any(x & 0x80 for x in data[6:10]):
        return 0
    return 10 + _synchsafe(data[6:10]) + (10 if data[3] == 4 and data[5] & 0x10 else 0)


def frame_walk(data: bytes) -> Tuple[float, int, int]:
    """Walk the MPEG audio frames of an MP3 image: (seconds, frames, sample_rate of the first frame).

    A candidate header only counts if the next frame header follows it (or the data ends inside / right after
    it), the usual guard against sync words inside tags or payload."""
    pos = _id3v2_extent(data)
    end = len(data)
    if end >= 128 and data[end - 128:end - 125] == b"TAG":     # ID3v1 trailer
        end -= 128
    seconds, frames, first_rate = 0.0, 0, 0
    while pos + 4 <= end:
        h = parse_frame_header(data[pos:pos + 4])
        if h is None:
            pos += 1
            continue
        size, samples, rate, _ = h
        nxt = pos + size
        if frames == 0 and nxt + 4 <= end and parse_frame_header(data[nxt:nxt + 4]) is None:
            pos += 1                                            # false sync before the first real frame
            continue
        seconds += samples / rate
        frames += 1
        first_rate = first_rate or rate
        pos = nxt
    if frames == 0:
        raise NoMp3("no MPEG audio frame found")
    return seconds, frames, first_rate


# ---------------------------------------------------------------------------------------------- ID3v2 TLEN
def _frames_of(tag_body: bytes, major: int) -> List[Tuple[bytes, bytes, bytes]]:
    """[(id, flags, payload)] of an ID3v2.3 / v2.4 tag body (padding dropped)."""
    out, pos = [], 0
    while pos + 10 <= len(tag_body):
        fid = tag_body[pos:pos + 4]
        if fid[0] == 0:
            break                                               # padding
        size = _synchsafe(tag_body[pos + 4:pos + 8]) if major == 4 else struct.unpack(">I", tag_body[pos + 4:pos + 8])[0]
        if not all(48 <= c <= 57 or 65 <= c <= 90 for c in fid) or pos + 10 + size > len(tag_body):
            raise ID3Error("malformed ID3 frame")
        out.append((fid, tag_body[pos + 8:pos + 10], tag_body[pos + 10:pos + 10 + size]))
        pos += 10 + size
    return out


def _split_tag(data: bytes):
    """(major, flags, body_without_extended_header, extent) or None when there is no usable ID3v2.3/2.4 tag."""
    extent = _id3v2_extent(data)
    if not extent:
        return None
    major, flags = data[3], data[5]
    if major not in (3, 4) or flags & 0x80:                     # v2.2 and unsynchronised tags: read-only / unsupported
        return major, flags, None, extent
    body = data[10:10 + _synchsafe(data[6:10])]
    if flags & 0x40:                                            # extended header
        if len(body) < 4:
            return major, flags, None, extent
        skip = _synchsafe(body[:4]) if major == 4 else 4 + struct.unpack(">I", body[:4])[0]
        body = body[skip:]
    return major, flags, body, extent


def _decode_text(payload: bytes) -> str:
    if not payload:
        return ""
    enc, raw = payload[0], payload[1:]
    codec = {0: "latin-1", 1: "utf-16", 2: "utf-16-be", 3: "utf-8"}.get(enc, "latin-1")
    return raw.decode(codec, errors="replace").split("\x00")[0].strip()


def read_tlen_seconds(data: bytes) -> Optional[int]:
    """The `TLEN` value as the reference interprets it: whole seconds (tagger.rs:176-178)."""
    t = _split_tag(data)
    if t is None or t[2] is None:
        return None
    try:
        for fid, _, payload in _frames_of(t[2], t[0]):
            if fid == b"TLEN":
                txt = _decode_text(payload)
                return int(txt) if txt.isdigit() else None
    except ID3Error:
        return None
    return None


def with_tlen_seconds(data: bytes, seconds: int) -> bytes:
    """The file image with `TLEN = seconds` (tagger.rs:193 `set_duration(value.as_secs() as u32)`): the frame is
    replaced or appended in an existing ID3v2.3/2.4 tag (other frames byte-for-byte, padding reused when it
    fits), or a new ID3v2.4 tag is put in front."""
    new_frame_payload = b"\x00" + str(int(seconds)).encode("ascii")
    t = _split_tag(data)
    if t is not None and t[2] is None:
        raise ID3Error("unsupported ID3v2 tag (v2.2 or unsynchronised)")
    if t is None:
        major, frames, extent = 4, [], 0
    else:
        major, _, body, extent = t
        frames = [f for f in _frames_of(body, major) if f[0] != b"TLEN"]
    frames.append((b"TLEN", b"\x00\x00", new_frame_payload))
    body = b"".join(fid + (_to_synchsafe(len(p)) if major == 4 else struct.pack(">I", len(p))) + fl + p for fid, fl, p in frames)
    old_room = extent - 10 if extent else 0
    room = old_room if len(body) <= old_room and not (t and t[1] & 0x50) else len(body) + 256
    header = b"ID3" + bytes([major, 0, 0]) + _to_synchsafe(room)
    return header + body + b"\x00" * (room - len(body)) + data[extent:]


# ---------------------------------------------------------------------------------------------- the function
def mp3_duration(path, use_parallel: bool = False, header_walk_ok: bool = True, cache: bool = True) -> float:
    """mp3_reader.rs:68-108, three steps:
      1. the tag's length if it has one (whole seconds, tagger.rs:176-178)                                   :71-75
      2. else the `mp3-duration` crate's walk over the frame headers -- NOTHING is written                   :76-79
      3. only if both fail: the sum over the decoded frames, which is then cached in the tag as TLEN = whole
         seconds (so every later call answers step 1 with the truncated value)                               :80-106
    This mirror has one frame walk (`frame_walk`) standing in for both the crate (step 2) and the decoder sum
    (step 3), which count the same samples on a well-formed stream.  An ordinary MP3 therefore returns its exact
    duration and is never modified.  `header_walk_ok=False` reproduces the case where the crate gives up and the
    reference falls through to step 3, the only path that rewrites the input file (source of the `ov < m - 1`
    quirk); `cache=False` suppresses even that write.
    `use_parallel` is accepted for signature parity (the reference's rayon variant computes the same sum)."""
    del use_parallel
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise NoFile(str(path)) from e
    tagged = read_tlen_seconds(data)
    if tagged is not None:
        return float(tagged)
    seconds, _, _ = frame_walk(data)
    if header_walk_ok:
        return seconds
    if cache:
        try:
            image = with_tlen_seconds(data, int(seconds))
            tmp = f"{path}.tmp-tlen"
            with open(tmp, "wb") as f:
                f.write(image)
            os.replace(tmp, path)
        except OSError as e:
            raise ID3Error(str(path)) from e
    return seconds


def claimed_samples(duration_s: float, sr: int) -> int:
    """`(m_duration.as_secs_f64() * sr as f64) as usize` (src/matcher/mod.rs:78): the `with_size` length claim."""
    return int(duration_s * sr)
