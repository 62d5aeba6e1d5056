"""mp3_duration mirror (src/matcher/mp3_reader.rs:68-108): frame walk, TLEN-as-seconds tag cache and
the length claim, on synthetic MPEG frame streams (the reference's own fixture res/local/Interlude.mp3 is absent)."""
import importlib
import json
import os
import struct

import pytest

md = importlib.import_module("audio_matcher_b200.mp3_duration")   # the package also exports the function of that name


def _frame(bitrate_idx=9, sr_idx=0, pad=0, version=3, layer_bits=1, mono=False):
    """One MPEG audio frame with a zero payload (default: MPEG1 Layer III, 128 kbit/s, 44.1 kHz, stereo)."""
    b1 = 0xE0 | (version << 3) | (layer_bits << 1) | 1
    b2 = (bitrate_idx << 4) | (sr_idx << 2) | (pad << 1)
    b3 = 0xC0 if mono else 0x00
    hdr = bytes([0xFF, b1, b2, b3])
    size = md.parse_frame_header(hdr)[0]
    return hdr + b"\x00" * (size - 4)


def _stream(n, **kw):
    # 44.1 kHz CBR needs a padded frame now and then; alternate to exercise both sizes
    return b"".join(_frame(pad=i % 2, **kw) for i in range(n))


def test_frame_header_tables():
    assert md.parse_frame_header(_frame()[:4]) == (417, 1152, 44100, 2)
    assert md.parse_frame_header(_frame(pad=1)[:4]) == (418, 1152, 44100, 2)
    assert md.parse_frame_header(_frame(sr_idx=1, bitrate_idx=14, mono=True)[:4]) == (960, 1152, 48000, 1)   # 320 kbit/s
    assert md.parse_frame_header(_frame(version=2, bitrate_idx=8)[:4]) == (208, 576, 22050, 2)               # MPEG2 L3 64k
    assert md.parse_frame_header(_frame(layer_bits=3, bitrate_idx=4)[:4])[1] == 384                          # Layer I
    assert md.parse_frame_header(b"\xff\xfb\xf0\x00") is None and md.parse_frame_header(b"\xff\xfb\x0c\x00") is None
    assert md.parse_frame_header(b"ID3\x04") is None


def test_frame_walk_counts_samples_like_the_decoder_sum():
    n = 281                                                     # 281 * 1152 / 44100 = 7.3404 s (the Interlude-sized case)
    secs, frames, rate = md.frame_walk(_stream(n))
    assert frames == n and rate == 44100 and abs(secs - n * 1152 / 44100) < 1e-9
    # junk with a false sync word in front, ID3v1 trailer behind
    junk = b"\x00\x11\xff\xfb\x90" + b"\x22" * 20
    secs2, frames2, _ = md.frame_walk(junk + _stream(n) + b"TAG" + b"\x00" * 125)
    assert frames2 == n and abs(secs2 - secs) < 1e-9
    with pytest.raises(md.NoMp3):
        md.frame_walk(b"\x00" * 5000)


def test_tag_cache_turns_the_duration_into_whole_seconds(tmp_path):
    """An ordinary untagged MP3 answers from the frame headers and is NOT modified (mp3_reader.rs:76-79).  Only when
    the header walk of the `mp3-duration` crate fails does the reference fall through to its decoder sum and cache the
    result in the tag (:80-106) -- the quirk behind `ov < m - 1`: that call returns 7.34 s, then
    TLEN = 7 is stored and every later call answers 7 s (tagger.rs:176-178, :193)."""
    p = tmp_path / "interlude.mp3"
    original = _stream(281)
    p.write_bytes(original)
    for _ in range(2):
        assert abs(md.mp3_duration(p) - 281 * 1152 / 44100) < 1e-9 and p.read_bytes() == original
    assert abs(md.mp3_duration(p, header_walk_ok=False, cache=False) - 281 * 1152 / 44100) < 1e-9 and p.read_bytes() == original
    first = md.mp3_duration(p, header_walk_ok=False)
    assert abs(first - 281 * 1152 / 44100) < 1e-9
    data = p.read_bytes()
    assert data[:3] == b"ID3" and md.read_tlen_seconds(data) == 7
    assert md.frame_walk(data)[1] == 281                        # audio untouched behind the new tag
    assert md.mp3_duration(p) == 7.0
    sr = 44100
    m = md.frame_walk(data)[1] * 1152
    ov = round(7.0 * sr)
    assert ov == 308700 and m - ov - 1 == 15011                 # the 15,011-offset gap per chunk boundary
    assert md.claimed_samples(first, sr) == int(first * sr)


def test_existing_tag_frames_survive_the_update(tmp_path):
    def v23_frame(fid, payload):
        return fid + struct.pack(">I", len(payload)) + b"\x00\x00" + payload
    body = v23_frame(b"TIT2", b"\x00Interlude") + v23_frame(b"TLEN", b"\x00999") + v23_frame(b"TALB", b"\x03Album")
    pad = 64
    tag = b"ID3\x03\x00\x00" + md._to_synchsafe(len(body) + pad) + body + b"\x00" * pad
    audio = _stream(50)
    data = tag + audio
    assert md.read_tlen_seconds(data) == 999
    out = md.with_tlen_seconds(data, 12)
    assert md.read_tlen_seconds(out) == 12 and len(out) == len(data)      # fits into the old padding: same layout
    t = md._split_tag(out)
    ids = [f[0] for f in md._frames_of(t[2], t[0])]
    assert ids == [b"TIT2", b"TALB", b"TLEN"] and out[len(tag):] == audio
    # a tag with a length answers without touching the file
    p = tmp_path / "tagged.mp3"
    p.write_bytes(out)
    assert md.mp3_duration(p) == 12.0 and p.read_bytes() == out
    # utf-16 text frame, v2.4 synchsafe sizes
    body4 = b"TLEN" + md._to_synchsafe(7) + b"\x00\x00" + b"\x01" + "42".encode("utf-16")[:6]
    assert md.read_tlen_seconds(b"ID3\x04\x00\x00" + md._to_synchsafe(len(body4)) + body4 + audio) == 42


def test_errors(tmp_path):
    with pytest.raises(md.NoFile):
        md.mp3_duration(tmp_path / "missing.mp3")
    p = tmp_path / "not.mp3"
    p.write_bytes(b"RIFF" + b"\x00" * 4000)
    with pytest.raises(md.NoMp3):
        md.mp3_duration(p)
    unsync = b"ID3\x03\x00\x80" + md._to_synchsafe(0) + _stream(3)
    with pytest.raises(md.ID3Error):
        md.with_tlen_seconds(unsync, 1)


GOLD = os.path.join(os.path.dirname(__file__), "golden", "mp3_duration.json")
REF_FIXTURE = "/root/reference/res/id3test.mp3"


def test_golden_values_match_the_reference_expectations():
    """tests/golden/mp3_duration.json was produced from the reference's fixture res/id3test.mp3: the tag reading must be
    the 7 s the reference's tagger test expects (tagger.rs:791), the frame walk must truncate to the 7 s its
    mp3_duration test expects (mp3_reader.rs:112-121) and give the 323,712-sample snippet of SURVEY.md 8a row 5."""
    with open(GOLD) as f:
        g = json.load(f)
    assert g["tlen_seconds"] == g["reference_expectation"]["tagger.rs:791 Length"] == 7
    assert int(g["seconds"]) == g["reference_expectation"]["mp3_reader.rs:112-121 as_secs"] == 7
    assert g["frames"] == 281 and g["sample_rate"] == 44100 and g["samples"] == 323712
    assert round(7.0 * 44100) == 308700 and g["samples"] - 308700 - 1 == 15011
    assert "TLEN" in g["frame_ids"] and g["id3_major"] == 3


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="reference fixture only exists in the build container")
def test_mirror_on_the_reference_fixture(tmp_path):
    with open(GOLD) as f:
        g = json.load(f)
    data = open(REF_FIXTURE, "rb").read()
    assert len(data) == g["bytes"] and md.read_tlen_seconds(data) == g["tlen_seconds"]
    secs, frames, rate = md.frame_walk(data)
    assert (frames, rate) == (g["frames"], g["sample_rate"]) and abs(secs - g["seconds"]) < 1e-12
    p = tmp_path / "copy.mp3"
    p.write_bytes(data)
    assert md.mp3_duration(p) == 7.0 and p.read_bytes() == data          # tagged: answered from the tag, file untouched
    # drop the length from the tag -> the frames answer; only the decoder-sum fallback caches it as whole seconds
    t = md._split_tag(data)
    frames_wo = [f for f in md._frames_of(t[2], t[0]) if f[0] != b"TLEN"]
    body = b"".join(fid + struct.pack(">I", len(pl)) + fl + pl for fid, fl, pl in frames_wo)
    room = g["id3_extent"] - 10
    stripped = b"ID3\x03\x00\x00" + md._to_synchsafe(room) + body + b"\x00" * (room - len(body)) + data[g["id3_extent"]:]
    p.write_bytes(stripped)
    assert abs(md.mp3_duration(p) - g["seconds"]) < 1e-12 and p.read_bytes() == stripped      # step 2: exact, file untouched
    assert abs(md.mp3_duration(p, header_walk_ok=False) - g["seconds"]) < 1e-12                # step 3: cached afterwards
    again = p.read_bytes()
    assert md.read_tlen_seconds(again) == 7 and again[g["id3_extent"]:] == data[g["id3_extent"]:] and len(again) == len(data)
    assert md.mp3_duration(p) == 7.0
