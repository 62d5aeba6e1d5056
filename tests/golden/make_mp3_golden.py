"""Golden values for the mp3_duration mirror from the reference's own fixture res/id3test.mp3 (run in the build
container, where /root/reference is mounted; the fixture itself is not copied).

Reference expectations on that file: `tag.get::<Length>() == Some(Duration::from_secs(7))`
(src/worker/tagger.rs:791) and, for the same audio under its other name, `mp3_duration(..).as_secs() == 7`
(src/matcher/mp3_reader.rs:112-121)."""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
md = importlib.import_module("audio_matcher_b200.mp3_duration")

SRC = "/root/reference/res/id3test.mp3"
data = open(SRC, "rb").read()
t = md._split_tag(data)
seconds, frames, rate = md.frame_walk(data)
gold = {
    "source": "res/id3test.mp3 of the reference", "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(),
    "id3_major": t[0], "id3_extent": md._id3v2_extent(data), "frame_ids": [f[0].decode() for f in md._frames_of(t[2], t[0])],
    "tlen_seconds": md.read_tlen_seconds(data), "frames": frames, "sample_rate": rate, "seconds": seconds,
    "samples": frames * 1152, "reference_expectation": {"tagger.rs:791 Length": 7, "mp3_reader.rs:112-121 as_secs": 7},
}
with open(os.path.join(ROOT, "tests", "golden", "mp3_duration.json"), "w") as f:
    json.dump(gold, f, indent=1)
print(gold)
