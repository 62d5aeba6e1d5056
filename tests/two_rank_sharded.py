"""One rank of the two-process test of the C ABI's multi-GPU entry points (run by tests/test_gpu_parity.py):
  python tests/two_rank_sharded.py RANK NRANKS ID_FILE
Rank r uses GPU r.  Rank 0 creates the NCCL unique id (am_comm_get_unique_id) and writes it to ID_FILE; the others
poll for it -- no torch.distributed, no MPI.  Every rank generates the same synthetic stream, keeps only its shard
(am_shard_frames), calls am_calc_chunks_sharded and compares the result with the CPU oracle on the whole stream."""
import os
import sys
import time

rank, nranks, id_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import audio_matcher_b200 as am  # noqa: E402
from audio_matcher_b200.matcher import shard_chunks  # noqa: E402
from oracle import am_oracle as orc  # noqa: E402  (checker)

def make_comm(tag):
    """Rank 0 creates the unique id and publishes it in a file; the others poll for it."""
    path = f"{id_file}.{tag}"
    if rank == 0:
        uid = am.Comm.unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(path + ".tmp", path)
    else:
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id file")
            time.sleep(0.05)
        with open(path, "rb") as f:
            uid = f.read()
    return am.Comm(nranks, rank, uid)


sr, chunk_s, dist_s = 8000, 5.0, 12.0
pcm, snip, planted = orc.synth_case(sr, 83.0, 0.5, chunk_s=chunk_s, plant_period_s=12.5, plant_jitter_s=2.5)
x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
ref = orc.calc_chunks(x, s, sr, orc.make_config(chunk_s, len(s) / sr, dist_s, 0.13), scale=True, precision=64)
assert len(ref) >= 3, ref
conf = am.Config(chunk_size=chunk_s, peak_config=am.PeakConfig(dist_s, 0.13))
algo = am.CudaConvolve(snip, sr=sr, config=conf)
total = algo.num_chunks(len(pcm))
c0, nc = shard_chunks(total, nranks, rank)
lo, hi = algo.shard_frames(c0, nc, len(pcm))
shard = np.ascontiguousarray(pcm[lo:hi])
# second pass: a record of 1 peak per rank cannot hold a shard's peaks -> the library has to gather a second time
for tag, record in (("a", None), ("b", "1")):
    if record:
        os.environ["AM_GATHER_RECORD_PEAKS"] = record
    comm = make_comm(tag)
    os.environ.pop("AM_GATHER_RECORD_PEAKS", None)
    for samples in (shard, __import__("torch").from_numpy(shard).cuda()):      # host shard and device-resident shard
        got = algo.calc_chunks_sharded(samples, True, total_frames=len(pcm), buf_first_frame=lo, first_chunk=c0, num_chunks=nc, comm=comm)
        assert [(p.position.start, p.position.stop, p.chunk) for p in got] == [(p.start, p.end, p.chunk) for p in ref], (got, ref)
        for a, b in zip(got, ref):
            assert abs(a.height - b.height) <= 1e-4 * abs(b.height) and abs(a.prominence - b.prominence) <= 1e-4 * abs(b.prominence)
    comm.close()
algo.close()
print(f"rank {rank}: sharded ok, {len(got)} peaks")
