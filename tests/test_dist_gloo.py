"""world_size-2 gloo test of the multi-GPU host path: chunk-range sharding, all-gather of the per-rank
peak candidates, global sort + neighbour filter.  The per-rank candidates come from the oracle here (no
GPU in this container); on the GPU box the same gather/merge code runs over NCCL (tests/test_gpu_parity.py)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import audio_matcher_b200 as am
    from audio_matcher_b200.matcher import gather_peaks, merge_peaks, shard_chunks, shard_frames, Config
    from oracle import am_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sr = 8000
        pcm, snip, _ = orc.synth_case(sr, 60.0, 0.5, chunk_s=5.0, plant_period_s=12.5, plant_jitter_s=2.5)
        x, s = orc.pcm16_to_f32(pcm), orc.pcm16_to_f32(snip)
        cfg = orc.make_config(5.0, 0.5, 2.0, 0.13)
        conf = Config(chunk_size=5.0, peak_config=am.PeakConfig(2.0, 0.13))
        total = orc.num_chunks(len(x), sr, cfg)
        c0, nc = shard_chunks(total, world, rank)
        lo, hi = shard_frames(c0, nc, len(x), sr, conf, len(s))
        # the rank only touches its own frames (+ halo): zero everything else to prove it
        shard = x.copy()
        shard[:lo] = 0
        shard[hi:] = 0
        local = orc.calc_chunks(shard, s, sr, cfg, first_chunk=c0, n_chunks=nc, final_filter=False)
        local = [am.Peak(range(p.start, p.end), p.height, p.prominence, p.left_diff, p.right_diff, p.chunk) for p in local]
        merged = merge_peaks(gather_peaks(local), sr, 2.0)
        whole = orc.calc_chunks(x, s, sr, cfg)
        ok = [(p.position.start, p.height, p.prominence) for p in merged] == [(p.start, p.height, p.prominence) for p in whole]
        q.put((rank, ok, len(merged), len(local)))
    finally:
        dist.destroy_process_group()


def test_sharded_gather_merge_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == res[1][2] > 0                   # every rank holds the same merged result
    assert all(nloc > 0 for _, _, _, nloc in res)
