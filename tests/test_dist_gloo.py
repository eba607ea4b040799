"""world_size-2 gloo test of the N>1 host logic: batch sharding covers the batch exactly once and
the cross-rank reductions (MAX of timings, SUM of processed samples / NLL partials) agree."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sr_wavenet_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = shard.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    s, e = shard.shard_batch(B, r, w)
    per_utt = np.arange(B, dtype=np.float64) + 1.0          # stand-in for per-utterance NLL sums
    fake_ms = 10.0 + 5.0 * rank
    shard.barrier()
    tmax, = shard.reduce_scalars([fake_ms], "max")
    units, nll = shard.reduce_scalars([float(e - s), float(per_utt[s:e].sum())], "sum")
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([tmax, units, nll, s, e]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions(tmp_path):
    B, world = 7, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(world)]
    for row in rows:
        assert row[0] == 15.0                      # max over ranks
        assert row[1] == B                         # every utterance processed exactly once
        assert row[2] == B * (B + 1) / 2
    assert rows[0][4] == rows[1][3] and rows[0][3] == 0 and rows[1][4] == B
