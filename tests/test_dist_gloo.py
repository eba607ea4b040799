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


# ---- distillation step: data-parallel gradient exchange -------------------------------
def _grad_worker(rank, world, port, out_dir):
    """Each rank differentiates its batch shard with the loss normalised by the GLOBAL batch (model.py:379),
    then one all-reduce(SUM) of the flat gradient -- the scheme ParallelWaveNet.train_fast uses with NCCL."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import distill_torch as dt
    from sr_wavenet_b200 import synth
    shard.init_from_env(backend="gloo")
    case = np.load(os.path.join(out_dir, "case.npz"))
    dil, F, P = [1, 2], 2, 128
    w = {k: v.astype(np.float64) for k, v in synth.make_student_weights(dil, F, latent_channels=4, seed=3).items()}
    B = case["z"].shape[0]
    s, e = shard.shard_batch(B, rank, world)
    loss, power, _, g = dt.loss_and_grads(w, case["z"][s:e], case["truth"][s:e], case["enc"][s:e], case["tl"][s:e],
                                          dil, P, F, alpha=0.25, beta=1.0, gamma=1.0, batch_norm=B)
    names = sorted(g)
    flat = torch.from_numpy(np.concatenate([g[k].ravel() for k in names]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    lp = torch.tensor([loss, power], dtype=torch.float64)
    dist.all_reduce(lp, op=dist.ReduceOp.SUM)
    np.savez(os.path.join(out_dir, "g%d.npz" % rank), flat=flat.numpy(), lp=lp.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    from oracle import distill_torch as dt
    from sr_wavenet_b200 import synth
    rng = np.random.default_rng(1)
    B, T, P, C, M = 3, 640, 128, 4, 2
    case = dict(z=rng.logistic(0, 1, size=(B, T)), truth=synth.synthetic_audio(B, T).astype(np.float64),
                enc=rng.normal(0, 1, size=(B, T // P, C)), tl=rng.normal(0, 0.5, size=(B, T, 4 * M)))
    np.savez(os.path.join(str(tmp_path), "case.npz"), **case)
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    dil, F = [1, 2], 2
    w = {k: v.astype(np.float64) for k, v in synth.make_student_weights(dil, F, latent_channels=C, seed=3).items()}
    loss, power, _, g = dt.loss_and_grads(w, case["z"], case["truth"], case["enc"], case["tl"], dil, P, F,
                                          alpha=0.25, beta=1.0, gamma=1.0)
    full = np.concatenate([g[k].ravel() for k in sorted(g)])
    r0, r1 = (np.load(os.path.join(str(tmp_path), "g%d.npz" % r)) for r in range(2))
    np.testing.assert_allclose(r0["flat"], r1["flat"], rtol=0, atol=0)          # every rank holds the same gradient
    # the power loss is a squared norm over the whole batch tensor (model.py:371): separable per utterance, so the
    # shard sums reproduce the full-batch loss and gradient
    np.testing.assert_allclose(r0["flat"], full, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r0["lp"], [loss, power], rtol=1e-10)


def _dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    shard.init_from_env(backend="gloo")
    assert shard.is_distributed()
    rng = np.random.default_rng(100 + rank)                 # replicas start DIFFERENT (unseeded Glorot init)
    weights = torch.from_numpy(rng.normal(size=1000).astype(np.float32))
    m, v = torch.from_numpy(rng.normal(size=1000).astype(np.float32)), torch.full((1000,), float(rank))
    step = torch.tensor([float(3 + rank)])
    local_B = 3 + rank                                      # unequal shares
    gB = shard.global_batch(local_B, device="cpu")
    shard.broadcast_from_rank0([weights, m, v, step])
    # one step: bucket = [gradient | loss | power], each rank's terms divided by the GLOBAL batch
    g = np.full(1000, float(local_B)) / gB
    bucket = torch.from_numpy(np.concatenate([g, [local_B * 2.0 / gB, local_B * 5.0]]).astype(np.float32))
    shard.all_reduce_sum(bucket)
    weights -= 0.1 * bucket[:1000]
    np.save(os.path.join(out_dir, "dp%d.npy" % rank), np.concatenate([weights.numpy(), m.numpy(), v.numpy(), step.numpy(),
                                                                        bucket[-2:].numpy(), [gB]]))
    dist.destroy_process_group()


def test_two_rank_replicas_are_synchronised(tmp_path):
    """The data-parallel distillation step's host logic (ParallelWaveNet._sync_replicas / train_fast): replicas that
    start from different weights / Adam state hold rank 0's after the first-step broadcast, the loss divisor is the sum
    of unequal local batches, and one all-reduce of [gradient | loss | power] leaves identical weights everywhere."""
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(str(tmp_path / ("dp%d.npy" % r))) for r in range(world))
    np.testing.assert_array_equal(a, b)
    ref = np.random.default_rng(100).normal(size=1000).astype(np.float32)
    np.testing.assert_allclose(a[:1000], ref - np.float32(0.1), rtol=1e-6)        # summed gradient = (3 + 4) / 7 = 1
    assert a[-1] == 7 and a[3000] == 3.0 and (a[2000:3000] == 0).all()
    np.testing.assert_allclose(a[-3:-1], [2.0, 35.0], rtol=1e-6)
