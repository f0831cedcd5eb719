"""world_size-2 `gloo` test (CPU) of the N>1 host logic: contiguous trajectory shards, no data-path collective,
optional gather of ragged shards, and the data-parallel gradient averaging used by the training step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import synth
from oracle import trpx_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trphysx_b200.parallel import ShardedRollout, shard_bounds, average_flat_grad
        cfg = synth.make_config("rossler")
        sd = O.to_torch(synth.gpt2_weights(cfg, 3, stress=True))
        x0 = torch.from_numpy(synth.normal_embeds((n_global, 1, 32), 5))
        sr = ShardedRollout(None, None, rank=rank, world_size=world)
        sl = sr.local_slice(n_global)
        assert (sl.start, sl.stop) == shard_bounds(n_global, world, rank)
        # the shard is rolled out independently (the CPU oracle stands in for the CUDA path on this box)
        local, _ = O.gpt2_generate(sd, cfg, x0[sl], 12, use_cache=True)
        full = sr.gather(local, n_global)
        ref, _ = O.gpt2_generate(sd, cfg, x0, 12, use_cache=True)
        ok_roll = bool(torch.allclose(full, ref, atol=1e-6))
        # data-parallel training: averaging equal-size shard gradients == global-batch gradient (MSE terms),
        # and the batch-independent regulariser is averaged, not summed
        cfg_l = synth.make_config("lorenz")
        sd_e = O.to_torch(synth.lorenz_embedding_weights(cfg_l, 9))
        states = torch.from_numpy(synth.lorenz_trajectories(8, 3, seed=2))
        lo, hi = shard_bounds(8, world, rank)
        _, _, g_local = O.lorenz_train_grads(sd_e, cfg_l, states[lo:hi])
        g_avg = average_flat_grad(g_local.clone(), world)
        _, _, g_ref = O.lorenz_train_grads(sd_e, cfg_l, states)
        ok_grad = O.rel_l2(g_avg, g_ref) < 1e-5
        q.put((rank, ok_roll, ok_grad, tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_rollout_and_grad_average_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_global = 7                     # ragged: shards of 4 and 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_roll, ok_grad, shape in res:
        assert ok_roll and ok_grad, (rank, ok_roll, ok_grad)
        assert shape == (n_global, 12, 32)
