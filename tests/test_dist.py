"""world_size-2 gloo tests (CPU) of the data-parallel logic: shard -> per-rank gradient (oracle as
the gradient source) -> one flat all-reduce -> 1/world scale -> replicated TF-form Adam must equal
the single-process result on the whole batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import convvae_ref as R
from vae_npvc_b200 import vcc2016_vae_arch
from vae_npvc_b200.parallel import allreduce_flat_grad_, broadcast_params_, reduce_loss_scalars, shard_bounds


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    arch = vcc2016_vae_arch()
    P = R.init_params(arch, 0 if rank == 0 else 7)             # rank 1 starts different: broadcast must fix it
    theta = torch.tensor(R.flatten_params(arch, P, np.float64))
    broadcast_params_(theta, src=0)
    P = R.unflatten_params(arch, theta.numpy())
    x, y, eps = R.make_inputs(arch, n)
    lo, hi = shard_bounds(n, rank, world)
    o = R.forward(arch, P, x[lo:hi], y[lo:hi], eps[lo:hi], with_grads=True)
    g = torch.tensor(R.flatten_params(arch, o["grads"], np.float64))
    scale = allreduce_flat_grad_(g)
    losses = reduce_loss_scalars(torch.tensor([float(o["G"]), float(o["D_KL"]), float(o["logP"])], dtype=torch.float64))
    th1, _, _ = R.adam_step(theta.numpy(), g.numpy() * scale, 0.0, 0.0, 1, 1e-4, 0.5, 0.999)
    out[rank] = (th1, (g * scale).numpy(), losses.numpy())
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 16, 16384):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_dp_equals_single_process():
    world, n = 2, 8
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    arch = vcc2016_vae_arch()
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, n)
    ref = R.forward(arch, P, x, y, eps, with_grads=True)
    g = R.flatten_params(arch, ref["grads"], np.float64)
    th_ref, _, _ = R.adam_step(R.flatten_params(arch, P, np.float64), g, 0.0, 0.0, 1, 1e-4, 0.5, 0.999)
    for r in range(world):
        th1, gr, losses = out[r]
        assert np.abs(gr - g).max() <= 1e-12 * np.abs(g).max()          # mean of per-rank means == global mean
        assert np.abs(th1 - th_ref).max() <= 1e-9
        assert abs(losses[0] - ref["G"]) <= 1e-9 * abs(ref["G"])
    assert np.array_equal(out[0][0], out[1][0])                          # replicas stay bit-identical
