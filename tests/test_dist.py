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


class _Machine(object):
    """What VAETrainer touches of a machine (theta, variables(), the device step state), on CPU tensors."""

    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        self.theta = torch.randn(10, generator=g)
        self.state = torch.zeros(4, dtype=torch.int64)

    def variables(self):
        return {"Encoder/dense/kernel": self.theta[:6].view(2, 3), "Encoder/dense/bias": self.theta[6:]}


def _restore_worker(rank, world, port, root, out):
    """Rank 0 owns the only logdir that holds checkpoints; rank 1 is given a directory that does not exist (no shared
    filesystem, per-rank timestamped logdirs).  After restore() both replicas must hold rank 0's variables, Adam
    slots and step."""
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tv = importlib.import_module("trainer.vae")
    arch = {"training": {"lr": 1e-4, "beta1": 0.5, "beta2": 0.999, "max_iter": 1}}
    logdir = os.path.join(root, "train") if rank == 0 else os.path.join(root, "rank1-does-not-exist")
    m = _Machine(10 + rank)
    t = tv.VAETrainer({"G": 0.0}, arch, None, {"logdir": None, "restore_from": logdir}); t.machine = m
    step = t.restore()
    st = t._state
    out[rank] = (step, t.global_step, m.theta.clone().numpy(), st["m"].clone().numpy(), st["v"].clone().numpy(), m.state.tolist())
    # a second logdir with nothing in it: every rank must agree that there is nothing to restore
    t2 = tv.VAETrainer({"G": 0.0}, arch, None, {"logdir": None, "restore_from": os.path.join(root, "empty-%d" % rank)}); t2.machine = _Machine(3)
    out["none%d" % rank] = t2.restore()
    dist.destroy_process_group()


def test_restore_reads_on_rank0_and_broadcasts(tmp_path):
    import importlib
    tv = importlib.import_module("trainer.vae")
    arch = {"training": {"lr": 1e-4, "beta1": 0.5, "beta2": 0.999, "max_iter": 1}}
    src = _Machine(1)
    t0 = tv.VAETrainer({"G": 0.0}, arch, None, {"logdir": str(tmp_path / "train")}); t0.machine = src
    st = t0._ensure_state(src); st["m"].fill_(0.25); st["v"].fill_(0.5)
    t0.global_step = 77; path = t0.save()
    assert path.endswith("model.ckpt-77") and not [f for f in os.listdir(tmp_path / "train") if ".tmp." in f]     # atomic rename left no temp file
    t_fresh = tv.VAETrainer({"G": 0.0}, arch, None, {"logdir": str(tmp_path / "fresh")}); t_fresh.machine = _Machine(5)
    assert t_fresh.save().endswith("model.ckpt-0")                  # a trainer that never stepped can still save (zero Adam slots)
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_restore_worker, args=(2, _free_port(), str(tmp_path), out), nprocs=2, join=True)
    for r in range(2):
        step, gs, theta, m, v, state = out[r]
        assert step == 77 and gs == 77 and state[1:3] == [77, 77]
        assert np.array_equal(theta, src.theta.numpy()) and np.all(m == 0.25) and np.all(v == 0.5)
        assert out["none%d" % r] is None
