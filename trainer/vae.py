"""``trainer.vae.VAETrainer`` -- the reference's trainer plugin surface (``trainer/vae.py:9-111``,
ctor contract of ``trainer/gan.py:14-23``) over the B200-native engine.

``VAETrainer(loss, arch, args, dirs)`` / ``.train(nIter, machine=None, summary_op=None)`` keep the
reference signatures.  ``_optimize`` is TF-form Adam over ALL variables (one flat buffer);
``train`` is the ``for step in range(max_iter): sess.run(opt['g'])`` hot loop with the same 60 s
status line and 300 s checkpoint cadence.  Under ``torch.distributed`` (world_size > 1) the flat
gradient is all-reduced once per step over NCCL before the replicated Adam step.
"""
import logging
import os
import re
import time

import torch
import torch.distributed as dist

from vae_npvc_b200.parallel import allreduce_flat_grad_, broadcast_params_, world_info


def latest_checkpoint(logdir):
    """Newest ``model.ckpt-<step>`` in `logdir` (what ``tf.train.get_checkpoint_state`` resolves in
    ``util/wrapper.py:39-60``), or None."""
    best, best_step = None, -1
    if logdir and os.path.isdir(logdir):
        for f in os.listdir(logdir):
            m = re.fullmatch(r'model\.ckpt-(\d+)', f)
            if m and int(m.group(1)) > best_step:
                best, best_step = os.path.join(logdir, f), int(m.group(1))
    return best


class VAETrainer(object):
    def __init__(self, loss, arch, args, dirs):
        self.loss = loss
        self.arch = arch
        self.args = args
        self.dirs = dirs
        self.machine = getattr(loss, 'machine', None)
        self.opt = self._optimize()
        # training.log in the logdir (trainer/gan.py:20-23); an explicit handler, because
        # logging.basicConfig is a no-op once the host application has configured logging
        self.logger = logging.getLogger('vae_npvc_b200.training.%x' % id(self))
        self.logger.setLevel(logging.INFO)
        if dirs and dirs.get('logdir'):
            os.makedirs(dirs['logdir'], exist_ok=True)
            self.logger.addHandler(logging.FileHandler(os.path.join(dirs['logdir'], 'training.log')))

    # -- optimiser (trainer/vae.py:10-28) ---------------------------------------------------
    def _optimize(self):
        tr = self.arch['training']
        self.lr, self.b1, self.b2 = tr['lr'], tr['beta1'], tr['beta2']
        self.global_step = 0
        self._state = None
        self.rank, self.world = 0, 1
        self.use_graph = os.environ.get('NPVC_GRAPH', '1') != "0"
        return {'g': self._train_step, 'global_step': lambda: self.global_step}

    def _ensure_state(self, machine):
        if self._state is None or self._state['machine'] is not machine:
            th = machine.theta
            self._state = dict(machine=machine, grad=torch.empty_like(th), m=torch.zeros_like(th), v=torch.zeros_like(th),
                               losses=torch.zeros(3, device=th.device), graphs={})
            self.rank, self.world = world_info()
            broadcast_params_(th, src=0)       # identical replicas: rank 0's initial variables
            self._sync_step_state(machine)
        return self._state

    def _sync_step_state(self, machine):
        """The device-resident counters (npvc_step_state) follow `global_step` (after a restore)."""
        state = getattr(machine, 'state', None)
        if state is not None:
            state[1] = self.global_step              # draws: the sampler's pass counter
            state[2] = self.global_step              # step: the pass that follows makes it the t of its Adam update

    def _eager_step(self, machine, st, x, y, eps=None):
        out = machine.loss_and_grad(x, y, st['grad'], eps=eps, losses=st['losses'])
        scale = allreduce_flat_grad_(st['grad'], self.world)      # one 3.76 MB bucket over NVLink
        if eps is not None:                                       # caller-supplied draw: the pass did not advance the device counters
            machine.state[2] += 1
        machine.engine.adam_step(machine.theta, st['grad'], st['m'], st['v'], machine.state,
                                 self.lr, self.b1, self.b2, 1e-8, scale)
        return out['losses']

    def _train_step(self, x=None, y=None, eps=None):
        """One ``sess.run(opt['g'])``: fwd + bwd (+ all-reduce) + Adam on one batch of frames.

        Nothing the host computes changes from step to step (the Adam step count and the sampler's counters live on
        the device), so after two eager steps on a batch shape the whole step -- ~60 kernel launches, the side-stream
        fork / join of the weight gradients, the all-reduce -- is captured ONCE in a CUDA graph and replayed: the
        reference's regime of 16 frames per step (architecture-vae-vcc2016.json:23) is launch-bound otherwise.
        ``NPVC_GRAPH=0`` keeps every step eager."""
        machine = self.machine
        st = self._ensure_state(machine)
        if x is None:
            x, y = self.loss.feed
            if hasattr(x, 'dequeue'):
                x, y = x.dequeue()
        self.global_step += 1
        if eps is not None or not self.use_graph:
            return self._eager_step(machine, st, x, y, eps)
        xf, yl = machine._frames(x), machine._labels(y)
        machine.frame_offset = self.rank * xf.shape[0]
        key = (xf.shape[0],)
        g = st['graphs'].get(key)
        if g is None:
            g = st['graphs'][key] = dict(seen=0, graph=None, x=torch.empty_like(xf), y=torch.empty_like(yl))
        if g['graph'] is None:
            g['seen'] += 1
            if g['seen'] <= 2:                                   # warm-up: workspace, tensor maps, kernel attributes, NCCL
                return self._eager_step(machine, st, xf, yl)
            g['x'].copy_(xf); g['y'].copy_(yl)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                self._eager_step(machine, st, g['x'], g['y'])
            g['graph'] = graph                                   # (the capture did not run the step)
        g['x'].copy_(xf, non_blocking=True); g['y'].copy_(yl, non_blocking=True)
        g['graph'].replay()
        machine.engine._packed_for = None                        # theta changed behind the operand packs
        return st['losses']

    # -- status line (trainer/vae.py:31-52) -------------------------------------------------
    def _refresh_status(self, sess=None):
        losses = self._state['losses'].tolist() if self._state else [float('nan')] * 3
        msg = 'Iter {:05d}: '.format(self.global_step)
        msg += 'log P(x|z, y) = {:.3e} '.format(losses[2])
        msg += 'D_KL(z) = {:.3e} '.format(losses[1])
        print('\r{}'.format(msg), end='', flush=True)
        self.logger.info(msg)
        for hd in self.logger.handlers:
            hd.flush()
        return msg

    # -- TensorBoard scalars (model/vae.py:132-133: 'KL-div', 'logPx'; Supervisor's summary thread) -------
    def _write_summaries(self, losses=None):
        """Event file in the logdir with the reference's scalar tags, at the trainer's global step.
        Best effort: without the tensorboard package this is a no-op."""
        if not (self.dirs and self.dirs.get('logdir')):
            return False
        if getattr(self, '_tb', None) is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
            except Exception:
                self._tb = False
                return False
            self._tb = SummaryWriter(self.dirs['logdir'])
        if self._tb is False:
            return False
        if losses is None:
            losses = self._state['losses'].tolist() if self._state else [float('nan')] * 3
        self._tb.add_scalar('KL-div', losses[1], self.global_step)
        self._tb.add_scalar('logPx', losses[2], self.global_step)
        self._tb.flush()
        return True

    def save(self, path=None):
        """Checkpoint: the variables under their TF names + Adam slots + global_step (the
        reference's Supervisor autosave, trainer/vae.py:78-84, as one torch file)."""
        st = self._ensure_state(self.machine)        # (a trainer that has not stepped yet saves zero Adam slots)
        path = path or os.path.join(self.dirs['logdir'], 'model.ckpt-{}'.format(self.global_step))
        tmp = '{}.tmp.{}'.format(path, os.getpid())
        torch.save({'variables': {k: v.cpu() for k, v in self.machine.variables().items()},
                    'adam_m': st['m'].cpu(), 'adam_v': st['v'].cpu(), 'global_step': self.global_step}, tmp)
        os.replace(tmp, path)                        # atomic: a crash mid-write never leaves a truncated newest checkpoint
        return path

    def restore(self, logdir=None, ckpt=None, machine=None):
        """Resume from a checkpoint written by ``save`` (the Supervisor's restore-on-start of
        ``trainer/vae.py:78-84``; ``load`` of ``util/wrapper.py:32-62``): `ckpt` names a file inside `logdir`,
        otherwise the newest ``model.ckpt-<step>`` there is taken.  Restores the variables, the Adam slots
        and ``global_step``; returns the step, or None when there is nothing to restore."""
        machine = machine if machine is not None else self.machine
        rank, world = world_info()
        if world > 1:
            return self._restore_distributed(logdir, ckpt, machine, rank)
        return self._restore_local(logdir, ckpt, machine)

    def _restore_distributed(self, logdir, ckpt, machine, rank):
        """Data-parallel resume: rank 0 alone reads the file (ranks need no shared filesystem and cannot disagree on
        which checkpoint is the newest); the variables, both Adam slots and global_step are then broadcast."""
        self.machine = machine
        st = self._ensure_state(machine)             # every rank: its first call broadcasts rank 0's variables (a collective)
        step = self._restore_local(logdir, ckpt, machine) if rank == 0 else None
        dev = machine.theta.device
        flag = torch.tensor([-1 if step is None else int(step)], dtype=torch.int64, device=dev)
        dist.broadcast(flag, src=0)
        if int(flag.item()) < 0:
            return None
        for t in (machine.theta, st['m'], st['v']):
            dist.broadcast(t, src=0)
        self.global_step = int(flag.item())
        self._sync_step_state(machine)
        if hasattr(machine, 'engine') and hasattr(machine.engine, '_packed_for'):
            machine.engine._packed_for = None
        return self.global_step

    def _restore_local(self, logdir, ckpt, machine):
        logdir = logdir or (self.dirs or {}).get('restore_from') or (self.dirs or {}).get('logdir')
        path = os.path.join(logdir, ckpt) if (ckpt and logdir) else latest_checkpoint(logdir)
        if not path or not os.path.exists(path):
            return None
        ck = torch.load(path, map_location='cpu')
        views = machine.variables()
        missing = sorted(set(views) - set(ck['variables']))
        if missing:
            raise ValueError('checkpoint {} lacks variables: {}'.format(path, ', '.join(missing[:4])))
        for name, v in views.items():
            t = ck['variables'][name]
            if tuple(t.shape) != tuple(v.shape):
                raise ValueError('checkpoint {}: {} has shape {}, the architecture wants {}'.format(
                    path, name, tuple(t.shape), tuple(v.shape)))
            v.copy_(t.to(v.device))
        if hasattr(machine, 'engine') and hasattr(machine.engine, '_packed_for'):
            machine.engine._packed_for = None          # the variables changed behind the operand packs
        self.machine = machine
        st = self._ensure_state(machine)
        if 'adam_m' in ck and 'adam_v' in ck:
            st['m'].copy_(ck['adam_m'].to(st['m'].device)); st['v'].copy_(ck['adam_v'].to(st['v'].device))
        self.global_step = int(ck.get('global_step', 0))
        self._sync_step_state(machine)
        self.logger.info('Restored {} (global step {})'.format(path, self.global_step))
        return self.global_step

    # -- hot loop (trainer/vae.py:73-99) ----------------------------------------------------
    def train(self, nIter=None, machine=None, summary_op=None, status_secs=60, save_secs=300, summary_secs=120):
        if machine is not None:
            self.machine = machine
        if self.machine is None:
            raise ValueError('VAETrainer needs the machine: pass `machine=` or a loss from machine.loss()')
        rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
        if self.global_step == 0 and self.dirs and (self.dirs.get('restore_from') or self.dirs.get('logdir')):
            # tf.train.Supervisor restores the newest checkpoint of its logdir before the first step; under
            # torch.distributed rank 0 reads it and broadcasts variables, Adam slots and the step
            self.restore(ckpt=getattr(self.args, 'ckpt', None))
        t_status = t_save = t_summary = time.time()
        for step in range(self.arch['training']['max_iter'] if nIter is None else min(nIter, self.arch['training']['max_iter'])):
            self.opt['g']()
            now = time.time()
            if rank0 and now - t_status >= status_secs:
                self._refresh_status(); t_status = now
            if rank0 and now - t_summary >= summary_secs:
                self._write_summaries(); t_summary = now
            if rank0 and self.dirs and self.dirs.get('logdir') and now - t_save >= save_secs:
                self.save(); t_save = now
        torch.cuda.synchronize()
        if rank0 and self.dirs and self.dirs.get('logdir'):
            self._refresh_status()
            self._write_summaries()
            self.save()
