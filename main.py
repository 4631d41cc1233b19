"""Training driver with the reference's CLI and call sequence (``main.py:13-78``):

    python main.py --model ConvVAE --trainer VAETrainer --architecture architecture-vae-vcc2016.json

flags -> plugin lookup (``--model_module`` / ``--trainer_module``) -> logdir -> architecture copied
into the logdir -> Tanhize from ./etc/x{max,min}.npf -> ``read`` -> ``MODEL(arch)`` -> ``loss`` ->
``TRAINER(loss, arch, args, dirs).train(...)``; argparse replaces tf.app.flags, torch replaces tf.
Launch under ``torchrun`` for data-parallel training (one process per GPU, NCCL).
"""
import argparse
import json
import os
from importlib import import_module

import numpy as np
import torch
import torch.distributed as dist

from analyzer import Tanhize, read
from util.wrapper import validate_log_dirs

parser = argparse.ArgumentParser()
parser.add_argument('--logdir_root', default=None, help='root of log dir')
parser.add_argument('--logdir', default=None, help='log dir')
parser.add_argument('--restore_from', default=None, help='restore from dir (not from *.ckpt)')
parser.add_argument('--gpu_cfg', default=None, help='GPU configuration')
parser.add_argument('--summary_freq', type=int, default=1000, help='Update summary')
parser.add_argument('--ckpt', default=None, help='specify the ckpt in restore_from (if there are multiple ckpts)')
parser.add_argument('--architecture', default='architecture-vawgan-vcc2016.json', help='network architecture')
parser.add_argument('--model_module', default='model.vae', help='Model module')
parser.add_argument('--model', default=None, help='Model: ConvVAE, VAWGAN')
parser.add_argument('--trainer_module', default='trainer.vae', help='Trainer module')
parser.add_argument('--trainer', default=None, help='Trainer: VAETrainer, VAWGANTrainer')


def main(argv=None):
    ''' NOTE: The input is rescaled to [-1, 1] '''
    args = parser.parse_args(argv)
    if args.model is None or args.trainer is None:
        raise ValueError(
            '\n  Both `model` and `trainer` should be assigned.' +
            '\n  Use `python main.py --help` to see applicable options.')
    MODEL = getattr(import_module(args.model_module, package=None), args.model)
    TRAINER = getattr(import_module(args.trainer_module, package=None), args.trainer)

    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')

    # the (timestamped) logdir is chosen once, on rank 0, and shared: every rank then names the same directory for
    # training.log / checkpoints / the architecture copy, and a resume finds one history (only rank 0 writes)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    dirs = validate_log_dirs(args) if (not distributed or dist.get_rank() == 0) else None
    if distributed:
        box = [dirs]
        dist.broadcast_object_list(box, src=0, device=torch.device('cuda', torch.cuda.current_device()))
        dirs = box[0]
    os.makedirs(dirs['logdir'], exist_ok=True)
    with open(args.architecture) as f:
        arch = json.load(f)
    with open(os.path.join(dirs['logdir'], os.path.basename(args.architecture)), 'w') as f:
        json.dump(arch, f, indent=4)

    machine = MODEL(arch)
    normalizer = Tanhize(
        xmax=np.fromfile('./etc/xmax.npf'),          # float64, as the reference reads them (main.py:58-59)
        xmin=np.fromfile('./etc/xmin.npf'),
        engine=machine.engine)
    image, label = read(
        file_pattern=arch['training']['datadir'],
        batch_size=arch['training']['batch_size'],
        capacity=2048,
        min_after_dequeue=1024,
        normalizer=normalizer,
        engine=machine.engine)

    loss = machine.loss(image, label)
    trainer = TRAINER(loss, arch, args, dirs)
    trainer.train(nIter=arch['training']['max_iter'], machine=machine)


if __name__ == '__main__':
    main()
