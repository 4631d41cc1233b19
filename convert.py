"""Conversion driver with the reference's CLI and call sequence (``convert.py:14-116``):

    python convert.py --src SF1 --trg TM3 --model ConvVAE --checkpoint logdir/train/<ts>/model.ckpt-<id>

whole-utterance features -> Tanhize -> [T,1,513,1] -> ``encode`` (z_mu) -> ``decode(z, y_trg)`` ->
un-Tanhize; log-Gaussian f0 transform.  WORLD synthesis (``pw2wav``) needs pyworld, which this image
does not have: converted features are written as ``<src>-<trg>-<name>.bin`` records (same layout as
the input) and wav synthesis is attempted only when pyworld + soundfile import.
"""
import argparse
import glob
import json
import os
from datetime import datetime
from importlib import import_module

import numpy as np
import torch

from analyzer import SPEAKERS, Tanhize, read_whole_features

parser = argparse.ArgumentParser()
parser.add_argument('--checkpoint', default=None, help='root of log dir')
parser.add_argument('--src', default='SF1', help='source speaker [SF1 - SM2]')
parser.add_argument('--trg', default='TM3', help='target speaker [SF1 - TM3]')
parser.add_argument('--output_dir', default='./logdir', help='root of output dir')
parser.add_argument('--module', default='model.vae', help='Module')
parser.add_argument('--model', default=None, help='Model')
parser.add_argument('--file_pattern', default='./dataset/vcc2016/bin/Testing Set/{}/*.bin', help='file pattern')

FS = 16000


def make_output_name(output_dir, filename, src, trg, ext):
    basename = os.path.splitext(os.path.split(str(filename, 'utf8'))[-1])[0]
    print('Processing {}'.format(basename))
    return os.path.join(output_dir, '{}-{}-{}.{}'.format(src, trg, basename, ext))


def get_default_output(logdir_root):
    started = datetime.now().strftime('%m%d-%H%M-%S-%Y')
    logdir = os.path.join(logdir_root, 'output', started)
    print('Using default logdir: {}'.format(logdir))
    return logdir


def convert_f0(f0, src, trg, etc='./etc'):
    """Log-Gaussian normalised transform (convert.py:51-57)."""
    mu_s, std_s = np.fromfile(os.path.join(etc, '{}.npf'.format(src)), np.float32)
    mu_t, std_t = np.fromfile(os.path.join(etc, '{}.npf'.format(trg)), np.float32)
    lf0 = np.where(f0 > 1., np.log(np.maximum(f0, 1e-30)), f0)
    lf0 = np.where(lf0 > 1., (lf0 - mu_s) / std_s * std_t + mu_t, lf0)
    lf0 = np.where(lf0 > 1., np.exp(lf0), lf0)
    return lf0.astype(np.float32)


def load_checkpoint(machine, path):
    ck = torch.load(path, map_location='cpu')
    views = machine.variables()
    for name, t in ck['variables'].items():
        views[name].copy_(t.to(views[name].device))
    return ck.get('global_step', 0)


def main(argv=None):
    args = parser.parse_args(argv)
    if args.model is None:
        raise ValueError('\n  You MUST specify `model`.' + '\n    Use `python convert.py --help` to see applicable options.')
    MODEL = getattr(import_module(args.module, package=None), args.model)

    logdir, ckpt = os.path.split(args.checkpoint)
    arch_file = glob.glob(os.path.join(logdir, 'architecture*.json'))[0]     # should only be 1 file
    with open(arch_file) as fp:
        arch = json.load(fp)
    machine = MODEL(arch)
    load_checkpoint(machine, args.checkpoint)
    normalizer = Tanhize(xmax=np.fromfile('./etc/xmax.npf'), xmin=np.fromfile('./etc/xmin.npf'), engine=machine.engine)
    output_dir = get_default_output(args.output_dir)
    os.makedirs(output_dir, exist_ok=True)
    y_t_id = SPEAKERS.index(args.trg)
    for feat in read_whole_features(args.file_pattern.format(args.src)):
        x = normalizer.forward_process(feat['sp'])                   # [T,513] on the GPU
        x = x.view(-1, 1, x.shape[-1], 1)                            # nh_to_nchw (convert.py:60-63)
        y_t = torch.full((x.shape[0],), y_t_id, dtype=torch.int64, device=x.device)
        z = machine.encode(x)
        x_t = machine.decode(z, y_t)                                 # NOTE: the API yields NHWC format
        sp = normalizer.backward_process(x_t.reshape(x.shape[0], -1)).cpu().numpy()
        f0 = convert_f0(feat['f0'], args.src, args.trg)
        rec = np.concatenate([sp, feat['ap'], f0[:, None], feat['en'][:, None],
                              np.full((sp.shape[0], 1), y_t_id, np.float32)], 1).astype(np.float32)
        rec.tofile(make_output_name(output_dir, feat['filename'], args.src, args.trg, 'bin'))
        try:
            import soundfile as sf
            from analyzer import pw2wav
            feat = dict(feat, sp=sp, f0=f0)
            sf.write(make_output_name(output_dir, feat['filename'], args.src, args.trg, 'wav'), pw2wav(feat), FS)
        except ImportError:
            pass                                                     # WORLD vocoder is host-side, absent here


if __name__ == '__main__':
    main()
