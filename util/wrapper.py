"""Log-dir helper with the reference's naming (``util/wrapper.py:99-132``): default
``logdir/train/<MMDD-HHMM-SS-YYYY>``; the three flags keep their meaning (and the reference's
unbound-variable bugs on the non-default paths are fixed)."""
import os
from datetime import datetime


def validate_log_dirs(args):
    ''' Create a default log dir (if necessary) '''
    def get_default_logdir(logdir_root):
        started = datetime.now().strftime('%m%d-%H%M-%S-%Y')
        logdir = os.path.join(logdir_root, 'train', started)
        print('Using default logdir: {}'.format(logdir))
        return logdir

    if args.logdir and args.restore_from:
        raise ValueError('You can only specify one of the following: --logdir and --restore_from')
    if args.logdir and args.logdir_root:
        raise ValueError('You can only specify either --logdir or --logdir_root')
    logdir_root = args.logdir_root if args.logdir_root else 'logdir'
    logdir = args.logdir if args.logdir else get_default_logdir(logdir_root)
    restore_from = args.restore_from if args.restore_from else logdir
    return {'logdir': logdir, 'logdir_root': logdir_root, 'restore_from': restore_from}
