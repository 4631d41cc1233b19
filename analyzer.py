"""``analyzer`` -- the data side of the reference's plugin surface (``analyzer.py`` of
JeremyCCHsu/vae-npvc) for the B200 engine: the binary-frame reader feeding a pinned-memory
asynchronous loader, the Tanhize normaliser, and the whole-utterance reader of the convert path.

Record layout (``analyzer.py:19-22``, ``README.md:106-125``): one frame = 1029 float32 =
``sp[513] | ap[513] | f0 | en | speaker`` (4,116 bytes).  Training consumes only ``sp`` and
``speaker`` (``analyzer.py:116,127``).

WORLD analysis / synthesis (``wav2pw`` / ``pw2wav``, ``analyzer.py:25-72,162-187``) stays on the host
and needs pyworld + librosa, which this image does not ship: those two entry points raise.
"""
import glob
import os
import threading

import numpy as np
import torch

FFT_SIZE = 1024
SP_DIM = FFT_SIZE // 2 + 1
FEAT_DIM = SP_DIM + SP_DIM + 1 + 1 + 1   # [sp, ap, f0, en, s]
RECORD_BYTES = FEAT_DIM * 4               # all features saved in float32
EPSILON = 1e-10

_DEFAULT_SPEAKERS = ['SF1', 'SF2', 'SF3', 'SM1', 'SM3', 'TF1', 'TF2', 'TM1', 'TM2', 'TM3']


def _load_speakers(path='./etc/speakers.tsv'):
    """Label ids == line order of etc/speakers.tsv (``analyzer.py:18``)."""
    if os.path.exists(path):
        with open(path) as f:
            return [s.strip() for s in f.readlines() if s.strip()]
    return list(_DEFAULT_SPEAKERS)


SPEAKERS = _load_speakers()


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('analyzer: a CUDA device is required (no CPU fallback for the hot path)')
    return torch.device('cuda', torch.cuda.current_device())


class Tanhize(object):
    ''' Normalizing `x` to [-1, 1]  (analyzer.py:75-87); per-bin xmin/xmax, CUDA kernels '''
    def __init__(self, xmin, xmax, engine=None):
        self.xmin = np.asarray(xmin)
        self.xmax = np.asarray(xmax)
        self.xscale = self.xmax - self.xmin
        self._engine = engine
        self._dev = None

    def _bind(self):
        if self._dev is None:
            if self._engine is None:
                from vae_npvc_b200 import vcc2016_vae_arch
                from vae_npvc_b200.engine import Engine
                self._engine = Engine(vcc2016_vae_arch())
            d = self._engine.device
            self._dev = (torch.as_tensor(self.xmin, dtype=torch.float32, device=d).contiguous(),
                         torch.as_tensor(self.xmax, dtype=torch.float32, device=d).contiguous())
        return self._engine, self._dev

    def _as_frames(self, x):
        eng, _ = self._bind()
        t = torch.as_tensor(x).to(eng.device, torch.float32)
        shape = t.shape
        return t.reshape(-1, len(self.xmin)).contiguous(), shape

    def forward_process(self, x):
        eng, (lo, hi) = self._bind()
        t, shape = self._as_frames(x)
        return eng.tanhize_forward(t, lo, hi).reshape(shape)

    def backward_process(self, x):
        eng, (lo, hi) = self._bind()
        t, shape = self._as_frames(x)
        return eng.tanhize_backward(t, lo, hi).reshape(shape)


# ------------------------------------------------------------------------------------------------
# host side of the frame reader (no CUDA needed: unit-tested on CPU)
# ------------------------------------------------------------------------------------------------
class RecordFiles(object):
    """Memory-mapped ``*.bin`` files seen as one sequence of 4,116-byte records."""
    def __init__(self, files, record_floats=FEAT_DIM):
        self.files = list(files)
        self.record_floats = record_floats
        self.maps = []
        for f in self.files:
            n = os.path.getsize(f) // (record_floats * 4)
            self.maps.append(np.memmap(f, dtype=np.float32, mode='r', shape=(n, record_floats)) if n else
                             np.zeros((0, record_floats), np.float32))

    def n_records(self):
        return sum(m.shape[0] for m in self.maps)


class ShuffleReader(object):
    """``tf.train.string_input_producer`` + ``FixedLengthRecordReader`` + ``shuffle_batch``
    (``analyzer.py:111-135``, ``main.py:62-68``): files in a freshly shuffled order every epoch, records
    read sequentially into a shuffle buffer of `capacity` frames, batches drawn at random from the
    buffer, which never holds fewer than `min_after_dequeue` frames after a dequeue."""
    def __init__(self, record_files, batch_size, capacity, min_after_dequeue, seed=0):
        self.rf = record_files
        self.batch = batch_size
        self.min_after = min_after_dequeue
        self.capacity = max(capacity, min_after_dequeue + batch_size)
        self.rng = np.random.RandomState(seed)
        self.pool = np.empty((self.capacity, record_files.record_floats), np.float32)
        self.fill = 0
        self._order, self._fi, self._pos = [], 0, 0
        if record_files.n_records() == 0:
            raise ValueError('no records found')

    def _next_chunk(self, want):
        """Up to `want` sequential records from the (shuffled-per-epoch) file stream."""
        while True:
            if self._fi >= len(self._order):
                self._order = list(self.rng.permutation(len(self.rf.maps)))
                self._fi, self._pos = 0, 0
            m = self.rf.maps[self._order[self._fi]]
            if self._pos >= m.shape[0]:
                self._fi += 1; self._pos = 0
                continue
            take = min(want, m.shape[0] - self._pos)
            out = m[self._pos:self._pos + take]
            self._pos += take
            return out

    def _refill(self):
        while self.fill < self.capacity:
            c = self._next_chunk(self.capacity - self.fill)
            self.pool[self.fill:self.fill + len(c)] = c
            self.fill += len(c)

    def next_batch(self, out):
        """Write one batch of records into `out` [batch, record_floats] (e.g. a pinned buffer)."""
        self._refill()
        idx = self.rng.choice(self.fill, self.batch, replace=False)
        np.take(self.pool, idx, axis=0, out=out)
        # compact: move the tail records into the holes left by the dequeued ones
        keep = np.ones(self.fill, bool); keep[idx] = False
        tail = np.nonzero(keep[self.fill - self.batch:])[0] + (self.fill - self.batch)
        holes = idx[idx < self.fill - self.batch]
        self.pool[holes] = self.pool[tail[:len(holes)]]
        self.fill -= self.batch
        return out


class FrameLoader(object):
    """Pinned-memory asynchronous loader: a host thread fills pinned record batches
    (double-buffered), the consumer copies them to the GPU on a side stream and runs the fused
    ``slice sp | Tanhize | cast speaker`` kernel (``npvc_unpack_records``)."""
    def __init__(self, reader, normalizer=None, engine=None, depth=3, fmt='NCHW'):
        self.reader, self.fmt = reader, fmt
        self.dev = _device()
        if engine is None:
            engine = normalizer._bind()[0] if normalizer is not None else None
        if engine is None:
            from vae_npvc_b200 import vcc2016_vae_arch
            from vae_npvc_b200.engine import Engine
            engine = Engine(vcc2016_vae_arch())
        self.engine = engine
        self.norm = normalizer._bind()[1] if normalizer is not None else (None, None)
        rf = reader.rf.record_floats
        self.host = [torch.empty(reader.batch, rf, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.free = threading.Semaphore(depth)
        self.ready = []
        self.cv = threading.Condition()
        self.stream = torch.cuda.Stream(device=self.dev)
        self._peek = None
        self._pending = []                  # events of H2D copies whose pinned slots are not yet back with the producer
        self._stop = False
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _produce(self):
        k = 0
        while not self._stop:
            self.free.acquire()
            if self._stop:
                break
            buf = self.host[k % len(self.host)]
            self.reader.next_batch(buf.numpy())
            with self.cv:
                self.ready.append(k % len(self.host)); self.cv.notify()
            k += 1

    def _release_copied(self, block=False):
        """Pinned slots go back to the producer once their H2D copy has completed -- polled, so the consumer does not
        stall on the copy it has just issued; it blocks on the OLDEST copy only when the producer is out of slots."""
        while self._pending and (block or self._pending[0].query()):
            self._pending.pop(0).synchronize()
            self.free.release()
            block = False

    def _next_device_batch(self):
        self._release_copied()
        while True:
            with self.cv:
                if self.ready:
                    slot = self.ready.pop(0)
                    break
                if not self._pending:
                    self.cv.wait(0.05)
                    continue
            self._release_copied(block=True)            # every slot is in flight: wait for the oldest copy, not for the producer
        with torch.cuda.stream(self.stream):
            rec = self.host[slot].to(self.dev, non_blocking=True)
            x, y = self.engine.unpack_records(rec, SP_DIM, self.norm[0], self.norm[1])
            done = torch.cuda.Event(); done.record(self.stream)
        cur = torch.cuda.current_stream()
        cur.wait_event(done)
        # allocated on the loader's stream, consumed on the caller's: the caching allocator must not hand the
        # blocks back to the loader's stream while the training step still reads them
        x.record_stream(cur); y.record_stream(cur)
        self._pending.append(done)          # the pinned slot may be refilled once the H2D copy is done (_release_copied)
        if self.fmt == 'NCHW':
            x = x.view(-1, 1, SP_DIM, 1)
        elif self.fmt == 'NHWC':
            x = x.view(-1, SP_DIM, 1, 1)
        return x, y

    def dequeue(self, peek=False):
        if self._peek is not None:
            b = self._peek
            if not peek:
                self._peek = None
            return b
        b = self._next_device_batch()
        if peek:
            self._peek = b
        return b

    def close(self):
        self._stop = True
        self.free.release()


class QueueHandle(object):
    """What ``read`` returns in place of TF's symbolic dequeue tensors: ``machine.loss(image,
    label)`` / the trainer call ``.dequeue()`` to get the next device batch."""
    def __init__(self, loader, which):
        self.loader, self.which = loader, which

    def dequeue(self, peek=False):
        return self.loader.dequeue(peek=peek)


def read(file_pattern, batch_size, record_bytes=RECORD_BYTES, capacity=256, min_after_dequeue=128,
         num_threads=8, format='NCHW', normalizer=None, seed=0, engine=None):
    '''
    Read only `sp` and `speaker`  (analyzer.py:90-135)
    Return: (`feature`, `speaker`) queue handles; each dequeue yields
        `feature`: [b, 1, 513, 1] (NCHW) float32 CUDA, `speaker`: [b,] int64 CUDA
    Under torch.distributed the file list is sharded by rank (utterance sharding).
    '''
    from vae_npvc_b200.parallel import shard_bounds, world_info
    files = sorted(glob.glob(file_pattern))
    rank, world = world_info()
    if world > 1:
        lo, hi = shard_bounds(len(files), rank, world)
        files = files[lo:hi]
    if not files:
        raise ValueError('no files match {}'.format(file_pattern))
    rf = RecordFiles(files, record_bytes // 4)
    reader = ShuffleReader(rf, batch_size, capacity, min_after_dequeue, seed=seed + rank)
    loader = FrameLoader(reader, normalizer=normalizer, engine=engine, fmt=format)
    return QueueHandle(loader, 'feature'), QueueHandle(loader, 'speaker')


def read_whole_features(file_pattern, num_epochs=1):
    '''
    Whole-utterance reader of the convert path (analyzer.py:138-158): yields one dict per file with
    `sp`, `ap` [T,513], `f0`, `en` [T], `speaker` [T] int64, `filename`.
    '''
    files = sorted(glob.glob(file_pattern))
    print('{} files found'.format(len(files)))
    for _ in range(num_epochs):
        for f in files:
            value = np.fromfile(f, np.float32).reshape(-1, FEAT_DIM)
            yield {
                'sp': value[:, :SP_DIM],
                'ap': value[:, SP_DIM:2 * SP_DIM],
                'f0': value[:, SP_DIM * 2],
                'en': value[:, SP_DIM * 2 + 1],
                'speaker': value[:, SP_DIM * 2 + 2].astype(np.int64),
                'filename': f.encode('utf8'),
            }


def wav2pw(x, fs=16000, fft_size=FFT_SIZE):
    raise ImportError('WORLD analysis needs pyworld (not in this image); it stays on the host, out of the B200 path')


def pw2wav(features, feat_dim=513, fs=16000):
    raise ImportError('WORLD synthesis needs pyworld (not in this image); it stays on the host, out of the B200 path')
