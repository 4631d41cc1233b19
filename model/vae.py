"""``model.vae.ConvVAE`` -- the reference's model plugin surface (``model/vae.py:8-145`` of
JeremyCCHsu/vae-npvc) backed by the B200-native engine.

Same module path, class name, constructor and methods as the reference, so
``getattr(import_module('model.vae'), 'ConvVAE')`` (``main.py:39-40``, ``convert.py:29-30``) finds
it: ``ConvVAE(arch, is_training=False)``, ``.loss(x, y) -> {'G','D_KL','logP'}``,
``.encode(x) -> z_mu``, ``.decode(z, y) -> xh (NHWC)``, ``.generate = .decode``.
Tensors are eager torch CUDA tensors instead of TF graph nodes; all arithmetic is the CUDA library.
"""
import torch

from vae_npvc_b200.engine import Engine


class LossDict(dict):
    """``loss`` result: the reference's dict keys plus what an eager trainer needs to re-run the
    step (the TF graph re-evaluated ``loss`` on a fresh queue batch at every ``sess.run``)."""
    machine = None
    feed = None          # (x, y): tensors (fixed batch) or queue handles (fresh batch per step)


class ConvVAE(object):
    def __init__(self, arch, is_training=False, device=None, seed=0, max_chunk=0):
        '''
        Conditional conv-VAE over 513-bin frames (model/vae.py:9-34).
        `arch`: network architecture (`dict`, the architecture-*.json content verbatim)
        `is_training`: unused (kept, as in the reference, for historical reasons)
        '''
        self.arch = arch
        self._sanity_check()
        self.is_training = is_training
        self.engine = Engine(arch, device=device, max_chunk=max_chunk)
        self.device = self.engine.device
        # variables: one flat fp32 buffer, TF variable order / layouts (npvc_param_table)
        self.theta = self.engine.init_theta(seed)
        self.y_emb = self.variables()['y_embedding/y_emb']
        # the tf.random_normal of GaussianSampleLayer (util/layers.py:154) is drawn inside the sampler kernel; its
        # seed and the pass / step counters live in 32 bytes of device memory (npvc_step_state)
        self.state = self.engine.new_step_state(seed + 1)
        self.frame_offset = 0        # data-parallel ranks: rank * frames-per-rank (distinct noise from one seed)
        self.generate = self.decode  # for VAE-GAN extension (model/vae.py:34)

    def _sanity_check(self):
        for net in ['encoder', 'generator']:
            assert len(self.arch[net]['output']) == len(self.arch[net]['kernel']) == len(self.arch[net]['stride'])

    # -- helpers ---------------------------------------------------------------------------
    def variables(self):
        """{TF variable name: tensor view in TF shape} over the flat parameter buffer."""
        return self.engine.named_views(self.theta)

    def _frames(self, x):
        """[N,1,513,1] NCHW (analyzer.py:116-122) or [N,513] -> contiguous float32 CUDA [N,513]."""
        if not torch.is_tensor(x):
            x = torch.as_tensor(x)
        x = x.to(self.device, torch.float32, non_blocking=True)
        return x.reshape(x.shape[0], -1).contiguous()

    def _labels(self, y):
        if not torch.is_tensor(y):
            y = torch.as_tensor(y)
        return y.to(self.device, torch.int64, non_blocking=True).reshape(-1).contiguous()

    # -- reference API ---------------------------------------------------------------------
    def loss(self, x, y, eps=None):
        """model/vae.py:106-137.  Returns {'G': -logPx + D_KL, 'D_KL', 'logP'} (0-dim tensors)."""
        feed = (x, y)
        if hasattr(x, 'dequeue'):                      # analyzer.read() queue handles
            x, y = x.dequeue(peek=True)
        xf, yl = self._frames(x), self._labels(y)
        out = self.engine.loss_fwd_bwd(self.theta, xf, yl, eps, grad=None, outputs=False, state=self.state,
                                       frame_offset=self.frame_offset)
        loss = LossDict(G=out['losses'][0], D_KL=out['losses'][1], logP=out['losses'][2])
        loss.machine, loss.feed = self, feed
        return loss

    def loss_and_grad(self, x, y, grad, eps=None, outputs=False, losses=None):
        """Forward + backward into the flat `grad` buffer (what optimizer.minimize differentiates,
        trainer/vae.py:24).  eps=None: the sampler draws in-kernel and the device step state advances.
        Returns the engine's output dict (losses = [G, D_KL, logP])."""
        xf, yl = self._frames(x), self._labels(y)
        return self.engine.loss_fwd_bwd(self.theta, xf, yl, eps, grad=grad, outputs=outputs, losses=losses,
                                        state=self.state, frame_offset=self.frame_offset)

    def encode(self, x):
        """model/vae.py:139-141: z_mu only."""
        mu, _ = self.engine.encode(self.theta, self._frames(x))
        return mu

    def decode(self, z, y):
        """model/vae.py:143-145: generator output in NHWC [N,513,1,1]."""
        xh = self.engine.decode(self.theta, z.to(self.device, torch.float32).contiguous(), self._labels(y))
        return xh.view(xh.shape[0], xh.shape[1], 1, 1)
