"""Host-side engine: PyTorch owns device memory and streams, libnpvc_b200.so does the work.

PyTorch is plumbing here (allocation, streams, torch.distributed); every arithmetic op of the
path runs in the hand-written CUDA library through the C-ABI.  There is no CPU fallback: without a
CUDA device (or without the built extension) the constructors raise.
"""
import functools
import math
import os

import torch

from . import lib as _lib


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream     # (of the current device: see _on_device)


def _on_device(fn):
    """Run an Engine method with the engine's device current: the library binds its device tables, kernel
    attributes and launches to the current device, and `_stream()` is that device's current stream."""
    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)
    return t


class Engine:
    """One ConvVAE launch plan + its caller-owned workspace."""

    def __init__(self, arch, device=None, max_chunk=0):
        if not torch.cuda.is_available():
            raise RuntimeError("vae_npvc_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("vae_npvc_b200 runs on CUDA devices only (got %s)" % self.device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.arch = arch
        self.handle = _lib.Handle(arch, max_chunk)
        self.lib = self.handle.lib
        self.table = self.handle.param_table()
        self.n_params = self.handle.param_count()
        self.z_dim = arch["z_dim"]
        self.in_h = arch["hwc"][0]
        self._ws = None
        self._ws_train = False

    # ------------------------------------------------------------------ parameters
    @_on_device
    def init_theta(self, seed=0, perturb=0.0):
        """Flat fp32 parameter vector with the TF default initialisers of the path:
        glorot_uniform kernels, zero biases / LN offsets, unit LN scales.  perturb > 0 moves
        biases / LN params off 0 / 1 (benchmarks: exercise those code paths)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        theta = torch.empty(self.n_params, dtype=torch.float32)
        for p in self.table:
            sl = theta[p["offset"]:p["offset"] + p["size"]]
            if p["init"] == 0:
                lim = math.sqrt(6.0 / (p["fan_in"] + p["fan_out"]))
                sl.copy_((torch.rand(p["size"], generator=g) * 2 - 1) * lim)
            else:
                base = 1.0 if p["init"] == 2 else 0.0
                sl.fill_(base)
                if perturb:
                    sl.add_((torch.rand(p["size"], generator=g) * 2 - 1) * perturb)
        return theta.to(self.device)

    def named_views(self, flat):
        """{tf_variable_name: view of `flat` in its TF shape}."""
        return {p["name"]: flat[p["offset"]:p["offset"] + p["size"]].view(p["shape"]) for p in self.table}

    # ------------------------------------------------------------------ workspace
    @_on_device
    def workspace(self, n, train):
        need = self.handle.workspace_bytes(n, train)
        # the training layout is a superset of the inference layout with identical offsets for
        # the packed operands, so one buffer sized for the larger request serves both
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            torch.cuda.empty_cache()
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._packed_for = None
        return self._ws

    _packed_for = None

    @_on_device
    def pack(self, theta, ws=None):
        ws = ws if ws is not None else self.workspace(1, False)
        _lib.check(self.lib.npvc_pack_weights(self.handle.h, _ptr(_f32c(theta, "theta")), ws.data_ptr(), ws.numel(), _stream()))
        self._packed_for = (theta.data_ptr(), theta._version, ws.data_ptr())

    def _poison(self, ws):
        """NPVC_DEBUG_POISON=1 (tests / bring-up): every byte a pass does not write itself reads as a NaN pattern."""
        if os.environ.get("NPVC_DEBUG_POISON"):
            ws.fill_(0xFF); self._packed_for = None

    def _ensure_packed(self, theta, ws):
        if self._packed_for != (theta.data_ptr(), theta._version, ws.data_ptr()):
            self.pack(theta, ws)

    # ------------------------------------------------------------------ path entry points
    @_on_device
    def encode(self, theta, x):
        """x [n,513] -> (mu, lv) [n,z]   (model/vae.py:72-82)."""
        x = _f32c(x, "x").view(-1, self.in_h)
        n = x.shape[0]
        ws = self.workspace(n, False)
        self._poison(ws)
        self._ensure_packed(theta, ws)
        mu = torch.empty(n, self.z_dim, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        _lib.check(self.lib.npvc_encode(self.handle.h, _ptr(theta), _ptr(x), n, _ptr(mu), _ptr(lv), ws.data_ptr(), ws.numel(), _stream()))
        return mu, lv

    @_on_device
    def sample(self, mu, lv, eps):
        z = torch.empty_like(mu)
        _lib.check(self.lib.npvc_sample(self.handle.h, _ptr(_f32c(mu, "mu")), _ptr(_f32c(lv, "lv")), _ptr(_f32c(eps, "eps")), mu.shape[0], _ptr(z), _stream()))
        return z

    @_on_device
    def decode(self, theta, z, y):
        """z [n,z], y [n] int64 -> xh [n,513]   (model/vae.py:84-103)."""
        z = _f32c(z, "z")
        n = z.shape[0]
        if y.dtype != torch.int64 or not y.is_cuda or not y.is_contiguous() or y.numel() != n:
            raise ValueError("y must be a contiguous int64 CUDA tensor of n labels")
        ws = self.workspace(n, False)
        self._poison(ws)
        self._ensure_packed(theta, ws)
        xh = torch.empty(n, self.in_h, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.npvc_decode(self.handle.h, _ptr(theta), _ptr(z), _ptr(y), n, _ptr(xh), ws.data_ptr(), ws.numel(), _stream()))
        return xh

    @_on_device
    def new_step_state(self, seed=0):
        """Device-resident {seed, draws, step, reserved} of a training loop (include/npvc_b200.h, npvc_step_state)."""
        st = torch.zeros(4, dtype=torch.int64, device=self.device)
        st[0] = int(seed) & 0x7FFFFFFFFFFFFFFF
        return st

    @_on_device
    def normal_draw(self, state, n, frame_offset=0):
        """The N(0,1) draw the next train pass over these frames will make (tests / debugging)."""
        eps = torch.empty(n, self.z_dim, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.npvc_normal_draw(self.handle.h, state.data_ptr(), int(frame_offset), n, _ptr(eps), _stream()))
        return eps

    @_on_device
    def loss_fwd_bwd(self, theta, x, y, eps=None, grad=None, outputs=True, losses=None, state=None, frame_offset=0):
        """Forward + losses (+ backward into `grad` when given)   (model/vae.py:106-130).
        eps: the N(0,1) draw of the sampler as a tensor (parity tests), or None with `state` (new_step_state()):
        drawn in-kernel, the state's pass / step counters advance on the device.
        Returns dict(losses=[G, D_KL, logP] device tensor, z, mu, lv, xh)."""
        x = _f32c(x, "x").view(-1, self.in_h)
        n = x.shape[0]
        if y.dtype != torch.int64 or not y.is_cuda or y.numel() != n:
            raise ValueError("y must be an int64 CUDA tensor of n labels")
        if eps is None:
            if state is None or state.dtype != torch.int64 or state.numel() != 4 or not state.is_cuda:
                raise ValueError("without eps a device step state (new_step_state()) is required")
        else:
            eps = _f32c(eps, "eps")
            if eps.shape != (n, self.z_dim):
                raise ValueError("eps must be [n, z_dim]")
        ws = self.workspace(n, True)
        self._poison(ws)
        repack = self._packed_for != (theta.data_ptr(), theta._version, ws.data_ptr())
        out = {}
        if outputs:
            for k, w in (("z", self.z_dim), ("mu", self.z_dim), ("lv", self.z_dim), ("xh", self.in_h)):
                out[k] = torch.empty(n, w, dtype=torch.float32, device=self.device)
        if losses is None:
            losses = torch.empty(3, dtype=torch.float32, device=self.device)
        if grad is not None:
            _f32c(grad, "grad")
        if eps is not None:
            _lib.check(self.lib.npvc_loss_fwd_bwd(
                self.handle.h, _ptr(_f32c(theta, "theta")), _ptr(x), _ptr(y), _ptr(eps), n,
                _ptr(out.get("z")), _ptr(out.get("mu")), _ptr(out.get("lv")), _ptr(out.get("xh")),
                _ptr(losses), _ptr(grad), 1 if repack else 0, ws.data_ptr(), ws.numel(), _stream()))
        else:
            _lib.check(self.lib.npvc_train_fwd_bwd(
                self.handle.h, _ptr(_f32c(theta, "theta")), _ptr(x), _ptr(y), state.data_ptr(), int(frame_offset), n,
                _ptr(out.get("z")), _ptr(out.get("mu")), _ptr(out.get("lv")), _ptr(out.get("xh")),
                _ptr(losses), _ptr(grad), 1 if repack else 0, ws.data_ptr(), ws.numel(), _stream()))
        self._packed_for = (theta.data_ptr(), theta._version, ws.data_ptr())
        out["losses"] = losses
        return out

    @_on_device
    def adam_step(self, theta, grad, m, v, step, lr, beta1, beta2, eps=1e-8, grad_scale=1.0):
        """TF-form Adam on the flat buffers (trainer/vae.py:16-24).  `step`: the 1-based t as a Python int, or a
        device step state (new_step_state()) whose `step` counter the kernel reads."""
        if torch.is_tensor(step):
            _lib.check(self.lib.npvc_adam_step_dev(self.handle.h, _ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(),
                                                   step.data_ptr(), lr, beta1, beta2, eps, grad_scale, _stream()))
        else:
            _lib.check(self.lib.npvc_adam_step(self.handle.h, _ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(),
                                               int(step), lr, beta1, beta2, eps, grad_scale, _stream()))
        self._packed_for = None      # theta changed behind torch's version counter

    @_on_device
    def tanhize_forward(self, x, xmin, xmax, out=None):
        out = torch.empty_like(x) if out is None else out
        _lib.check(self.lib.npvc_tanhize_forward(self.handle.h, _ptr(_f32c(x, "x")), _ptr(xmin), _ptr(xmax), x.shape[0], x.shape[1], _ptr(out), _stream()))
        return out

    @_on_device
    def tanhize_backward(self, x, xmin, xmax, out=None):
        out = torch.empty_like(x) if out is None else out
        _lib.check(self.lib.npvc_tanhize_backward(self.handle.h, _ptr(_f32c(x, "x")), _ptr(xmin), _ptr(xmax), x.shape[0], x.shape[1], _ptr(out), _stream()))
        return out

    @_on_device
    def unpack_records(self, records, sp_dim, xmin=None, xmax=None):
        """[n, 1029] float32 records -> (x [n,513] Tanhize'd, y [n] int64)   (analyzer.py:111-127)."""
        records = _f32c(records, "records")
        n, rf = records.shape
        x = torch.empty(n, sp_dim, dtype=torch.float32, device=self.device)
        y = torch.empty(n, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.npvc_unpack_records(self.handle.h, _ptr(records), n, rf, sp_dim, _ptr(xmin), _ptr(xmax), _ptr(x), _ptr(y), _stream()))
        return x, y

    @_on_device
    def debug_buffer(self, name, n):
        per = self.lib.npvc_debug_buffer(self.handle.h, name.encode(), self._ws.data_ptr(), None, n, None)
        if per < 0:
            raise KeyError(name)
        plan_buf = [b for b in self.handle.plan()["bufs"] if b["name"] == name]
        cnt = (plan_buf[0]["fixed"] + plan_buf[0]["per_frame"] * n) if plan_buf else per
        out = torch.empty(cnt, dtype=torch.float32, device=self.device)
        self.lib.npvc_debug_buffer(self.handle.h, name.encode(), self._ws.data_ptr(), out.data_ptr(), n, _stream())
        return out

    def launch_count(self):
        return self.handle.launch_count()
