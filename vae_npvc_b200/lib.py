"""ctypes binding of libnpvc_b200.so (include/npvc_b200.h).  No torch types cross this boundary:
device pointers travel as integers (``tensor.data_ptr()``), the stream as ``cudaStream_t``.
"""
import ctypes as C
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnpvc_b200.so")
MAX_LAYERS = 8


class NpvcArch(C.Structure):
    _fields_ = [
        ("in_h", C.c_int32), ("z_dim", C.c_int32), ("y_dim", C.c_int32), ("n_enc", C.c_int32),
        ("enc_out", C.c_int32 * MAX_LAYERS), ("enc_kernel", C.c_int32 * MAX_LAYERS),
        ("enc_stride", C.c_int32 * MAX_LAYERS),
        ("gen_h", C.c_int32), ("gen_c", C.c_int32), ("n_gen", C.c_int32),
        ("gen_out", C.c_int32 * MAX_LAYERS), ("gen_kernel", C.c_int32 * MAX_LAYERS),
        ("gen_stride", C.c_int32 * MAX_LAYERS),
    ]


class NpvcParamDesc(C.Structure):
    _fields_ = [
        ("name", C.c_char * 96), ("offset", C.c_int64), ("size", C.c_int64), ("rank", C.c_int32),
        ("shape", C.c_int32 * 4), ("fan_in", C.c_int32), ("fan_out", C.c_int32), ("init", C.c_int32),
    ]


# every symbol include/npvc_b200.h declares: name -> (restype, argtypes)
_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
SYMBOLS = {
    "npvc_create": (C.c_int, [C.POINTER(NpvcArch), _I64, C.POINTER(_P)]),
    "npvc_destroy": (None, [_P]),
    "npvc_last_error": (C.c_char_p, []),
    "npvc_version": (C.c_char_p, []),
    "npvc_param_count": (_I64, [_P]),
    "npvc_param_tensors": (_I32, [_P]),
    "npvc_param_table": (C.c_int, [_P, C.POINTER(NpvcParamDesc), _I32]),
    "npvc_workspace_bytes": (_I64, [_P, _I64, _I32]),
    "npvc_plan_json": (C.c_char_p, [_P]),
    "npvc_plan_table": (_I64, [_P, C.c_char_p, C.POINTER(C.c_int32), _I64]),
    "npvc_launch_count": (_I64, [_P]),
    "npvc_pack_weights": (C.c_int, [_P, _P, _P, _I64, _P]),
    "npvc_encode": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _I64, _P]),
    "npvc_sample": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P]),
    "npvc_decode": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "npvc_loss_fwd_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _I32, _P, _I64, _P]),
    "npvc_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, _P]),
    "npvc_train_fwd_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32, _P, _I64, _P]),
    "npvc_normal_draw": (C.c_int, [_P, _P, _I64, _I64, _P, _P]),
    "npvc_adam_step_dev": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _P]),
    "npvc_tanhize_forward": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P]),
    "npvc_tanhize_backward": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P]),
    "npvc_unpack_records": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "npvc_profile_enable": (C.c_int, [_P, _I32]),
    "npvc_profile_json": (C.c_char_p, [_P]),
    "npvc_debug_buffer": (_I64, [_P, C.c_char_p, _P, _P, _I64, _P]),
}

_lib = None


def load():
    """Load the shared library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libnpvc_b200.so is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C vae_npvc_b200/csrc`). There is no CPU / PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().npvc_last_error().decode()


def check(rc):
    if rc != 0:
        msg = last_error()
        if rc == 1:
            raise ValueError(msg)
        raise RuntimeError("npvc_b200 error %d: %s" % (rc, msg))


def arch_struct(arch):
    """architecture-*.json dict (verbatim, unused keys tolerated) -> npvc_arch.
    Consumed keys: hwc, z_dim, y_dim, encoder.{kernel,stride,output}, generator.{hwc,kernel,
    stride,output} (model/vae.py:37-39,72-103)."""
    enc, gen = arch["encoder"], arch["generator"]
    for net in (enc, gen):   # model/vae.py:37-39 _sanity_check
        assert len(net["output"]) == len(net["kernel"]) == len(net["stride"])
    if len(enc["output"]) > MAX_LAYERS or len(gen["output"]) > MAX_LAYERS:
        raise ValueError("at most %d layers per network" % MAX_LAYERS)
    a = NpvcArch()
    a.in_h = arch["hwc"][0]
    if arch["hwc"][1] != 1 or arch["hwc"][2] != 1:
        raise ValueError("only [H,1,1] frames (width 1, 1 channel) are supported")
    a.z_dim, a.y_dim = arch["z_dim"], arch["y_dim"]
    a.n_enc = len(enc["output"])
    for i, (o, k, s) in enumerate(zip(enc["output"], enc["kernel"], enc["stride"])):
        if k[1] != 1 or s[1] != 1:
            raise ValueError("only [k,1] kernels / [s,1] strides are supported")
        a.enc_out[i], a.enc_kernel[i], a.enc_stride[i] = o, k[0], s[0]
    gh, gw, gc = gen["hwc"]
    if gw != 1:
        raise ValueError("generator.hwc width must be 1")
    a.gen_h, a.gen_c = gh, gc
    a.n_gen = len(gen["output"])
    for i, (o, k, s) in enumerate(zip(gen["output"], gen["kernel"], gen["stride"])):
        if k[1] != 1 or s[1] != 1:
            raise ValueError("only [k,1] kernels / [s,1] strides are supported")
        a.gen_out[i], a.gen_kernel[i], a.gen_stride[i] = o, k[0], s[0]
    return a


class Handle:
    """Owns one npvc_handle (host plan; device tables appear on the first device call)."""

    def __init__(self, arch, max_chunk=0):
        self.lib = load()
        self._h = C.c_void_p()
        self.arch_c = arch_struct(arch)
        check(self.lib.npvc_create(C.byref(self.arch_c), int(max_chunk), C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.npvc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def h(self):
        return self._h

    def param_count(self):
        return int(self.lib.npvc_param_count(self._h))

    def param_table(self):
        n = int(self.lib.npvc_param_tensors(self._h))
        arr = (NpvcParamDesc * n)()
        check(self.lib.npvc_param_table(self._h, arr, n))
        return [dict(name=d.name.decode(), offset=int(d.offset), size=int(d.size),
                     shape=tuple(d.shape[i] for i in range(d.rank)), fan_in=d.fan_in,
                     fan_out=d.fan_out, init=d.init) for d in arr]

    def workspace_bytes(self, n, train):
        return int(self.lib.npvc_workspace_bytes(self._h, int(n), 1 if train else 0))

    def plan(self):
        return json.loads(self.lib.npvc_plan_json(self._h).decode())

    def plan_table(self, name):
        import numpy as np
        n = int(self.lib.npvc_plan_table(self._h, name.encode(), None, 0))
        out = np.empty(n, np.int32)
        self.lib.npvc_plan_table(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out

    def launch_count(self):
        return int(self.lib.npvc_launch_count(self._h))

    def profile_enable(self, on=True):
        check(self.lib.npvc_profile_enable(self._h, 1 if on else 0))

    def profile(self):
        return json.loads(self.lib.npvc_profile_json(self._h).decode())
