"""Host-side launch recorder (test infrastructure): the product's host objects (plan.o, engine.o) linked against
tests/cuda_stub/cudart_stub.cpp instead of libcudart.  "Device" buffers are numpy arrays, kernels are not executed;
what engine.cu would have launched -- kernels, launch geometry, UmmaArgs, tensor maps, stream / event operations --
comes back as Python dicts.  Nothing here is on a product path.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from vae_npvc_b200 import lib as plib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "vae_npvc_b200", "csrc")
SRC = os.path.join(HERE, "cuda_stub", "cudart_stub.cpp")
OUT = os.path.join(HERE, "cuda_stub", "_build", "libnpvc_b200_hoststub.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


class StubTmap(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("rank", C.c_int32), ("base", C.c_uint64), ("dims", C.c_uint64 * 5),
                ("strides", C.c_uint64 * 5), ("box", C.c_uint32 * 5), ("estrides", C.c_uint32 * 5),
                ("interleave", C.c_int32), ("swizzle", C.c_int32), ("l2promo", C.c_int32), ("oob", C.c_int32)]


_UMMA_INTS = ["K", "N", "BN", "kblocks", "sw", "stages", "tmem_cols", "Rb", "Ra", "Ab", "FB", "TA", "RbH", "rows_tile", "frames",
              "m_tiles", "n_tiles", "acc_sets", "tapT", "tapC", "tapP", "b_tile_al", "d_sw", "rows_al", "tiles_per_split", "ld"]


class StubUmma(C.Structure):
    _fields_ = [(k, C.c_int32) for k in _UMMA_INTS] + [
        ("c_ptr", C.c_uint64), ("c_fs", C.c_int64), ("c_R", C.c_int32), ("c_rs", C.c_int32), ("c_off", C.c_int32),
        ("c_flen", C.c_int32), ("c_pred", C.c_int32), ("c_split", C.c_int32), ("out_ptr", C.c_uint64)] + [
        (k, C.c_int32) for k in ("a_boxes", "ln_on", "ln_store_c", "ln_L", "ln_Cn", "ln_out_flen", "ln_out_off", "b_res")] + [("ln_aout", C.c_uint64)]


class StubLaunch(C.Structure):
    _fields_ = [("name", C.c_char * 192), ("grid", C.c_uint32 * 3), ("block", C.c_uint32 * 3), ("smem", C.c_uint64),
                ("stream", C.c_uint64), ("seq", C.c_uint64), ("cluster_x", C.c_uint32), ("has_umma", C.c_int32),
                ("tmap", C.c_int32 * 4), ("u", StubUmma)]


class StubOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("stream", C.c_uint64), ("event", C.c_uint64), ("seq", C.c_uint64),
                ("ptr", C.c_uint64), ("bytes", C.c_uint64)]


_lib = None


def load():
    """Build (when stale) and load the stub-linked library."""
    global _lib
    if _lib is not None:
        return _lib
    objs = [os.path.join(CSRC, "plan.o"), os.path.join(CSRC, "engine.o")]
    for o in objs:
        if not os.path.exists(o):
            raise RuntimeError("%s is missing: run __graft_entry__.build() first" % o)
    deps = objs + [SRC, os.path.join(CSRC, "launch_args.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        # -Bsymbolic: the product objects must bind to THIS file's cuda* symbols even when a real libcudart (torch's) is
        # already loaded in the process
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I" + CUDA_INC, SRC] + objs + ["-o", OUT],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the CUDA runtime stub failed:\n" + r.stderr[-3000:])
    dll = C.CDLL(OUT)
    for name, (res, args) in plib.SYMBOLS.items():
        fn = getattr(dll, name); fn.restype = res; fn.argtypes = args
    dll.stub_n_launches.restype = dll.stub_n_tmaps.restype = dll.stub_n_ops.restype = C.c_int64
    dll.stub_get_launch.argtypes = [C.c_int64, C.POINTER(StubLaunch)]
    dll.stub_get_tmap.argtypes = [C.c_int64, C.POINTER(StubTmap)]
    dll.stub_get_op.argtypes = [C.c_int64, C.POINTER(StubOp)]
    _lib = dll
    return dll


def _aligned(nbytes, align=256):
    raw = np.empty(nbytes + align, np.uint8)                # pages are not touched: a 5 GB workspace costs nothing
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class Recording(object):
    """What one call launched."""
    def __init__(self, dll, ws_lo, ws_hi, rc, err):
        self.rc, self.err, self.ws_lo, self.ws_hi = rc, err, ws_lo, ws_hi
        self.launches, self.tmaps, self.ops = [], [], []
        for i in range(dll.stub_n_launches()):
            L = StubLaunch(); dll.stub_get_launch(i, C.byref(L))
            d = dict(name=L.name.decode(), grid=tuple(L.grid), block=tuple(L.block), smem=int(L.smem), stream=int(L.stream),
                     seq=int(L.seq), cluster_x=int(L.cluster_x), tmap=tuple(L.tmap), umma=None)
            if L.has_umma:
                d["umma"] = {k: int(getattr(L.u, k)) for k, _ in StubUmma._fields_}
            self.launches.append(d)
        for i in range(dll.stub_n_tmaps()):
            t = StubTmap(); dll.stub_get_tmap(i, C.byref(t))
            r = t.rank
            self.tmaps.append(dict(dtype=t.dtype, rank=r, base=int(t.base), dims=list(t.dims)[:r], strides=list(t.strides)[:r - 1],
                                   box=list(t.box)[:r], estrides=list(t.estrides)[:r], interleave=t.interleave, swizzle=t.swizzle,
                                   l2promo=t.l2promo, oob=t.oob))
        for i in range(dll.stub_n_ops()):
            o = StubOp(); dll.stub_get_op(i, C.byref(o))
            self.ops.append(dict(kind=o.kind, stream=int(o.stream), event=int(o.event), seq=int(o.seq), ptr=int(o.ptr), bytes=int(o.bytes)))


def record_loss_fwd_bwd(arch, n, env=None, with_grad=True, max_chunk=0, fail_cluster=0, calls=1):
    """npvc_loss_fwd_bwd over n frames under the recording stub (library switches from `env`); the recording covers
    the LAST of `calls` calls."""
    dll = load()
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        a = plib.arch_struct(arch)
        h = C.c_void_p()
        rc = dll.npvc_create(C.byref(a), int(max_chunk), C.byref(h))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc:
        raise ValueError(dll.npvc_last_error().decode())
    try:
        npar = dll.npvc_param_count(h)
        z, H = arch["z_dim"], arch["hwc"][0]
        theta = np.zeros(npar, np.float32); grad = np.zeros(npar, np.float32)
        x = np.zeros((n, H), np.float32); y = np.zeros(n, np.int64); eps = np.zeros((n, z), np.float32)
        losses = np.zeros(3, np.float32)
        nbytes = dll.npvc_workspace_bytes(h, n, 1)
        ws = _aligned(nbytes)
        dll.stub_fail_cluster_launches(int(fail_cluster))
        for _ in range(calls):
            dll.stub_reset()
            rc = dll.npvc_loss_fwd_bwd(h, theta.ctypes.data, x.ctypes.data, y.ctypes.data, eps.ctypes.data, n, None, None, None, None,
                                       losses.ctypes.data, grad.ctypes.data if with_grad else None, 1, ws.ctypes.data, nbytes, None)
        rec = Recording(dll, ws.ctypes.data, ws.ctypes.data + nbytes, rc, dll.npvc_last_error().decode() if rc else "")
        rec.launch_count = int(dll.npvc_launch_count(h))
        rec.plan, rec.train = json.loads(dll.npvc_plan_json(h).decode()), True
        return rec
    finally:
        dll.stub_fail_cluster_launches(0)
        dll.npvc_destroy(h)


def record_encode_decode(arch, n, env=None, max_chunk=0):
    """npvc_encode followed by npvc_decode (the convert.py path) over n frames under the recording stub."""
    dll = load()
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        a = plib.arch_struct(arch)
        h = C.c_void_p()
        rc = dll.npvc_create(C.byref(a), int(max_chunk), C.byref(h))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc:
        raise ValueError(dll.npvc_last_error().decode())
    try:
        npar = dll.npvc_param_count(h)
        z, H = arch["z_dim"], arch["hwc"][0]
        theta = np.zeros(npar, np.float32)
        x = np.empty((n, H), np.float32); y = np.zeros(n, np.int64)
        mu = np.empty((n, z), np.float32); lv = np.empty((n, z), np.float32); xh = np.empty((n, H), np.float32)
        nbytes = dll.npvc_workspace_bytes(h, n, 0)
        ws = _aligned(nbytes)
        dll.stub_reset()
        rc = dll.npvc_pack_weights(h, theta.ctypes.data, ws.ctypes.data, nbytes, None)
        rc = rc or dll.npvc_encode(h, theta.ctypes.data, x.ctypes.data, n, mu.ctypes.data, lv.ctypes.data, ws.ctypes.data, nbytes, None)
        rc = rc or dll.npvc_decode(h, theta.ctypes.data, mu.ctypes.data, y.ctypes.data, n, xh.ctypes.data, ws.ctypes.data, nbytes, None)
        rec = Recording(dll, ws.ctypes.data, ws.ctypes.data + nbytes, rc, dll.npvc_last_error().decode() if rc else "")
        rec.launch_count = int(dll.npvc_launch_count(h))
        rec.plan, rec.train = json.loads(dll.npvc_plan_json(h).decode()), False
        return rec
    finally:
        dll.npvc_destroy(h)


def buffer_ranges(plan, chunk, train):
    """[(name, first float, floats per frame, is-split)] of the per-frame workspace buffers for a pass over `chunk` frames
    (Plan::buf_offset of plan.cpp restated: 64-float aligned buffers behind the operand-pack arena)."""
    rup = lambda v: (v + 63) // 64 * 64
    off, out, start = rup(plan["arena_w"]), [], {}
    for i, b in enumerate(plan["bufs"]):
        own = rup(b["fixed"] + b["per_frame"] * chunk)
        if b.get("alias", -1) >= 0:                      # a tenant of another buffer: same start, its own extent, no storage
            if not (b["train_only"] and not train):
                out.append((b["name"], start[b["alias"]], start[b["alias"]] + own, b["per_frame"], b["split"]))
            continue
        sz = 0 if ((b["train_only"] and not train) or b.get("elide")) else own
        start[i] = off
        out.append((b["name"], off, off + sz, b["per_frame"], b["split"]))
        off += sz
    return out
