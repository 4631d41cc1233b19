"""CPU tests of the host side of the product: C-ABI surface, parameter table, launch plan
(executed by the numpy interpreter against the oracle), error behaviour without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import plan_interp as PI
from parity_util import check_branches, lrelu_branches
from oracle import convvae_ref as R
from vae_npvc_b200 import lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(__file__)), "include", "npvc_b200.h")


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))


def test_library_exports_every_declared_symbol():
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(npvc_[a-z_0-9]+)\s*\(", src))
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    dll = C.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name
    assert b"sm_100a" in lib.load().npvc_version()


def test_param_table_matches_reference_variable_order(arch):
    h = lib.Handle(arch)
    tab = h.param_table()
    specs = R.param_specs(arch)
    assert h.param_count() == 939162 and len(tab) == 44
    off = 0
    for t, (name, shape, fi, fo, kind) in zip(tab, specs):
        assert t["name"] == name and t["shape"] == tuple(shape) and t["offset"] == off
        assert (t["fan_in"], t["fan_out"]) == (fi, fo) and t["init"] == {"glorot": 0, "zeros": 1, "ones": 2}[kind]
        off += t["size"]


@pytest.mark.parametrize("n,umma", [(1, 0), (3, 0), (3, 1), (8, 1)])
def test_plan_matches_oracle(arch, n, umma, monkeypatch):
    """umma=0: CUDA-core plan, exact in fp64.  umma=1: the tcgen05 routing with tf32 hi+lo operand
    planes and packs (bf16x3: hi.hi + hi.lo + lo.hi, ~2^-17 per product, so fp64 agreement drops to ~1e-5)."""
    monkeypatch.setenv("NPVC_UMMA", str(umma))
    tol_out, tol_g = (1e-12, 1e-9) if not umma else (3e-5, 5e-5)
    h = lib.Handle(arch)
    plan = h.plan()
    tables = {k: h.plan_table(k) for k in ("pack_src", "pack16_src", "pack_list", "unpack_ptr", "unpack_idx")}
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, n)
    it = PI.Interp(plan, tables, R.flatten_params(arch, P, np.float64), n, x, y, eps)
    out = it.loss_fwd_bwd()
    pos = lrelu_branches(it.buf, arch, P, n)                      # the oracle differentiates with the plan's lrelu branches
    check_branches(pos, R.forward(arch, P, x, y, eps, with_acts=True)["acts"])
    ref = R.forward(arch, P, x, y, eps, with_grads=True, lrelu_pos=pos)
    assert (sum(op.get("umma", 0) for op in plan["ops"]) > 10) == bool(umma)
    for k in ("mu", "lv", "z", "xh"):
        assert rel(out[k], ref[k]) < tol_out, k
    for k in ("D_KL", "logP", "G"):
        assert rel(out[k], ref[k]) < (1e-6 if not umma else 1e-5), k
    gref = R.flatten_params(arch, ref["grads"], np.float64)
    for t in h.param_table():
        sl = slice(t["offset"], t["offset"] + t["size"])
        assert rel(out["grad"][sl], gref[sl]) < tol_g, t["name"]


def test_pack_tables_are_consistent(arch, monkeypatch):
    monkeypatch.setenv("NPVC_UMMA", "0")
    h = lib.Handle(arch)
    plan = h.plan()
    src, ptr, idx = (h.plan_table(k) for k in ("pack_src", "unpack_ptr", "unpack_idx"))
    assert len(src) == plan["arena_w"] and len(ptr) == plan["n_params"] + 1
    assert src.max() < plan["n_params"] and src.min() >= -1
    assert idx.max() < plan["arena_dw"] and (np.diff(ptr) >= 0).all() and ptr[-1] == len(idx)
    # every conv / dense kernel element receives its gradient from the packed arena; the
    # 1025-tap kernel gathers one full Toeplitz diagonal (513 entries) per tap and channel
    tab = {t["name"]: t for t in h.param_table()}
    t3 = tab["Generator/conv2d_transpose_3/kernel"]
    cnt = np.diff(ptr)[t3["offset"]:t3["offset"] + t3["size"]].reshape(1025, 8)
    assert cnt[512].min() == 513 and cnt[0].max() == 1 and cnt.sum() == 513 * 513 * 8
    e0 = tab["Encoder/Conv2d-0/Conv2d-0/kernel"]
    assert (np.diff(ptr)[e0["offset"]:e0["offset"] + e0["size"]] == 1).all()
    lnp = tab["Encoder/Conv2d-0/layernorm.scale"]
    assert (np.diff(ptr)[lnp["offset"]:lnp["offset"] + lnp["size"]] == 0).all()


def test_workspace_sizes(arch):
    h = lib.Handle(arch, 4096)
    a, b = h.workspace_bytes(100, False), h.workspace_bytes(100, True)
    assert 0 < a < b
    assert h.workspace_bytes(4096, True) == h.workspace_bytes(100000, True)      # chunked above max_chunk
    assert h.workspace_bytes(200, True) > b


def test_tensor_path_plan_features(arch, monkeypatch):
    """Plan-level features of the tensor path: operand planes, tap metadata of the conv-shaped views, the
    parity-split dgrad of the 8-channel transposed conv (interleaved views, padded gradient frame), the
    speaker term inside the merge GEMM."""
    monkeypatch.setenv("NPVC_UMMA", "1")
    plan = lib.Handle(arch).plan()
    ops = {o["name"]: o for o in plan["ops"]}
    bufs = {b["name"]: b for b in plan["bufs"]}
    for name in ("a_e0", "a_g2", "dc_e1", "dc_g2", "zs", "hm", "dhm", "dhz", "dxh"):
        assert bufs[name]["split"] == 1 and bufs[name]["per_frame"] % 8 == 0, name        # 16-byte aligned lo plane
    for name in ("c_e0", "da_g2", "xh", "mu"):
        assert bufs[name]["split"] == 0, name
    assert ops["conv_e1"]["tap"] == [7, 16, 3] and ops["convT_g2"]["tap"] == [3, 16, 1]
    assert ops["dgrad_g1"]["tap"] == [7, 16, 3] and ops["dgrad_e1"]["tap"] == [3, 32, 1]
    assert ops["convT_g3"]["tap"] == [0, 0, 0] and "dgrad_g2" not in ops
    ev, od = ops["dgrad_g2_even"], ops["dgrad_g2_odd"]
    assert ev["tap"] == od["tap"] == [4, 16, 3] and ev["K"] == od["K"] == 56
    assert (ev["A"]["R"], od["A"]["R"]) == (86, 85) and ev["A"]["rs"] == od["A"]["rs"] == 48
    assert (ev["A"]["off"], od["A"]["off"]) == (0, 24) and (ev["C"]["off"], od["C"]["off"]) == (0, 16) and ev["C"]["rs"] == 32
    # the last 16-element tap of the last row stays inside the frame's own plane (zeros), never a neighbour's bits
    last = ev["A"]["off"] + (ev["A"]["R"] - 1) * ev["A"]["rs"] + ev["tap"][0] * ev["tap"][1]
    assert last <= bufs["dc_g2"]["per_frame"] == 4144
    # merge: K = z + padded one-hot columns, no table / segmented-sum ops left
    assert ops["merge"]["K"] == 128 + 16 and ops["wgrad_merge_z"]["K"] == 144 and "segsum_dhm" not in ops and "zcat" in ops


def test_bad_architecture_raises_value_error(arch):
    bad = dict(arch); bad["encoder"] = dict(arch["encoder"]); bad["encoder"]["output"] = [16, 32, 64, 128, 255]
    with pytest.raises(ValueError):
        lib.Handle(bad)
    bad2 = dict(arch); bad2["generator"] = dict(arch["generator"]); bad2["generator"]["output"] = [32, 16, 8]
    with pytest.raises(AssertionError):          # model/vae.py:37-39 _sanity_check
        lib.Handle(bad2)
    with pytest.raises(ValueError):              # frames per internal pass are bounded (32-bit row / tile counts)
        lib.Handle(arch, max_chunk=1 << 21)
    assert lib.Handle(arch, max_chunk=1 << 20).workspace_bytes(1 << 22, True) == lib.Handle(arch, max_chunk=1 << 20).workspace_bytes(1 << 20, True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_fails_loudly(arch):
    from vae_npvc_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(arch)
    h = lib.Handle(arch)
    buf = np.zeros(1024, np.float32)
    rc = h.lib.npvc_pack_weights(h.h, buf.ctypes.data, 256 * 1024 * 1024 * 16, 1 << 40, None)
    assert rc != 0 and ("CUDA" in lib.last_error() or "device" in lib.last_error())


@pytest.mark.parametrize("umma", [0, 1])
def test_alternative_architectures(alt_arch, umma, monkeypatch):
    """Generic plan: other kernel sizes / strides / asymmetric pads / padded channel counts."""
    monkeypatch.setenv("NPVC_UMMA", str(umma))
    tol = 1e-10 if not umma else 5e-5          # bf16x3 planes: ~2^-17 per product, amplified by Layernorm over few elements
    from oracle import convvae_loops as L
    h = lib.Handle(alt_arch)
    plan = h.plan()
    tables = {k: h.plan_table(k) for k in ("pack_src", "pack16_src", "pack_list", "unpack_ptr", "unpack_idx")}
    P = R.init_params(alt_arch, 0)
    x, y, eps = R.make_inputs(alt_arch, 3)
    ref = R.forward(alt_arch, P, x, y, eps, with_grads=True)
    lo = L.forward(alt_arch, P, x, y, eps)
    assert rel(ref["xh"], lo["xh"]) < 1e-12 and rel(ref["mu"], lo["mu"]) < 1e-12          # both oracle formulations
    out = PI.Interp(plan, tables, R.flatten_params(alt_arch, P, np.float64), 3, x, y, eps).loss_fwd_bwd()
    for k in ("mu", "lv", "z", "xh"):
        assert rel(out[k], ref[k]) < tol, k
    assert rel(out["grad"], R.flatten_params(alt_arch, ref["grads"], np.float64)) < tol
    assert h.param_count() == R.n_params(alt_arch)


def _random_arch(rs):
    """A valid random architecture inside the plan's documented limits (kernel >= stride, channel counts that are
    multiples of 4, a stride-1 kernel > 3 as the last generator layer): random depths, strides 1..4, even / odd
    kernels (asymmetric SAME pads and crops), channel counts that need padding."""
    n_gen = rs.randint(2, 4)
    strides = [int(rs.choice([2, 3, 4])) for _ in range(n_gen - 1)] + [1]
    gen_h = int(rs.randint(3, 12))
    in_h = gen_h * int(np.prod(strides))
    gk = [[int(s + rs.randint(0, 6)), 1] for s in strides[:-1]]
    gk += [[int(rs.choice([5, 6, 2 * in_h - 1, in_h // 2 * 2 + 1, in_h + 4])), 1]]
    gout = [int(rs.choice([4, 8, 12, 16, 24, 32])) for _ in range(n_gen - 1)] + [1]
    n_enc = rs.randint(1, 4)
    es = [int(rs.choice([1, 2, 3, 4])) for _ in range(n_enc)]
    return {"hwc": [in_h, 1, 1], "z_dim": int(rs.choice([4, 8, 12, 32])), "y_dim": int(rs.randint(1, 7)),
            "encoder": {"kernel": [[int(s + rs.randint(0, 6)), 1] for s in es], "stride": [[s, 1] for s in es],
                        "output": [int(rs.choice([4, 8, 12, 16, 24, 32, 64])) for _ in range(n_enc)]},
            "generator": {"hwc": [gen_h, 1, int(rs.choice([3, 5, 8, 16, 33]))], "kernel": gk,
                          "stride": [[s, 1] for s in strides], "output": gout}}


@pytest.mark.parametrize("seed", range(12))
def test_random_architectures(seed, monkeypatch):
    """Plan generality: the launch plan of a random architecture, executed by the numpy interpreter in both operand
    formats (fp32 CUDA-core packs / bf16 hi-lo planes), against the oracle -- outputs and every gradient."""
    arch = _random_arch(np.random.RandomState(seed))
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, 3)
    ref = R.forward(arch, P, x, y, eps, with_grads=True)
    gref = R.flatten_params(arch, ref["grads"], np.float64)
    for umma, tol in ((0, 1e-10), (1, 2e-4)):
        monkeypatch.setenv("NPVC_UMMA", str(umma))
        h = lib.Handle(arch)
        tables = {k: h.plan_table(k) for k in ("pack_src", "pack16_src", "pack_list", "unpack_ptr", "unpack_idx")}
        out = PI.Interp(h.plan(), tables, R.flatten_params(arch, P, np.float64), 3, x, y, eps).loss_fwd_bwd()
        for k in ("mu", "lv", "z", "xh"):
            assert rel(out[k], ref[k]) < tol, (umma, k, arch)
        assert rel(out["grad"], gref) < tol, (umma, arch)
        assert h.param_count() == R.n_params(arch)
