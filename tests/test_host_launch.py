"""Host launch logic of engine.cu without a GPU: the product's host objects run against a recording CUDA-runtime stub
(tests/hoststub.py, tests/cuda_stub/) and every launch is checked against the invariants the kernels rely on --
shared-memory and TMEM budgets, tile geometry, the cuTensorMapEncodeTiled rules, tensor extents inside the caller's
workspace, bytes per TMA box == bytes the kernels expect per stage, fork / join structure of the internal streams --
for the reference architecture, the alternative and random ones, small to full batch sizes and every library switch
(including the ones that have not run on a GPU yet)."""
import numpy as np
import pytest

import hoststub as HS
from conftest import ALT_ARCHS
from test_plan import _random_arch
from vae_npvc_b200 import vcc2016_vae_arch

SMEM_MAX = 227 * 1024
BF16, SWZ_BYTES = 9, {1: 32, 2: 64, 3: 128}        # CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_{32,64,128}B


def _rup(v, m):
    return (v + m - 1) // m * m


def check_tmap(t, rec, what):
    """cuTensorMapEncodeTiled's documented requirements + the tensor lies inside the caller's workspace."""
    assert t["dtype"] == BF16 and t["interleave"] == 0 and t["swizzle"] in SWZ_BYTES, what
    assert t["base"] % 16 == 0, (what, "global address must be 16-byte aligned")
    assert all(1 <= d <= 2 ** 32 for d in t["dims"]), (what, t["dims"])
    assert all(s % 16 == 0 and 0 < s < 2 ** 40 for s in t["strides"]), (what, t["strides"])
    assert all(1 <= b <= 256 for b in t["box"]) and all(e == 1 for e in t["estrides"]), (what, t["box"])
    inner = t["box"][0] * 2
    assert inner % 16 == 0 and inner <= SWZ_BYTES[t["swizzle"]], (what, "inner box bytes vs swizzle span", t["box"], t["swizzle"])
    last = t["base"] + t["dims"][0] * 2 + sum((d - 1) * s for d, s in zip(t["dims"][1:], t["strides"]))
    assert rec.ws_lo <= t["base"] and last <= rec.ws_hi, (what, "tensor extent leaves the workspace", last - rec.ws_hi)


def box_bytes(t):
    return 2 * int(np.prod(t["box"]))


def check_recording(rec, n):
    assert rec.rc == 0, rec.err
    assert rec.launch_count >= len(rec.launches) > 0
    for L in rec.launches:
        name, u = L["name"], L["umma"]
        assert all(g >= 1 for g in L["grid"]) and 1 <= L["block"][0] <= 1024 and L["smem"] <= SMEM_MAX, L
        assert L["grid"][1] <= 65535 and L["grid"][2] <= 65535, L
        if L["cluster_x"] > 1:
            assert L["cluster_x"] == 2 and L["grid"][0] % 2 == 0, L
        if u is None:
            continue
        tm = [rec.tmaps[i] for i in L["tmap"]]
        assert all(i >= 0 for i in L["tmap"]), (name, "kernel launched without encoded tensor maps")
        for i, t in enumerate(tm):
            check_tmap(t, rec, (name, i, u["K"], u["N"]))
        pair = "ILb1" in name
        assert pair == (L["cluster_x"] == 2), name
        assert u["frames"] >= 1 and u["m_tiles"] >= 1 and u["rows_tile"] <= 128 and u["rows_tile"] == u["RbH"] * u["Ab"] * u["FB"], u
        tc = u["tmem_cols"]
        assert tc in (32, 64, 128, 256, 512), u
        if "umma_fwd_kernel_t" in name:
            BN, sw, st = u["BN"], u["sw"], u["stages"]
            assert BN % 16 == 0 and 16 <= BN <= 256 and u["acc_sets"] in (1, 2, 4) and tc >= u["acc_sets"] * 2 * BN, u
            groups = (L["block"][0] - 64) // 128
            issuer2 = L["block"][0] - 64 - 128 * groups                # second MMA issuer warp: tap mode, >= 2 accumulator sets / stages / tiles per CTA
            assert issuer2 in (0, 32) and groups in (1, 2, 4) and groups <= u["acc_sets"] and u["acc_sets"] % groups == 0, (L["block"], u)
            if issuer2:
                assert u["tapT"] > 0 and u["acc_sets"] == 4 and u["stages"] >= 4 and u["m_tiles"] > L["grid"][0], (L["block"], u)
            assert u["n_tiles"] * BN >= u["N"] and (u["n_tiles"] - 1) * BN < u["N"], u
            assert tm[0]["rank"] == 4 and tm[1]["rank"] == 4 and tm[2]["rank"] == 2 and tm[3]["rank"] == 2
            if u["tapT"] > 0:                                     # tap mode
                assert not pair and sw == 2 * u["tapC"] and sw in (32, 64, 128) and st >= 2 and u["n_tiles"] == 1, u
                assert u["b_tile_al"] % 1024 == 0 and u["b_tile_al"] >= BN * sw, u
                ring = u["tapT"] * 2 * u["b_tile_al"] + st * u["tapP"] * 2 * 128 * sw
                assert box_bytes(tm[0]) == u["rows_tile"] * sw and box_bytes(tm[2]) == BN * sw, (u, tm[0]["box"], tm[2]["box"])
                assert L["grid"][0] <= 148 and L["grid"][0] <= u["m_tiles"]
            else:                                                 # window mode (single CTA or CTA pair)
                assert sw in (64, 128) and st >= 1 and u["kblocks"] == -(-u["K"] // (sw // 2)), u
                bt = (BN // 2 if pair else BN) * sw
                if u["b_res"]:                                    # resident weights in front of a ring of activation stages
                    assert not pair and u["n_tiles"] == 1 and st >= 3 and bt % 1024 == 0 and L["grid"][0] < u["m_tiles"], u
                    ring = u["kblocks"] * 2 * bt + st * 2 * 128 * sw
                else:
                    ring = st * (2 * 128 * sw + 2 * bt)
                assert box_bytes(tm[0]) == u["rows_tile"] * sw and box_bytes(tm[2]) == bt, (u, tm[0]["box"], tm[2]["box"])
                if pair:
                    assert sw == 128 and bt % 1024 == 0 and u["m_tiles"] >= 2 and L["grid"][0] <= 148, u
                    assert L["grid"][0] // 2 <= -(-u["m_tiles"] // 2) * u["n_tiles"], (L["grid"], u)
                else:
                    assert L["grid"][0] <= min(148, u["m_tiles"] * u["n_tiles"]), (L["grid"], u)
            # 1024-byte alignment slack + ring + barriers + TMEM slot + bias staging of up to 4 epilogue groups
            assert L["smem"] >= 1023 + ring + 8 * (2 * st + 10) + 16 + 4096 + (32768 if u["ln_on"] else 0), (name, L["smem"], ring, u)
            assert bool(u["ln_on"]) == ("umma_fwd_kernel_tILb0ELb1E" in name), (name, u["ln_on"])     # the LNF instantiation <-> ln.on
            if u["ln_on"]:
                # Layernorm epilogue: whole frames per tile, one N tile, plane rows of the frame inside its (padded) output frame
                assert not pair and u["Ra"] == 1 and u["n_tiles"] == 1 and u["N"] % 8 == 0 and u["ln_L"] == u["c_R"] * u["N"], u
                assert u["ln_out_off"] % 8 == 0 and u["ln_out_flen"] % 8 == 0 and u["ln_out_off"] + u["ln_L"] <= u["ln_out_flen"], u
                assert not rec.train or u["ln_store_c"] == 1, u
            assert tm[0]["box"] == tm[1]["box"] and tm[2]["box"] == tm[3]["box"]
        else:                                                     # weight-gradient kernel
            BN, dsw, st, ral = u["BN"], u["d_sw"], u["stages"], u["rows_al"]
            assert dsw in (32, 64, 128) and 16 <= BN <= 256 and BN % 16 == 0 and tc >= 2 * BN, u
            assert ral % 16 == 0 and ral >= u["rows_tile"] and ral - u["rows_tile"] < 16 and st >= 1, u
            dw = dsw // 2
            if pair:
                assert dsw == 128 and BN % (2 * dw) == 0, u
                d_boxes = BN // dw // 2
            else:
                d_boxes = -(-BN // dw)
            assert u["a_boxes"] == (1 if (not pair and u["K"] <= 64) else 2), u
            stage = 2 * (u["a_boxes"] * ral * 128) + 2 * d_boxes * _rup(ral * dsw, 1024)
            assert L["smem"] >= 1023 + st * stage + 8 * (2 * st + 1) + 8 + 4, (name, L["smem"], st, stage, u)
            gx, gy, gz = L["grid"]
            assert gx * 128 >= u["K"] and (gx - (2 if pair else 1)) * 128 < u["K"], (L["grid"], u)
            assert gy * BN >= u["N"] and (gy - 1) * BN < u["N"] and gz * u["tiles_per_split"] >= u["m_tiles"] > (gz - 1) * u["tiles_per_split"], (L["grid"], u)
            assert box_bytes(tm[0]) == u["rows_tile"] * 128 and box_bytes(tm[2]) == u["rows_tile"] * dsw, (u, tm[0]["box"], tm[2]["box"])
            assert all(t["rank"] == 4 for t in tm) and u["out_ptr"] >= rec.ws_lo and u["out_ptr"] < rec.ws_hi
    check_streams(rec)


def check_streams(rec, caller=0):
    """Everything issued on an internal stream happens-before the end of the call on the caller's stream (event record /
    wait edges, transitively), and nothing on an internal stream starts before the call's first operation."""
    clock, snap, last = {}, {}, {}                          # clock[s][t] = last seq of stream t that stream s has observed
    for o in sorted(rec.ops, key=lambda o: o["seq"]):
        s = o["stream"]
        c = clock.setdefault(s, {})
        if o["kind"] == 1:                                  # event record: what the stream has done / seen so far
            e = dict(c); e[s] = o["seq"]
            snap[o["event"]] = e
        elif o["kind"] == 2:                                # stream wait
            assert o["event"] in snap, "wait on an event that was never recorded"
            for k, v in snap[o["event"]].items():
                c[k] = max(c.get(k, -1), v)
        else:                                               # launch / memset / memcpy
            if s != caller:
                assert caller in c, "internal stream %#x ran without a fork from the caller's stream" % s
            last[s] = o["seq"]
        c[s] = o["seq"]
    seen = clock.get(caller, {})
    for s, q in last.items():
        if s != caller:
            assert seen.get(s, -1) >= q, "work on internal stream %#x is not joined into the caller's stream" % s


def observed_before(rec, seq, caller=0):
    """{stream: last seq of that stream the caller's stream has observed} just before the caller's operation `seq`."""
    clock, snap = {}, {}
    for o in sorted(rec.ops, key=lambda o: o["seq"]):
        s = o["stream"]; c = clock.setdefault(s, {})
        if s == caller and o["seq"] == seq:
            return dict(c)
        if o["kind"] == 1:
            e = dict(c); e[s] = o["seq"]; snap[o["event"]] = e
        elif o["kind"] == 2:
            for k, v in snap[o["event"]].items():
                c[k] = max(c.get(k, -1), v)
        c[s] = o["seq"]
    raise AssertionError("no operation %d on the caller's stream" % seq)


def test_packs_run_beside_the_first_layer():
    """Training call: the fused first layer reads only the fp32 pack, so the speaker table and the bf16 planes are
    issued on the side stream; the first tensor-path GEMM (their first reader) waits for them, the first layer does not.
    With NPVC_OVERLAP=0 everything stays on the caller's stream."""
    r = HS.record_loss_fwd_bwd(vcc2016_vae_arch(), 16384)
    by = lambda frag: [L for L in r.launches if frag in L["name"]]
    p16, e0, um = by("pack16_kernel")[0], by("e0_fwd_kernel")[0], [L for L in r.launches if L["umma"]][0]
    ptab = by("fewrows_fwd_kernel")[0]
    assert p16["stream"] != 0 and ptab["stream"] == p16["stream"] and e0["stream"] == 0 and um["stream"] == 0
    pk = [L for L in r.launches if "pack_list_kernel" in L["name"] or "npvc::pack_kernel" in L["name"]][0]
    assert pk["stream"] == 0 and pk["seq"] < e0["seq"]
    assert observed_before(r, e0["seq"]).get(p16["stream"], -1) < ptab["seq"]            # not waited for
    assert observed_before(r, um["seq"]).get(p16["stream"], -1) >= max(p16["seq"], ptab["seq"])
    r0 = HS.record_loss_fwd_bwd(vcc2016_vae_arch(), 16384, {"NPVC_OVERLAP": "0"})
    assert {L["stream"] for L in r0.launches} == {0}
    rf = HS.record_loss_fwd_bwd(vcc2016_vae_arch(), 64, {"NPVC_FUSE": "0"})                # no fused first layer: joined at once
    first = [L for L in rf.launches if L["stream"] == 0 and "pack" not in L["name"]][0]
    side = [L for L in rf.launches if L["stream"] != 0 and L["seq"] < first["seq"]]
    assert side and observed_before(rf, first["seq"]).get(side[0]["stream"], -1) >= max(L["seq"] for L in side)


SWITCHES = {
    "default": {},
    "single_cta": {"NPVC_PAIR": "0"},
    "pair_wide": {"NPVC_PAIR": "2"},
    "wgrad_pair": {"NPVC_WGRAD_PAIR": "1"},
    "wgrad_pair_256": {"NPVC_WGRAD_PAIR": "2"},
    "everything": {"NPVC_PAIR": "2", "NPVC_WGRAD_PAIR": "2"},
    "window_only": {"NPVC_UMMA_TAP": "0"},
    "no_overlap": {"NPVC_OVERLAP": "0"},
    "no_resident_weights": {"NPVC_UMMA_BRES": "0"},
    "one_mma_issuer": {"NPVC_UMMA_DUAL": "0"},
    "ln_bwd_prefetch": {"NPVC_PREFETCH_MIN": "0", "NPVC_PREFETCH_E0_MIN": "0"},
}


@pytest.mark.parametrize("switch", sorted(SWITCHES))
@pytest.mark.parametrize("n", [1, 37, 64, 300, 4096, 16384])
def test_reference_architecture_launches(n, switch):
    check_recording(HS.record_loss_fwd_bwd(vcc2016_vae_arch(), n, SWITCHES[switch]), n)


@pytest.mark.parametrize("switch", ["default", "everything"])
@pytest.mark.parametrize("name", sorted(ALT_ARCHS))
def test_alternative_architecture_launches(name, switch):
    for n in (3, 29, 2048):
        check_recording(HS.record_loss_fwd_bwd(ALT_ARCHS[name], n, SWITCHES[switch]), n)


@pytest.mark.parametrize("seed", range(12))
def test_random_architecture_launches(seed):
    arch = _random_arch(np.random.RandomState(seed))
    for n, sw in ((5, "default"), (1000, "default"), (1000, "everything")):
        check_recording(HS.record_loss_fwd_bwd(arch, n, SWITCHES[sw]), n)


def test_chunked_and_repeated_calls():
    """Calls larger than max_chunk run as chunks (the tensor-map cache must follow the changing frame count of the last
    chunk), a second call reuses the cached maps, forward-only calls launch no backward work."""
    arch = vcc2016_vae_arch()
    r = HS.record_loss_fwd_bwd(arch, 40, max_chunk=16)
    check_recording(r, 40)
    frames = sorted({L["umma"]["frames"] for L in r.launches if L["umma"]})
    assert frames == [8, 16]
    r1 = HS.record_loss_fwd_bwd(arch, 16384)
    r2 = HS.record_loss_fwd_bwd(arch, 16384, calls=2)
    assert r2.rc == 0 and r2.tmaps == [], "a repeated call on the same buffers must hit the tensor-map cache"
    key = lambda L: (L["name"], L["grid"], L["block"], L["smem"], L["cluster_x"], L["stream"] != 0)
    assert [key(L) for L in r2.launches] == [key(L) for L in r1.launches]
    r3 = HS.record_loss_fwd_bwd(arch, 64, with_grad=False)
    check_recording(r3, 64)
    assert not any("wgrad" in L["name"] or "ln_bwd" in L["name"] or "unpack" in L["name"] for L in r3.launches)
    r4 = HS.record_loss_fwd_bwd(arch, 40000, max_chunk=16384)                              # 2 full chunks + a ragged one
    check_recording(r4, 40000)
    assert len({L["stream"] for L in r4.launches}) == 2                                    # caller, wgrad side stream


def test_default_rule_pairs_only_the_measured_shapes():
    """pair_wanted (engine.cu): at cfg2 size the CTA-pair form runs G3 forward / dgrad, E4, heads, E4 dgrad and the merge
    dgrad (BN >= 128, >= 8 k-blocks, >= 64 M tiles); small batches keep the single-CTA form."""
    arch = vcc2016_vae_arch()
    big = HS.record_loss_fwd_bwd(arch, 16384)
    pairs = sorted((L["umma"]["K"], L["umma"]["N"]) for L in big.launches if L["cluster_x"] == 2 and "umma_fwd" in L["name"])
    assert pairs == sorted([(4104, 513), (513, 4104), (896, 256), (768, 256), (768, 384), (1672, 128)]), pairs
    wpairs = [(L["umma"]["K"], L["umma"]["N"]) for L in big.launches if L["cluster_x"] == 2 and "umma_wgrad" in L["name"]]
    assert wpairs == [(4104, 513)], wpairs                     # the pair weight-gradient kernel: the last generator layer only
    small = HS.record_loss_fwd_bwd(arch, 64)
    assert not any(L["cluster_x"] == 2 for L in small.launches)


def test_refused_cluster_launch_falls_back_to_the_single_cta_form():
    arch = vcc2016_vae_arch()
    r = HS.record_loss_fwd_bwd(arch, 16384, fail_cluster=1)
    check_recording(r, 16384)
    assert not any(L["cluster_x"] == 2 for L in r.launches)
    ref = HS.record_loss_fwd_bwd(arch, 16384, {"NPVC_PAIR": "0"})
    assert [(L["name"], L["grid"], L["block"], L["smem"]) for L in r.launches] == [(L["name"], L["grid"], L["block"], L["smem"]) for L in ref.launches]
    r2 = HS.record_loss_fwd_bwd(arch, 16384, {"NPVC_PAIR": "2"}, fail_cluster=1)           # an explicit request reports the failure
    assert r2.rc != 0


@pytest.mark.parametrize("n", [1, 23, 700, 16384, 40000])
def test_inference_path_launches(n):
    """encode -> decode (convert.py path; cfg3 runs as 16,384-frame chunks plus a ragged last chunk): inference workspace
    layout, no backward kernels."""
    rec = HS.record_encode_decode(vcc2016_vae_arch(), n)
    check_recording(rec, n)
    assert not any(k in L["name"] for L in rec.launches for k in ("wgrad", "ln_bwd", "unpack", "recon", "adam"))
    if n == 40000:
        assert sorted({L["umma"]["frames"] for L in rec.launches if L["umma"]}) == [40000 - 2 * 16384, 16384]


def _check_operand_rows_stay_in_their_plane(rec, chunk):
    """The rule learned from the NaN * 0 bug: an A / dC operand box may never read bits from outside its own
    frame's plane, not even to multiply them by a zero weight.  For every 4-D operand map: the bytes reachable inside
    one frame -- (k extent) + (rows - 1) * row stride + (row groups - 1) * group stride -- measured from where the map
    starts inside its bf16 plane, must end inside that plane.  (Columns beyond dims[0] are TMA zero fill and never read;
    accumulator rows the epilogue discards may hold anything: rows are independent.)"""
    bufs = HS.buffer_ranges(rec.plan, chunk, rec.train)
    checked = 0
    for L in rec.launches:
        if not L["umma"]:
            continue
        for i in L["tmap"]:
            t = rec.tmaps[i]
            if t["rank"] != 4:
                continue
            pos = (t["base"] - rec.ws_lo) // 4                      # float index of the map's first element (hi or lo plane)
            hit = [b for b in bufs if b[1] <= pos < b[2]]
            assert len(hit) == 1 and hit[0][4], (L["name"], "operand map outside the split buffers", pos)
            name, lo, hi, per_frame, _ = hit[0]
            plane = per_frame * 2                                       # bytes of one bf16 plane of a frame
            start = (t["base"] - rec.ws_lo - lo * 4) % (per_frame * 4) % plane
            s_r, s_a, s_f = t["strides"]
            assert s_f == per_frame * 4 or t["dims"][3] == 1, (name, "frame stride")
            inner = t["dims"][0] * 2
            u = L["umma"]
            if u["tapT"] > 0 and "umma_fwd" in L["name"]:
                # tap mode: the last halo row of a row group is read by the valid output rows only up to the phase of the last
                # tap (T - 1 = P * m + pz); what lies behind it feeds the discarded halo accumulator rows alone
                inner = ((u["tapT"] - 1) % u["tapP"] + 1) * u["tapC"] * 2
            reach = inner + (t["dims"][1] - 1) * (s_r if t["dims"][1] > 1 else 0) + (t["dims"][2] - 1) * (s_a if t["dims"][2] > 1 else 0)
            assert start + reach <= plane, (L["name"], name, "operand rows leave their plane by %d bytes" % (start + reach - plane), t)
            checked += 1
    return checked


@pytest.mark.parametrize("switch", ["default", "pair_wide", "window_only", "wgrad_pair_256"])
def test_operand_boxes_never_leave_their_frame(switch):
    arch = vcc2016_vae_arch()
    for n in (8, 300, 16384):
        rec = HS.record_loss_fwd_bwd(arch, n, SWITCHES[switch])
        assert _check_operand_rows_stay_in_their_plane(rec, min(n, 16384)) > 40
    for name in sorted(ALT_ARCHS):
        rec = HS.record_loss_fwd_bwd(ALT_ARCHS[name], 29, SWITCHES[switch])
        _check_operand_rows_stay_in_their_plane(rec, 29)
    for seed in range(12):
        rec = HS.record_loss_fwd_bwd(_random_arch(np.random.RandomState(seed)), 50, SWITCHES[switch])
        _check_operand_rows_stay_in_their_plane(rec, 50)


def _check_epilogue_stores_stay_in_their_frame(rec, chunk):
    """Forward / dgrad epilogues write row j of a frame at in-frame element j * rs + off, N columns wide: inside the
    frame (or clipped to [0, flen) when the view is predicated), in a per-frame buffer of the workspace."""
    bufs = HS.buffer_ranges(rec.plan, chunk, rec.train)
    n = 0
    for L in rec.launches:
        u = L["umma"]
        if not u or "umma_fwd" not in L["name"]:
            continue
        pos = (u["c_ptr"] - rec.ws_lo) // 4
        hit = [b for b in bufs if b[1] <= pos < b[2] and b[0] != "da_shared"]
        if len(hit) > 1:                                 # tenants of the shared gradient buffer: the one with this view's frame stride
            hit = [b for b in hit if b[3] == u["c_fs"]][:1]
        assert len(hit) == 1 and pos == hit[0][1], (L["name"], "output view does not start at a workspace buffer", pos)
        name, lo, hi, per_frame, split = hit[0]
        assert u["c_fs"] == per_frame and bool(u["c_split"]) == bool(split), (name, u)
        for j in (0, u["c_R"] - 1):
            inf = j * u["c_rs"] + u["c_off"]
            if u["c_pred"]:
                assert 0 < u["c_flen"] <= per_frame, (name, u)
            else:
                assert 0 <= inf and inf + u["N"] <= per_frame, (name, "row %d writes [%d, %d) of a %d-element frame" % (j, inf, inf + u["N"], per_frame))
        n += 1
    return n


@pytest.mark.parametrize("switch", ["default", "pair_wide", "window_only"])
def test_epilogue_stores_stay_in_their_frame(switch):
    for n in (8, 16384):
        rec = HS.record_loss_fwd_bwd(vcc2016_vae_arch(), n, SWITCHES[switch])
        assert _check_epilogue_stores_stay_in_their_frame(rec, n) >= 15
    for name in sorted(ALT_ARCHS):
        _check_epilogue_stores_stay_in_their_frame(HS.record_loss_fwd_bwd(ALT_ARCHS[name], 29, SWITCHES[switch]), 29)
    for seed in range(12):
        _check_epilogue_stores_stay_in_their_frame(HS.record_loss_fwd_bwd(_random_arch(np.random.RandomState(seed)), 50, SWITCHES[switch]), 50)


def test_fused_kernels_and_small_batch_tiles_of_the_default_plan():
    """What the default engine launches for the reference architecture: the fused first-layer kernels (and no launch of the
    ops they replace), one speaker-backward kernel per call, the Layernorm epilogue for inference passes only, and -- at
    16 frames -- narrow N tiles that spread the dense generator layer over tens of CTAs."""
    arch = vcc2016_vae_arch()
    tr = HS.record_loss_fwd_bwd(arch, 16384)
    names = [L["name"] for L in tr.launches]
    assert sum("e0_fwd_kernel" in n for n in names) == 1 and sum("e0_bwd_kernel" in n for n in names) == 1
    assert not any("rowgemm_kernel" in n or "wgrad_tiny_kernel" in n for n in names)          # the unfused first-layer kernels
    assert sum("speaker_bwd_kernel" in n for n in names) == 1 and not any("colsum_kernel" in n for n in names)
    assert sum("zero_pads_kernel" in n for n in names) == 1
    assert not any(L["umma"] and L["umma"]["ln_on"] for L in tr.launches)                   # training keeps the Layernorm kernels
    assert sum("ln_fwd_reg_kernel" in n for n in names) == 7                                  # 8 Layernorms - the fused first layer
    bufs = {b["name"]: b for b in tr.plan["bufs"]}
    assert bufs["dc_e0"]["elide"] == 1 and all(bufs[k]["alias"] == [i for i, b in enumerate(tr.plan["bufs"]) if b["name"] == "da_shared"][0]
                                               for k in bufs if k.startswith("da_") and k != "da_shared")
    inf = HS.record_encode_decode(arch, 16384)
    fused = [L for L in inf.launches if L["umma"] and L["umma"]["ln_on"]]
    assert sorted((L["umma"]["K"], L["umma"]["N"]) for L in fused) == sorted([(112, 32), (224, 64), (448, 128), (264, 96), (96, 48)])
    assert sum("ln_fwd" in L["name"] for L in inf.launches) == 2                              # E4 (two N tiles) and G2 (frames span tiles)
    unf = HS.record_loss_fwd_bwd(arch, 300, {"NPVC_FUSE": "0"})
    assert any("rowgemm_kernel" in L["name"] for L in unf.launches) and not any("e0_" in L["name"] for L in unf.launches)
    small = HS.record_loss_fwd_bwd(arch, 16)
    g3 = [L for L in small.launches if L["umma"] and "umma_fwd" in L["name"] and (L["umma"]["K"], L["umma"]["N"]) == (4104, 513)]
    assert len(g3) == 1 and g3[0]["grid"][0] >= 30 and g3[0]["umma"]["BN"] == 16, g3


def test_gemm_output_rows_lie_on_32_byte_boundaries():
    """An epilogue thread owns one output row, so only whole 32-byte sectors per thread store efficiently (DESIGN.md §6:
    the merge layer and the last layer were 1.6x / 1.13x slower before their rows were aligned).  Every forward / dgrad
    launch of the reference architecture must put its rows on 32-byte boundaries of the fp32 buffer / of both bf16 planes;
    the one known exception (G0's dgrad: 88-element rows of a dense buffer another GEMM reads as K = 1672) is listed."""
    r = HS.record_loss_fwd_bwd(vcc2016_vae_arch(), 16384)
    bad = []
    for L in r.launches:
        u = L["umma"]
        if u is None or "umma_fwd_kernel_t" not in L["name"]:
            continue
        el = 2 if u["c_split"] else 4                                   # bytes per element of a plane / of the fp32 buffer
        rows_ok = (u["c_off"] * el) % 32 == 0 and (u["c_fs"] * el) % 32 == 0 and (u["c_R"] == 1 or (u["c_rs"] * el) % 32 == 0)
        if not (rows_ok and u["c_ptr"] % 32 == 0 and (u["BN"] * el) % 32 == 0):
            bad.append((u["K"], u["N"]))
    assert bad == [(288, 88)], bad
