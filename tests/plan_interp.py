"""numpy (float64) interpreter of the launch plan exported by ``npvc_plan_json``.

Executes exactly the op list, views and operand packs the CUDA engine executes, on the CPU, so
the host-side plan (index math, paddings, pack tables) is checked against the oracle without a
GPU.  Test infrastructure only.
"""
import numpy as np

SP_WS, SP_THETA, SP_GRAD, SP_AW, SP_ADW, SP_USER = 1, 2, 3, 4, 5, 6
U_X, U_Y, U_EPS = 0, 1, 2
(OP_GEMM, OP_WGRAD, OP_LN_FWD, OP_LN_BWD, OP_SAMPLE, OP_SAMPLE_BWD, OP_RECON, _UNUSED7, OP_COLSUM,
 OP_ZERO, OP_PACK, OP_UNPACK, OP_PACK16, OP_ZCAT) = range(14)
PH_PACK, PH_ENC, PH_SAMPLE, PH_DEC, PH_LOSS, PH_BWD, PH_FINAL = range(7)

LN_EPS = 1e-5
ONE_PLUS_EPS = float(np.float32(1.0) + np.float32(1e-6))
LOG_2PI = float(np.float32(1.8378770664093453))


class Interp:
    def __init__(self, plan, tables, theta, n, x=None, y=None, eps=None, n_total=None):
        self.plan, self.n = plan, n
        self.n_total = n_total or n
        self.theta = np.asarray(theta, np.float64)
        self.grad = np.zeros_like(self.theta)
        self.aw = np.zeros(plan["arena_w"], np.float64)
        self.aw16 = np.zeros(plan["aw16_count"], np.float64)     # bf16 operand packs (values held as float64)
        self.adw = np.zeros(plan["arena_dw"], np.float64)
        self.tables = tables
        self.x = None if x is None else np.asarray(x, np.float64).reshape(-1)
        self.y = None if y is None else np.asarray(y, np.int64)
        self.eps = None if eps is None else np.asarray(eps, np.float64)
        self.bufs = []
        for b in plan["bufs"]:
            self.bufs.append(np.full(b["fixed"] + b["per_frame"] * n, np.nan))   # NaN: catch reads of unwritten data
        self.buf_index = {b["name"]: i for i, b in enumerate(plan["bufs"])}

    # ---- references / views -------------------------------------------------------------
    def flat(self, ref):
        sp = ref["space"]
        if sp == SP_WS:
            return self.bufs[ref["buf"]], 0
        if sp == SP_THETA:
            return self.theta, ref["off"]
        if sp == SP_GRAD:
            return self.grad, ref["off"]
        if sp == SP_AW:
            return self.aw, ref["off"]
        if sp == SP_ADW:
            return self.adw, ref["off"]
        if sp == SP_USER:
            return {U_X: self.x, U_EPS: None if self.eps is None else self.eps.reshape(-1)}[ref["buf"]], 0
        return None, 0

    def buf(self, name):
        return self.bufs[self.buf_index[name]]

    # ---- bf16 hi / lo planes (plan.h Buf::split): a split buffer holds hi + lo of the fp32 value ----
    @staticmethod
    def _bf16(a):
        b = np.asarray(a, np.float32).view(np.uint32)
        r = (b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
        return r.view(np.float32).astype(np.float64)

    @classmethod
    def split(cls, a):
        a32 = np.asarray(a, np.float64).astype(np.float32).astype(np.float64)
        hi = cls._bf16(a32)
        return hi, cls._bf16(a32 - hi)

    @classmethod
    def q(cls, a):
        hi, lo = cls.split(a)
        return hi + lo

    def is_split(self, ref):
        return ref["space"] == SP_WS and bool(self.plan["bufs"][ref["buf"]].get("split"))

    def view_index(self, v, rows, width):
        r = np.arange(rows)
        f, j = r // v["R"], r % v["R"]
        inf = (j * v["rs"] + v["off"])[:, None] + np.arange(width)[None, :]
        idx = (f * v["fs"])[:, None] + inf
        valid = np.ones_like(idx, bool)
        if v["pred"]:
            valid = (inf >= 0) & (inf < v["flen"])
        return idx, valid, f

    def gather(self, v, rows, width):
        arr, base = self.flat(v["ref"])
        idx, valid, _ = self.view_index(v, rows, width)
        idx = np.where(valid, idx, 0) + base
        assert idx.min() >= 0 and idx.max() < arr.size, "view out of bounds"
        out = np.where(valid, arr[idx], 0.0)
        assert not np.isnan(out).any(), "read of unwritten workspace"
        return out

    def scatter(self, v, rows, vals):
        arr, base = self.flat(v["ref"])
        idx, valid, _ = self.view_index(v, rows, vals.shape[1])
        assert idx[valid].min() + base >= 0 and idx[valid].max() + base < arr.size
        if self.is_split(v["ref"]):
            vals = self.q(vals)
        arr[idx[valid] + base] = vals[valid]

    # ---- ops ----------------------------------------------------------------------------
    def run_phase(self, phase, with_grad=True):
        for op in self.plan["ops"]:
            if op["phase"] != phase:
                continue
            if op["kind"] == OP_UNPACK and not with_grad:
                continue
            getattr(self, "op_%d" % op["kind"])(op)

    def op_10(self, op):   # PACK (fp32 operand packs)
        src = self.tables["pack_src"]
        n = self.plan["aw16_off"]
        lst = self.tables.get("pack_list")
        if lst is not None and lst.size:                 # tensor path: only the positions some op reads as fp32 are packed
            self.aw[:n] = np.nan
            self.aw[lst] = np.where(src[lst] >= 0, self.theta[np.maximum(src[lst], 0)], 0.0)
        else:
            self.aw[:n] = np.where(src >= 0, self.theta[np.maximum(src, 0)], 0.0)

    def op_12(self, op):   # PACK16 (bf16 hi / lo operand packs of the tensor path)
        src = self.tables["pack16_src"]                  # entry i -> hi pack element i and lo pack element i + half
        half = self.plan["aw16_count"] // 2
        assert src.size == half
        idx = np.maximum(src, 0) & ((1 << 29) - 1)
        from_arena = (src & (1 << 29)) != 0
        val = np.where(from_arena, self.aw[np.minimum(idx, self.aw.size - 1)], self.theta[np.minimum(idx, self.theta.size - 1)])
        hi, lo = self.split(np.where(src >= 0, val, 0.0))
        self.aw16[:half] = np.where(src >= 0, hi, 0.0); self.aw16[half:] = np.where(src >= 0, lo, 0.0)

    def op_13(self, op):   # ZCAT: zs = [z | one-hot(y)]
        zd, yp, n = op["i0"], op["i1"], self.n
        z = self.flat(op["r0"])[0][:n * zd].reshape(n, zd)
        oh = np.zeros((n, yp)); oh[np.arange(n), self.y] = 1.0
        zs = np.concatenate([z, oh], 1)
        self.flat(op["r1"])[0][:n * (zd + yp)] = (self.q(zs) if self.is_split(op["r1"]) else zs).reshape(-1)

    def op_11(self, op):   # UNPACK
        ptr, idx = self.tables["unpack_ptr"], self.tables["unpack_idx"]
        cs = np.concatenate([[0.0], np.cumsum(self.adw[idx])])
        self.grad += cs[ptr[1:]] - cs[ptr[:-1]]

    def _rows(self, op):
        return op["rows_fixed"] or self.n * op["A"]["R"]

    def op_0(self, op):    # GEMM
        rows, K, N = self._rows(op), op["K"], op["N"]
        A = self.gather(op["A"], rows, K)
        if op.get("umma"):
            # tensor-core path, bf16x3: A planes x K-major [N, kpad] bf16 hi / lo packs, the lo.lo term dropped
            kp = op["kpad"]
            Bh = self.aw16[op["bu_hi"]:op["bu_hi"] + N * kp].reshape(N, kp)[:, :K].T
            Bl = self.aw16[op["bu_lo"]:op["bu_lo"] + N * kp].reshape(N, kp)[:, :K].T
            Ah, Al = self.split(A)
            Cv = Ah @ Bh + (Al @ Bh + Ah @ Bl)
        else:
            barr, boff = self.flat(op["B"])
            Bm = barr[boff:boff + K * op["ldb"]].reshape(K, op["ldb"])[:, :N]
            Cv = A @ Bm
        cols = np.arange(N) % op["bias_mod"]
        for key in ("bias0", "bias1", "bias2"):
            if op[key]["space"]:
                arr, off = self.flat(op[key])
                Cv = Cv + arr[off + cols][None, :]
        self.scatter(op["C"], rows, Cv)

    def op_1(self, op):    # WGRAD
        rows, K, N = self._rows(op), op["K"], op["N"]
        A = self.gather(op["A"], rows, K)
        D = self.gather(op["C"], rows, N)
        arr, off = self.flat(op["B"])
        out = arr[off:off + K * op["ldb"]].reshape(K, op["ldb"])
        if op.get("umma"):
            Ah, Al = self.split(A)
            Dh, Dl = self.split(D)
            out[:, :N] += Ah.T @ Dh + (Al.T @ Dh + Ah.T @ Dl)
        else:
            out[:, :N] += A.T @ D

    def op_2(self, op):    # LN_FWD
        L, Cn, n = op["L"], op["Cn"], self.n
        x = self.flat(op["in"])[0][:n * L].reshape(n, L)
        garr, goff = self.flat(op["gamma"]); barr, boff = self.flat(op["beta"])
        gamma, beta = garr[goff:goff + Cn], barr[boff:boff + Cn]
        m = x.mean(1, keepdims=True)
        v = ((x - m) ** 2).mean(1, keepdims=True)
        rs = 1.0 / np.sqrt(v + LN_EPS)
        xh = (x - m) * rs
        c = np.arange(L) % Cn
        u = xh * gamma[c] + beta[c]
        a = np.maximum(u, 0.02 * u)
        self.flat(op["r0"])[0][:n] = m[:, 0]            # (mean, rstd) kept; backward recomputes xhat from c
        self.flat(op["rstd"])[0][:n] = rs[:, 0]
        out = self.flat(op["aout"])[0].reshape(n, op["out_flen"])
        out[:] = 0.0
        out[:, op["out_off"]:op["out_off"] + L] = self.q(a) if self.is_split(op["aout"]) else a

    def op_3(self, op):    # LN_BWD
        L, Cn, n = op["L"], op["Cn"], self.n
        dy = self.flat(op["in"])[0][:n * L].reshape(n, L)
        rs = self.flat(op["rstd"])[0][:n][:, None]
        xh = (self.flat(op["xhat"])[0][:n * L].reshape(n, L) - self.flat(op["r0"])[0][:n][:, None]) * rs   # "xhat" ref = raw conv output
        garr, goff = self.flat(op["gamma"]); barr, boff = self.flat(op["beta"])
        gamma, beta = garr[goff:goff + Cn], barr[boff:boff + Cn]
        assert not np.isnan(dy).any() and not np.isnan(xh).any()
        c = np.arange(L) % Cn
        u = xh * gamma[c] + beta[c]
        du = dy * np.where(u >= 0, 1.0, 0.02)
        dxh = du * gamma[c]
        s1 = dxh.mean(1, keepdims=True)
        s2 = (dxh * xh).mean(1, keepdims=True)
        dc = rs * (dxh - s1 - xh * s2)
        out = self.flat(op["aout"])[0].reshape(n, op["out_flen"])
        out[:] = 0.0
        out[:, op["out_off"]:op["out_off"] + L] = self.q(dc) if self.is_split(op["aout"]) else dc
        for key, val in (("dgamma", du * xh), ("dbeta", du), ("dbias", dc)):
            arr, off = self.flat(op[key])
            arr[off:off + Cn] += val.reshape(n, L // Cn, Cn).sum((0, 1))

    def op_4(self, op):    # SAMPLE (+KL)
        z, n = op["i0"], self.n
        hz = self.flat(op["r0"])[0][:n * 2 * z].reshape(n, 2 * z)
        mu, lv = hz[:, :z], hz[:, z:]
        self.flat(op["r1"])[0][:n * z] = mu.reshape(-1)
        self.flat(op["r2"])[0][:n * z] = lv.reshape(-1)
        if self.eps is not None:
            self.flat(op["r3"])[0][:n * z] = (mu + self.eps * np.sqrt(np.exp(lv))).reshape(-1)
            acc = self.buf("acc")
            acc[0] = np.nan_to_num(acc[0]) + (0.5 * (-lv + (np.exp(lv) + mu * mu) / ONE_PLUS_EPS - 1.0)).sum()

    def op_5(self, op):    # SAMPLE_BWD
        z, n = op["i0"], self.n
        dz = self.flat(op["r0"])[0][:n * z].reshape(n, z)
        hz = self.flat(op["r1"])[0][:n * 2 * z].reshape(n, 2 * z)
        mu, lv = hz[:, :z], hz[:, z:]
        inv_n = 1.0 / self.n_total
        dmu = dz + mu / ONE_PLUS_EPS * inv_n
        dlv = dz * self.eps * 0.5 * np.sqrt(np.exp(lv)) + 0.5 * (np.exp(lv) / ONE_PLUS_EPS - 1.0) * inv_n
        dhz = np.concatenate([dmu, dlv], 1)
        self.flat(op["r2"])[0][:n * 2 * z] = (self.q(dhz) if self.is_split(op["r2"]) else dhz).reshape(-1)
        arr, off = self.flat(op["r3"])
        arr[off:off + 2 * z] += dhz.sum(0)

    def op_6(self, op):    # RECON
        H, ld, n = op["i0"], op["i1"], self.n
        x = self.x.reshape(n, H)
        xh = self.flat(op["r1"])[0][:n * ld].reshape(n, ld)[:, :H]     # (rows at the pitch of dxh; pads never written)
        d = xh - x
        acc = self.buf("acc")
        acc[1] = np.nan_to_num(acc[1]) + (-0.5 * (LOG_2PI + d * d / ONE_PLUS_EPS)).sum()
        g = d / ONE_PLUS_EPS / self.n_total
        dxh = self.flat(op["r2"])[0].reshape(n, ld)
        dxh[:] = 0.0
        dxh[:, :H] = self.q(g) if self.is_split(op["r2"]) else g
        arr, off = self.flat(op["r3"])
        arr[off] += g.sum()

    def op_8(self, op):    # COLSUM
        N, rows = op["i0"], op["i1"]
        arr0, off0 = self.flat(op["r0"])
        src = arr0[off0:off0 + rows * N].reshape(rows, N)
        arr, off = self.flat(op["r1"])
        arr[off:off + N] += src.sum(0)

    def op_9(self, op):    # ZERO
        cnt = op["count"] * (self.n if op["per_frame_count"] else 1)
        self.flat(op["r0"])[0][:cnt] = 0.0

    # ---- whole passes -------------------------------------------------------------------
    def loss_fwd_bwd(self, with_grad=True):
        self.buf("acc")[:] = 0.0
        for ph in (PH_PACK, PH_ENC, PH_SAMPLE, PH_DEC, PH_LOSS):
            self.run_phase(ph)
        if with_grad:
            self.run_phase(PH_BWD)
            self.run_phase(PH_FINAL)
        z = self.plan["z"]
        acc = self.buf("acc")
        kl, lp = acc[0] / self.n_total, acc[1] / self.n_total
        return {"mu": self.buf("mu").reshape(self.n, z), "lv": self.buf("lv").reshape(self.n, z),
                "z": self.buf("z").reshape(self.n, z), "xh": self.buf("xh").reshape(self.n, -1)[:, :self.plan["out_dim"]],
                "D_KL": kl, "logP": lp, "G": -lp + kl, "grad": self.grad}
