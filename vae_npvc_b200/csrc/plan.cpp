// Plan builder: architecture -> parameter table, buffers, operand packs, op list.
// Reference semantics restated here are cited per block (paths relative to the reference repo).
#include "plan.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace npvc {

namespace {

inline int rup(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t rup64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }   // floor
inline int cdivs(int a, int b) { return -fdiv(-a, b); }                             // ceil, signed

struct EncL { int Ci, Co, k, s, Hi, Ho, pl, pr; };
struct GenL { int Ci, Cip, Co, k, s, Hi, Ho, cl, dense, lo, hi, wn; };

// dgrad of a transposed conv with 8 output channels: rows split by parity into two GEMMs (see the op list)
inline bool gen_parity_split(int Co, int s, int Hi, bool use_umma) { return use_umma && Co == 8 && (s * Co) % 8 == 0 && Hi >= 2; }

struct Builder {
  Plan& p;
  explicit Builder(Plan& pl) : p(pl) {}

  int add_param(const std::string& name, std::vector<int> shape, int fi, int fo, int init) {
    Param q; q.name = name; q.off = p.n_params; q.rank = (int)shape.size(); q.size = 1;
    for (size_t i = 0; i < shape.size(); i++) { q.shape[i] = shape[i]; q.size *= shape[i]; }
    q.fan_in = fi; q.fan_out = fo; q.init = init;
    p.n_params += q.size; p.params.push_back(q);
    return (int)p.params.size() - 1;
  }
  int add_buf(const std::string& name, int64_t per_frame, int64_t fixed, bool train_only, bool split = false) {
    Buf b; b.name = name; b.per_frame = per_frame; b.fixed = fixed; b.train_only = train_only; b.split = split;
    p.bufs.push_back(b); return (int)p.bufs.size() - 1;
  }
  // arena allocation (128-byte aligned); fills pack_src with -1
  int64_t aw_alloc(int64_t n) {
    int64_t off = p.arena_w; p.arena_w = rup64(off + n, 32);
    p.pack_src.resize(p.arena_w, -1); return off;
  }
  static Ref ws(int b) { Ref r; r.space = SP_WS; r.buf = b; return r; }
  static Ref th(int64_t off) { Ref r; r.space = SP_THETA; r.off = off; return r; }
  static Ref gr(int64_t off) { Ref r; r.space = SP_GRAD; r.off = off; return r; }
  static Ref aw(int64_t off) { Ref r; r.space = SP_AW; r.off = off; return r; }
  static Ref adw(int64_t off) { Ref r; r.space = SP_ADW; r.off = off; return r; }
  static Ref user(int slot) { Ref r; r.space = SP_USER; r.buf = slot; return r; }
  View view(Ref ref, int R, int64_t fs, int rs, int off, int flen, int pred = 0) const {
    View v; v.ref = ref; v.R = R; v.fs = fs; v.rs = rs; v.off = off; v.flen = flen; v.pred = pred;
    if (ref.space == SP_WS) v.split = p.bufs[ref.buf].split;     // split buffers: fs == per_frame (whole frames)
    return v;
  }
  Op& op(int kind, int phase, const std::string& name) {
    Op o; o.kind = kind; o.phase = phase; o.name = name; p.ops.push_back(o); return p.ops.back();
  }
};

void json_ref(std::ostringstream& o, const char* key, const Ref& r) {
  o << "\"" << key << "\":{\"space\":" << r.space << ",\"buf\":" << r.buf << ",\"off\":" << r.off << "}";
}
void json_view(std::ostringstream& o, const char* key, const View& v) {
  o << "\"" << key << "\":{";
  json_ref(o, "ref", v.ref);
  o << ",\"R\":" << v.R << ",\"fs\":" << v.fs << ",\"rs\":" << v.rs << ",\"off\":" << v.off
    << ",\"flen\":" << v.flen << ",\"pred\":" << v.pred << ",\"split\":" << v.split << "}";
}

}  // namespace

// Rows of one frame that go into one <= 128-row tensor-core tile: R itself when it fits (the tile then
// holds floor(128/R) whole frames), else the largest divisor of R in [16, 128] (a frame spans R/Rb tiles).
int umma_row_tile(int R) {
  if (R <= 128) return R;
  for (int d = 128; d >= 16; d--) if (R % d == 0) return d;
  return 0;
}

// a thread's 8 elements = 8 consecutive channels of one position, at most 4 such units per thread
int ln_group(int L, int Cn, int out_off, int out_flen) {
  if (L % 8 || out_off % 8 || out_flen % 8 || Cn % 8 || Cn > 2048) return 0;
  for (int G = 32; G <= 256; G *= 2)
    if (L <= 32 * G && (8 * G) % Cn == 0) return G;
  return 0;
}
int e0_bwd_group(int L, int Co) {
  if (L % 4 || Co % 4 || Co > 1024) return 0;
  for (int G = 32; G <= 256; G *= 2)
    if (L <= 16 * G && (4 * G) % Co == 0) return G;
  return 0;
}

int64_t Plan::buf_offset(int b, int64_t chunk, bool train) const {
  int64_t off = rup64(arena_w, 64);
  for (int i = 0; i < (int)bufs.size(); i++) {
    const Buf& q = bufs[i];
    int64_t sz = ((q.train_only && !train) || q.elide || q.alias >= 0) ? 0 : rup64(q.fixed + q.per_frame * chunk, 64);
    if (i == b) return q.alias >= 0 ? buf_offset(q.alias, chunk, train) : off;
    off += sz;
  }
  return off;
}
int64_t Plan::ws_floats(int64_t chunk, bool train) const { return buf_offset((int)bufs.size(), chunk, train); }

std::string build_plan(const npvc_arch& a, Plan& p, bool use_umma, bool fuse) {
  p = Plan();
  p.arch = a;
  Builder B(p);
  char msg[256];
  // tensor path: activations that feed GEMMs are bf16 hi / lo planes -> 16-byte alignment is 8 elements.
  // AL pads the channel counts this builder is free to pad; layers whose own channel counts are only
  // multiples of 4 stay on the CUDA-core kernels (which read the planes with 8-byte loads).
  const int AL = use_umma ? 8 : 4;
  const bool SPLIT = use_umma;

  // ---------------------------------------------------------------- validate (model/vae.py:37-39)
  if (a.n_enc < 1 || a.n_enc > NPVC_MAX_LAYERS || a.n_gen < 1 || a.n_gen > NPVC_MAX_LAYERS)
    return "encoder/generator must have 1..8 layers";
  if (a.in_h < 1 || a.z_dim < 4 || a.z_dim % 4 || a.y_dim < 1) return "bad in_h / z_dim (multiple of 4) / y_dim";
  if (a.n_gen < 2) return "generator needs at least one transposed conv before the dense layer";
  if (2 * a.z_dim > 1024) return "z_dim > 512 is not supported";
  if (a.gen_out[a.n_gen - 1] != 1) return "the last generator layer must have 1 output channel";

  // ---------------------------------------------------------------- geometry
  // tf.layers.conv2d(padding='same'): H_out = ceil(H/s), pad_total = max((H_out-1)s + k - H, 0)
  std::vector<EncL> E; {
    int H = a.in_h, ci = 1;
    for (int i = 0; i < a.n_enc; i++) {
      EncL l; l.Ci = ci; l.Co = a.enc_out[i]; l.k = a.enc_kernel[i]; l.s = a.enc_stride[i]; l.Hi = H;
      if (l.k < 1 || l.s < 1 || l.Co < 4 || l.Co % 4) return "encoder.output must be multiples of 4; kernel,stride >= 1";
      if (l.k < l.s) return "encoder kernel < stride is not supported";
      l.Ho = cdiv(H, l.s);
      int pt = std::max((l.Ho - 1) * l.s + l.k - H, 0);
      l.pl = pt / 2; l.pr = pt - l.pl;
      E.push_back(l); H = l.Ho; ci = l.Co;
    }
  }
  // tf.layers.conv2d_transpose(padding='same'): H_out = s*H, crop_left = (k-s)//2
  std::vector<GenL> G; {
    int H = a.gen_h, ci = a.gen_c;
    for (int i = 0; i < a.n_gen; i++) {
      GenL l; l.Ci = ci; l.Cip = (i == 0) ? rup(ci, AL) : ci; /* only gen_c is ours to pad */ l.Co = a.gen_out[i]; l.k = a.gen_kernel[i]; l.s = a.gen_stride[i];
      l.Hi = H; l.Ho = H * l.s;
      if (l.k < l.s || l.s < 1) return "generator kernel < stride is not supported";
      l.cl = std::max(l.k - l.s, 0) / 2;
      bool last = (i == a.n_gen - 1);
      l.dense = (l.s == 1 && l.k > 3);     // wide stride-1 layer -> dense Toeplitz GEMM
      if (!last && (l.Co % 4)) return "generator.output (all but last) must be multiples of 4";
      if (l.dense && !last) return "a stride-1 wide transposed conv is only supported as the last generator layer";
      if (!l.dense && last) return "the last generator layer must be stride-1 (dense) in this build";
      // window of input positions i in [j+lo, j+hi] feeding output positions s*j + r, r in [0,s)
      l.hi = fdiv(l.s - 1 + l.cl, l.s);
      l.lo = cdivs(l.cl - l.k + 1, l.s);
      l.wn = l.hi - l.lo + 1;
      G.push_back(l); H = l.Ho; ci = l.Co;
    }
    if (H * G.back().Co != a.in_h) {
      snprintf(msg, sizeof msg, "generator output (%d) != in_h (%d)", H * G.back().Co, a.in_h);
      return msg;
    }
  }
  const int z = a.z_dim, nE = a.n_enc, nG = a.n_gen;
  const int gh = a.gen_h, gc = a.gen_c, gcp = rup(gc, AL), Nm = gh * gcp;
  const int flat = E.back().Ho * E.back().Co;
  p.out_dim = a.in_h;
  const int xhld = rup(a.in_h, AL);

  // ---------------------------------------------------------------- parameter table
  // tf.trainable_variables() creation order: model/vae.py:20-24 (y_emb), then loss() ->
  // _encode (util/layers.py:55-65: conv kernel, bias, layernorm.offset, layernorm.scale;
  // model/vae.py:80-81 dense, dense_1), then _generate (model/vae.py:51-61 two slim FCs +
  // bias_add; :96-101 conv2d_transpose kernel, bias, ConvT-LN offset, scale).
  int P_emb = B.add_param("y_embedding/y_emb", {a.y_dim, z}, a.y_dim, z, 0);
  std::vector<int> P_ek(nE), P_eb(nE), P_eo(nE), P_es(nE);
  for (int i = 0; i < nE; i++) {
    char nm[96];
    snprintf(nm, sizeof nm, "Encoder/Conv2d-%d/Conv2d-%d/kernel", i, i);
    P_ek[i] = B.add_param(nm, {E[i].k, 1, E[i].Ci, E[i].Co}, E[i].k * E[i].Ci, E[i].k * E[i].Co, 0);
    snprintf(nm, sizeof nm, "Encoder/Conv2d-%d/Conv2d-%d/bias", i, i);
    P_eb[i] = B.add_param(nm, {E[i].Co}, 0, 0, 1);
    snprintf(nm, sizeof nm, "Encoder/Conv2d-%d/layernorm.offset", i);
    P_eo[i] = B.add_param(nm, {E[i].Co, 1, 1}, 0, 0, 1);
    snprintf(nm, sizeof nm, "Encoder/Conv2d-%d/layernorm.scale", i);
    P_es[i] = B.add_param(nm, {E[i].Co, 1, 1}, 0, 0, 2);
  }
  int P_dk[2], P_db[2];
  P_dk[0] = B.add_param("Encoder/dense/kernel", {flat, z}, flat, z, 0);
  P_db[0] = B.add_param("Encoder/dense/bias", {z}, 0, 0, 1);
  P_dk[1] = B.add_param("Encoder/dense_1/kernel", {flat, z}, flat, z, 0);
  P_db[1] = B.add_param("Encoder/dense_1/bias", {z}, 0, 0, 1);
  int P_fw[2], P_fb[3];
  P_fw[0] = B.add_param("Generator/fully_connected/weights", {z, gh * gc}, z, gh * gc, 0);
  P_fb[0] = B.add_param("Generator/fully_connected/biases", {gh * gc}, 0, 0, 1);
  P_fw[1] = B.add_param("Generator/fully_connected_1/weights", {z, gh * gc}, z, gh * gc, 0);
  P_fb[1] = B.add_param("Generator/fully_connected_1/biases", {gh * gc}, 0, 0, 1);
  P_fb[2] = B.add_param("Generator/BiasAdd/biases", {gh * gc}, 0, 0, 1);
  std::vector<int> P_gk(nG), P_gb(nG), P_go(nG, -1), P_gs(nG, -1);
  for (int i = 0; i < nG; i++) {
    char nm[96], base[64];
    if (i == 0) snprintf(base, sizeof base, "Generator/conv2d_transpose");
    else snprintf(base, sizeof base, "Generator/conv2d_transpose_%d", i);
    snprintf(nm, sizeof nm, "%s/kernel", base);
    P_gk[i] = B.add_param(nm, {G[i].k, 1, G[i].Co, G[i].Ci}, G[i].k * G[i].Co, G[i].k * G[i].Ci, 0);
    snprintf(nm, sizeof nm, "%s/bias", base);
    P_gb[i] = B.add_param(nm, {G[i].Co}, 0, 0, 1);
    if (i < nG - 1) {
      snprintf(nm, sizeof nm, "Generator/ConvT-LN%d.offset", i);
      P_go[i] = B.add_param(nm, {G[i].Co, 1, 1}, 0, 0, 1);
      snprintf(nm, sizeof nm, "Generator/ConvT-LN%d.scale", i);
      P_gs[i] = B.add_param(nm, {G[i].Co, 1, 1}, 0, 0, 2);
    }
  }
  auto poff = [&](int pi) { return p.params[pi].off; };
  if (p.n_params >= (int64_t)1 << 31) return "too many parameters";

  // ---------------------------------------------------------------- operand packs (arena_w)
  // forward part (mirrored by arena_dw for the weight gradients)
  std::vector<int64_t> A_we(nE), A_gf(nG);
  std::vector<int> ld_gf(nG);
  for (int e = 0; e < nE; e++) {               // [k*Ci, Co] == TF HWIO layout (identity copy)
    int64_t n = (int64_t)E[e].k * E[e].Ci * E[e].Co;
    A_we[e] = B.aw_alloc(n);
    for (int64_t i = 0; i < n; i++) p.pack_src[A_we[e] + i] = (int32_t)(poff(P_ek[e]) + i);
  }
  // heads: BH[(h,c), (head,d)] = W_head[c*Ho + h, d]   (slim.flatten of NCHW, model/vae.py:79)
  const int Ho4 = E.back().Ho, Co4 = E.back().Co;
  int64_t A_bh = B.aw_alloc((int64_t)flat * 2 * z);
  for (int h = 0; h < Ho4; h++) for (int c = 0; c < Co4; c++) for (int hd = 0; hd < 2; hd++) for (int d = 0; d < z; d++)
    p.pack_src[A_bh + ((int64_t)(h * Co4 + c)) * 2 * z + hd * z + d] = (int32_t)(poff(P_dk[hd]) + (int64_t)(c * Ho4 + h) * z + d);
  int64_t A_bhb = B.aw_alloc(2 * z);
  for (int hd = 0; hd < 2; hd++) for (int d = 0; d < z; d++) p.pack_src[A_bhb + hd * z + d] = (int32_t)(poff(P_db[hd]) + d);
  // merge: BZ/BY[d, (h,c')] = W[d, c*gh + h] (reshape to [-1, c, h, w], model/vae.py:94), c' padded to 4
  // The z-branch operand carries ypad extra rows: row z + s holds the per-speaker table P[s,:] (written by the
  // "ptab" GEMM at pack time), multiplied by one-hot(y) columns appended to z -- the speaker term rides
  // in the merge GEMM itself, and its weight-gradient rows ARE the per-speaker sums of the merge gradient.
  const int ypad = rup(a.y_dim, AL), zk = z + ypad;
  int64_t A_bz[2], A_bm[3];
  for (int w = 0; w < 2; w++) {
    A_bz[w] = B.aw_alloc((int64_t)(w == 0 ? zk : z) * Nm);
    for (int d = 0; d < z; d++) for (int h = 0; h < gh; h++) for (int c = 0; c < gc; c++)
      p.pack_src[A_bz[w] + (int64_t)d * Nm + h * gcp + c] = (int32_t)(poff(P_fw[w]) + (int64_t)d * gh * gc + c * gh + h);
  }
  for (int w = 0; w < 3; w++) {
    A_bm[w] = B.aw_alloc(Nm);
    for (int h = 0; h < gh; h++) for (int c = 0; c < gc; c++)
      p.pack_src[A_bm[w] + h * gcp + c] = (int32_t)(poff(P_fb[w]) + c * gh + h);
  }
  for (int g = 0; g < nG; g++) {
    const GenL& l = G[g];
    if (!l.dense) {
      // GF[(pos,c), (r,o)] = W[kk,0,o,c], kk = r + cl - s*(lo+pos): output position s*j + r gets
      // input position j+lo+pos through tap kk (out[o, s*i + kk - cl] += x[c,i] W[kk,0,o,c])
      int K = l.wn * l.Cip, N = l.s * l.Co; ld_gf[g] = N;
      A_gf[g] = B.aw_alloc((int64_t)K * N);
      for (int pos = 0; pos < l.wn; pos++) for (int c = 0; c < l.Ci; c++) for (int r = 0; r < l.s; r++) {
        int kk = r + l.cl - l.s * (l.lo + pos);
        if (kk < 0 || kk >= l.k) continue;
        for (int o = 0; o < l.Co; o++)
          p.pack_src[A_gf[g] + (int64_t)(pos * l.Cip + c) * N + r * l.Co + o] =
              (int32_t)(poff(P_gk[g]) + ((int64_t)kk * l.Co + o) * l.Ci + c);
      }
    } else {
      // dense Toeplitz T[(i,c), (pp,o)] = W[pp - i + cl, 0, o, c]
      int K = l.Hi * l.Ci, N = l.Ho * l.Co; ld_gf[g] = rup(N, 4);
      A_gf[g] = B.aw_alloc((int64_t)K * ld_gf[g]);
      for (int i = 0; i < l.Hi; i++) for (int c = 0; c < l.Ci; c++) for (int pp = 0; pp < l.Ho; pp++) {
        int kk = pp - i + l.cl;
        if (kk < 0 || kk >= l.k) continue;
        for (int o = 0; o < l.Co; o++)
          p.pack_src[A_gf[g] + (int64_t)(i * l.Ci + c) * ld_gf[g] + pp * l.Co + o] =
              (int32_t)(poff(P_gk[g]) + ((int64_t)kk * l.Co + o) * l.Ci + c);
      }
    }
  }
  const int64_t fwd_size = p.arena_w;
  p.arena_dw = fwd_size;
  // unpack CSR over the forward part (b2/b3 share b1's gradient vector)
  {
    std::vector<int32_t> cnt(p.n_params + 1, 0);
    auto dwpos = [&](int64_t q) -> int64_t {
      for (int w = 1; w < 3; w++) if (q >= A_bm[w] && q < A_bm[w] + Nm) return q - A_bm[w] + A_bm[0];
      return q;
    };
    for (int64_t q = 0; q < fwd_size; q++) if (p.pack_src[q] >= 0) cnt[p.pack_src[q] + 1]++;
    p.unpack_ptr.assign(p.n_params + 1, 0);
    for (int64_t t = 0; t < p.n_params; t++) p.unpack_ptr[t + 1] = p.unpack_ptr[t] + cnt[t + 1];
    p.unpack_idx.assign(p.unpack_ptr[p.n_params], 0);
    std::vector<int32_t> cur(p.unpack_ptr.begin(), p.unpack_ptr.end() - 1);
    for (int64_t q = 0; q < fwd_size; q++) if (p.pack_src[q] >= 0) p.unpack_idx[cur[p.pack_src[q]]++] = (int32_t)dwpos(q);
  }
  // backward-only packs
  std::vector<int64_t> A_gd(nG), A_ed(nE, -1);
  std::vector<int> ld_gd(nG);
  for (int g = 0; g < nG; g++) {
    const GenL& l = G[g];
    if (l.dense) {   // TD[(pp,o), (i,c)]
      int K = l.Ho * l.Co, N = l.Hi * l.Ci; ld_gd[g] = rup(N, 4);
      A_gd[g] = B.aw_alloc((int64_t)K * ld_gd[g]);
      for (int pp = 0; pp < l.Ho; pp++) for (int o = 0; o < l.Co; o++) for (int i = 0; i < l.Hi; i++) {
        int kk = pp - i + l.cl;
        if (kk < 0 || kk >= l.k) continue;
        for (int c = 0; c < l.Ci; c++)
          p.pack_src[A_gd[g] + (int64_t)(pp * l.Co + o) * ld_gd[g] + i * l.Ci + c] =
              (int32_t)(poff(P_gk[g]) + ((int64_t)kk * l.Co + o) * l.Ci + c);
      }
    } else {         // GD[(kk,o), c] = W[kk,0,o,c]: dgrad of a transposed conv is a strided conv
      ld_gd[g] = l.Cip;
      A_gd[g] = B.aw_alloc((int64_t)l.k * l.Co * l.Cip);
      for (int kk = 0; kk < l.k; kk++) for (int o = 0; o < l.Co; o++) for (int c = 0; c < l.Ci; c++)
        p.pack_src[A_gd[g] + (int64_t)(kk * l.Co + o) * l.Cip + c] = (int32_t)(poff(P_gk[g]) + ((int64_t)kk * l.Co + o) * l.Ci + c);
    }
  }
  int64_t A_bzd[2];
  for (int w = 0; w < 2; w++) {   // BZD/BYD[(h,c'), d]
    A_bzd[w] = B.aw_alloc((int64_t)Nm * z);
    for (int h = 0; h < gh; h++) for (int c = 0; c < gc; c++) for (int d = 0; d < z; d++)
      p.pack_src[A_bzd[w] + (int64_t)(h * gcp + c) * z + d] = (int32_t)(poff(P_fw[w]) + (int64_t)d * gh * gc + c * gh + h);
  }
  int64_t A_hd = B.aw_alloc((int64_t)2 * z * flat);   // HD[(head,d), (h,c)]
  for (int hd = 0; hd < 2; hd++) for (int d = 0; d < z; d++) for (int h = 0; h < Ho4; h++) for (int c = 0; c < Co4; c++)
    p.pack_src[A_hd + (int64_t)(hd * z + d) * flat + h * Co4 + c] = (int32_t)(poff(P_dk[hd]) + (int64_t)(c * Ho4 + h) * z + d);
  // encoder dgrad (transposed form): rows (frame, q), window of wn = ceil(k/s) output positions
  // j = q-(wn-1)+m; ED[(m,o), (r,c)] = W[kk,0,c,o], kk = r + s*(wn-1-m): padded input position
  // s*q + r receives dc[o, j] through tap kk = s*q + r - s*j.
  std::vector<int> e_wn(nE), e_q0(nE), e_q1(nE), e_pf(nE, 0), e_pb(nE, 0);
  for (int e = 1; e < nE; e++) {
    const EncL& l = E[e];
    e_wn[e] = cdiv(l.k, l.s); e_q0[e] = l.pl / l.s; e_q1[e] = (l.pl + l.Hi - 1) / l.s;
    e_pf[e] = std::max(0, (e_wn[e] - 1) - e_q0[e]); e_pb[e] = std::max(0, e_q1[e] - (l.Ho - 1));
    int K = e_wn[e] * l.Co, N = l.s * l.Ci;
    A_ed[e] = B.aw_alloc((int64_t)K * N);
    for (int m = 0; m < e_wn[e]; m++) for (int r = 0; r < l.s; r++) {
      int kk = r + l.s * (e_wn[e] - 1 - m);
      if (kk >= l.k) continue;
      for (int o = 0; o < l.Co; o++) for (int c = 0; c < l.Ci; c++)
        p.pack_src[A_ed[e] + (int64_t)(m * l.Co + o) * N + r * l.Ci + c] = (int32_t)(poff(P_ek[e]) + ((int64_t)kk * l.Ci + c) * l.Co + o);
    }
  }

  // ---------------------------------------------------------------- buffers
  p.buf_acc = B.add_buf("acc", 0, 8, false);                    // 2 doubles: sum KL, sum logP (+pad)
  // acc comes first so that theta-derived state (arena_w) sits at the same offsets in
  // the inference and training layouts
  p.buf_adw = B.add_buf("arena_dw", 0, p.arena_dw, true);
  // gradients w.r.t. the activations (fp32): da_l is written by ONE dgrad GEMM and read by the Layernorm backward that
  // follows it on the caller's stream, and the next da is written only after that -- all of them share one buffer
  // (sized for the largest: 48 KB per frame less workspace for the reference architecture)
  int64_t da_max = 0;
  for (int e = 0; e < nE; e++) da_max = std::max<int64_t>(da_max, (int64_t)E[e].Ho * E[e].Co);
  for (int g = 0; g + 1 < nG; g++) da_max = std::max<int64_t>(da_max, (int64_t)G[g].Ho * G[g].Co);
  const int b_da = B.add_buf("da_shared", da_max, 0, true);
  std::vector<int> b_ce(nE), b_me(nE), b_ae(nE), b_re(nE), b_dce(nE), b_dae(nE);
  std::vector<int> ae_flen(nE), ae_off(nE), dce_flen(nE), dce_off(nE);
  for (int e = 0; e < nE; e++) {
    char nm[32]; const EncL& l = E[e]; int L = l.Ho * l.Co;
    snprintf(nm, sizeof nm, "c_e%d", e);    b_ce[e] = B.add_buf(nm, L, 0, false);
    snprintf(nm, sizeof nm, "mean_e%d", e); b_me[e] = B.add_buf(nm, 1, 0, false);
    if (e < nE - 1) { ae_flen[e] = (E[e + 1].pl + l.Ho + E[e + 1].pr) * l.Co; ae_off[e] = E[e + 1].pl * l.Co; }
    else { ae_flen[e] = L; ae_off[e] = 0; }
    snprintf(nm, sizeof nm, "a_e%d", e);    b_ae[e] = B.add_buf(nm, ae_flen[e], 0, false, SPLIT);
    snprintf(nm, sizeof nm, "rstd_e%d", e); b_re[e] = B.add_buf(nm, 1, 0, false);
    dce_flen[e] = (e_pf[e] + l.Ho + e_pb[e]) * l.Co; dce_off[e] = e_pf[e] * l.Co;
    snprintf(nm, sizeof nm, "dc_e%d", e);   b_dce[e] = B.add_buf(nm, dce_flen[e], 0, true, SPLIT);
    snprintf(nm, sizeof nm, "da_e%d", e);   b_dae[e] = B.add_buf(nm, L, 0, true); p.bufs[b_dae[e]].alias = b_da;
  }
  p.buf_hz = B.add_buf("hz", 2 * z, 0, false);
  p.buf_mu = B.add_buf("mu", z, 0, false);
  p.buf_lv = B.add_buf("lv", z, 0, false);
  p.buf_z = B.add_buf("z", z, 0, false);
  // zs: the sampled z (or the caller's z in decode()) as operand planes for the merge GEMM
  int b_zs = B.add_buf("zs", zk, 0, false, SPLIT);
  int b_dz = B.add_buf("dz", z, 0, true), b_dhz = B.add_buf("dhz", 2 * z, 0, true, SPLIT);
  // The merge GEMM writes the interior of the padded frame: a lead of < 16 elements in front of the frame puts the interior
  // (and, with a frame length that is a multiple of 16, every row of the output) on 32-byte boundaries of both planes, which
  // is what the epilogue's 32-byte stores need (88 elements of front padding alone left it on 16-byte stores).
  const int hm_lead = (16 - (-G[0].lo * gcp) % 16) % 16;
  const int hm_flen = ((-G[0].lo + gh + G[0].hi) * gcp + hm_lead + 15) / 16 * 16, hm_off = -G[0].lo * gcp + hm_lead;
  int b_hm = B.add_buf("hm", hm_flen, 0, false, SPLIT);
  int b_dhm = B.add_buf("dhm", Nm, 0, true, SPLIT);
  std::vector<int> b_cg(nG, -1), b_mg(nG, -1), b_ag(nG, -1), b_rg(nG, -1), b_dcg(nG, -1), b_dag(nG, -1);
  std::vector<int> ag_flen(nG), ag_off(nG), dcg_flen(nG), dcg_off(nG);
  for (int g = 0; g < nG - 1; g++) {
    char nm[32]; const GenL& l = G[g]; int L = l.Ho * l.Co;
    snprintf(nm, sizeof nm, "c_g%d", g);    b_cg[g] = B.add_buf(nm, L, 0, false);
    snprintf(nm, sizeof nm, "mean_g%d", g); b_mg[g] = B.add_buf(nm, 1, 0, false);
    const GenL& nx = G[g + 1];
    if (!nx.dense) { ag_flen[g] = (-nx.lo + l.Ho + nx.hi) * l.Co; ag_off[g] = -nx.lo * l.Co; }
    else { ag_flen[g] = L; ag_off[g] = 0; }
    snprintf(nm, sizeof nm, "a_g%d", g);    b_ag[g] = B.add_buf(nm, ag_flen[g], 0, false, SPLIT);
    snprintf(nm, sizeof nm, "rstd_g%d", g); b_rg[g] = B.add_buf(nm, 1, 0, false);
    dcg_flen[g] = (l.cl + l.Ho + (l.k - l.s - l.cl)) * l.Co; dcg_off[g] = l.cl * l.Co;
    // parity-split dgrad (below): its last 16-element tap reaches past the k*Co window -- those elements meet
    // zero weights, so they must be zeros of this frame's own plane, never a neighbour's bits (NaN * 0)
    if (gen_parity_split(l.Co, l.s, l.Hi, use_umma)) dcg_flen[g] += cdiv(l.k * l.Co, 16) * 16 - l.k * l.Co;
    snprintf(nm, sizeof nm, "dc_g%d", g);   b_dcg[g] = B.add_buf(nm, dcg_flen[g], 0, true, SPLIT);
    snprintf(nm, sizeof nm, "da_g%d", g);   b_dag[g] = B.add_buf(nm, L, 0, true); p.bufs[b_dag[g]].alias = b_da;
  }
  // reconstruction rows at the 32-byte aligned pitch of dxh: the last layer's epilogue stores 32 bytes per thread (a
  // 513-float pitch left every row on 4-byte stores); the pad floats are never written nor read
  p.buf_xh = B.add_buf("xh", xhld, 0, false);
  int b_dxh = B.add_buf("dxh", xhld, 0, true, SPLIT);

  // ---------------------------------------------------------------- ops
  { Op& o = B.op(OP_PACK, PH_PACK, "pack"); o.count = p.arena_w; }

  {  // P[s,:] = emb[s,:] . BY + b1 + b2 + b3   (model/vae.py:51-61 with the y-branch hoisted per speaker)
    Op& o = B.op(OP_GEMM, PH_PACK, "ptab");
    o.rows_fixed = a.y_dim; o.A = B.view(B.th(poff(P_emb)), 1, z, 0, 0, z); o.K = z;
    o.B = B.aw(A_bz[1]); o.ldb = Nm; o.N = Nm; o.C = B.view(B.aw(A_bz[0] + (int64_t)z * Nm), 1, Nm, 0, 0, Nm);
    for (int w = 0; w < 3; w++) o.bias[w] = B.aw(A_bm[w]);
    o.bias_mod = Nm; o.a_scalar = 1;   // theta offsets are not 16B aligned in general
  }
  if (use_umma) B.op(OP_PACK16, PH_PACK, "pack16");
  // encoder: conv (F) + Layernorm + lrelu   (util/layers.py:47-66, model/vae.py:74-78)
  // LN forward keeps (mean, rstd) per frame; backward recomputes xhat = (c - mean) * rstd from the raw conv
  // output c_l (LN_BWD: `xhat` = c_l, r0 = mean) instead of storing a second activation-sized tensor
  std::vector<View> VA_e(nE);
  for (int e = 0; e < nE; e++) {
    const EncL& l = E[e]; char nm[32];
    if (e == 0) VA_e[e] = B.view(B.user(U_X), l.Ho, l.Hi, l.s * l.Ci, -l.pl * l.Ci, l.Hi * l.Ci, 1);
    else VA_e[e] = B.view(B.ws(b_ae[e - 1]), l.Ho, ae_flen[e - 1], l.s * l.Ci, 0, ae_flen[e - 1]);
    snprintf(nm, sizeof nm, "conv_e%d", e);
    Op& o = B.op(OP_GEMM, PH_ENC, nm);
    o.A = VA_e[e]; o.K = l.k * l.Ci; o.a_scalar = (e == 0); o.B = B.aw(A_we[e]); o.ldb = l.Co; o.N = l.Co;
    o.tap_T = l.k; o.tap_C = l.Ci; o.tap_s = l.s;
    o.C = B.view(B.ws(b_ce[e]), l.Ho, l.Ho * l.Co, l.Co, 0, l.Ho * l.Co);
    o.bias[0] = B.th(poff(P_eb[e])); o.bias_mod = l.Co;
    // first layer: conv + Layernorm + lrelu as one CUDA-core kernel (fused_e0.cuh) when the register-resident
    // Layernorm mapping applies
    const bool fuse_fwd = fuse && e == 0 && l.Ci == 1 && l.k <= 8 && ln_group(l.Ho * l.Co, l.Co, ae_off[e], ae_flen[e]) > 0;
    if (fuse_fwd) p.ops.back().fuse = FUSE_E0_FWD;
    else if (fuse) p.ops.back().fuse = FUSE_LN_FWD;
    snprintf(nm, sizeof nm, "ln_e%d", e);
    Op& q = B.op(OP_LN_FWD, PH_ENC, nm);
    q.in = B.ws(b_ce[e]); q.r0 = B.ws(b_me[e]); q.aout = B.ws(b_ae[e]); q.rstd = B.ws(b_re[e]);
    q.gamma = B.th(poff(P_es[e])); q.beta = B.th(poff(P_eo[e]));
    q.L = l.Ho * l.Co; q.Cn = l.Co; q.out_flen = ae_flen[e]; q.out_off = ae_off[e];
  }
  View V_af = B.view(B.ws(b_ae[nE - 1]), 1, flat, 0, 0, flat);
  {  // heads: (mu | lv) = f . BH + (b_mu | b_lv)   (model/vae.py:79-81)
    Op& o = B.op(OP_GEMM, PH_ENC, "heads");
    o.A = V_af; o.K = flat; o.B = B.aw(A_bh); o.ldb = 2 * z; o.N = 2 * z;
    o.C = B.view(B.ws(p.buf_hz), 1, 2 * z, 0, 0, 2 * z); o.bias[0] = B.aw(A_bhb); o.bias_mod = 2 * z;
  }
  {  // GaussianSampleLayer + GaussianKLD  (util/layers.py:152-156,170-183)
    Op& o = B.op(OP_SAMPLE, PH_SAMPLE, "sample_kl");
    o.r0 = B.ws(p.buf_hz); o.r1 = B.ws(p.buf_mu); o.r2 = B.ws(p.buf_lv); o.r3 = B.ws(p.buf_z); o.i0 = z;
  }
  // generator  (model/vae.py:84-103)
  // zero padding of the merge output (the merge GEMM writes the interior [hm_off, hm_off + Nm) of every frame): i0 / i1 = the
  // interior, so only the pads around it need clearing
  { Op& o = B.op(OP_ZERO, PH_DEC, "zero_hm"); o.r0 = B.ws(b_hm); o.count = hm_flen; o.per_frame_count = 1; o.i0 = hm_off; o.i1 = Nm; }
  {  // zs = [z | one-hot(y)]  (model/vae.py:64-70,89-90: embedding lookup + the two FCs of _merge)
    Op& o = B.op(OP_ZCAT, PH_DEC, "zcat"); o.r0 = B.ws(p.buf_z); o.r1 = B.ws(b_zs); o.i0 = z; o.i1 = ypad;
  }
  View V_z = B.view(B.ws(b_zs), 1, zk, 0, 0, zk);
  View V_hm_rows = B.view(B.ws(b_hm), 1, hm_flen, 0, hm_off, hm_flen);
  {
    Op& o = B.op(OP_GEMM, PH_DEC, "merge");
    o.A = V_z; o.K = zk; o.B = B.aw(A_bz[0]); o.ldb = Nm; o.N = Nm; o.C = V_hm_rows;
  }
  std::vector<View> VA_g(nG);
  for (int g = 0; g < nG; g++) {
    const GenL& l = G[g]; char nm[32];
    Ref src = (g == 0) ? B.ws(b_hm) : B.ws(b_ag[g - 1]);
    int sflen = (g == 0) ? hm_flen : ag_flen[g - 1];
    snprintf(nm, sizeof nm, "convT_g%d", g);
    Op& o = B.op(OP_GEMM, PH_DEC, nm);
    if (!l.dense) {
      VA_g[g] = B.view(src, l.Hi, sflen, l.Cip, g == 0 ? hm_lead : 0, sflen);
      o.A = VA_g[g]; o.K = l.wn * l.Cip; o.B = B.aw(A_gf[g]); o.ldb = ld_gf[g]; o.N = l.s * l.Co;
      o.tap_T = l.wn; o.tap_C = l.Cip; o.tap_s = 1;
      o.C = B.view(B.ws(b_cg[g]), l.Hi, l.Ho * l.Co, l.s * l.Co, 0, l.Ho * l.Co);
      o.bias[0] = B.th(poff(P_gb[g])); o.bias_mod = l.Co;
      if (fuse) p.ops.back().fuse = FUSE_LN_FWD;
      snprintf(nm, sizeof nm, "ln_g%d", g);
      Op& q = B.op(OP_LN_FWD, PH_DEC, nm);
      q.in = B.ws(b_cg[g]); q.r0 = B.ws(b_mg[g]); q.aout = B.ws(b_ag[g]); q.rstd = B.ws(b_rg[g]);
      q.gamma = B.th(poff(P_gs[g])); q.beta = B.th(poff(P_go[g]));
      q.L = l.Ho * l.Co; q.Cn = l.Co; q.out_flen = ag_flen[g]; q.out_off = ag_off[g];
    } else {
      VA_g[g] = B.view(src, 1, sflen, 0, 0, sflen);
      o.A = VA_g[g]; o.K = l.Hi * l.Ci; o.B = B.aw(A_gf[g]); o.ldb = ld_gf[g]; o.N = l.Ho * l.Co;
      o.C = B.view(B.ws(p.buf_xh), 1, xhld, 0, 0, xhld);
      o.bias[0] = B.th(poff(P_gb[g])); o.bias_mod = l.Co;
    }
  }
  {  // GaussianLogDensity + d/dxh  (util/layers.py:159-167, model/vae.py:120-128)
    Op& o = B.op(OP_RECON, PH_LOSS, "recon");
    o.r0 = B.user(U_X); o.r1 = B.ws(p.buf_xh); o.r2 = B.ws(b_dxh); o.r3 = B.gr(poff(P_gb[nG - 1]));
    o.i0 = a.in_h; o.i1 = xhld;
  }
  // ---- backward (autodiff of trainer/vae.py:24)
  {
    const GenL& l = G[nG - 1];
    View V_dxh_rows = B.view(B.ws(b_dxh), 1, xhld, 0, 0, xhld);
    Op& w = B.op(OP_WGRAD, PH_BWD, "wgrad_g_last");
    w.A = VA_g[nG - 1]; w.K = l.Hi * l.Ci; w.C = V_dxh_rows; w.N = l.Ho * l.Co; w.B = B.adw(A_gf[nG - 1]); w.ldb = ld_gf[nG - 1];
    Op& o = B.op(OP_GEMM, PH_BWD, "dgrad_g_last");
    o.A = V_dxh_rows; o.K = l.Ho * l.Co; o.B = B.aw(A_gd[nG - 1]); o.ldb = ld_gd[nG - 1]; o.N = l.Hi * l.Ci;
    Ref dst = (nG >= 2) ? B.ws(b_dag[nG - 2]) : B.ws(b_dhm);
    o.C = B.view(dst, 1, l.Hi * l.Ci, 0, 0, l.Hi * l.Ci);
  }
  for (int g = nG - 2; g >= 0; g--) {
    const GenL& l = G[g]; char nm[32];
    snprintf(nm, sizeof nm, "ln_bwd_g%d", g);
    Op& q = B.op(OP_LN_BWD, PH_BWD, nm);
    q.in = B.ws(b_dag[g]); q.xhat = B.ws(b_cg[g]); q.r0 = B.ws(b_mg[g]); q.rstd = B.ws(b_rg[g]); q.aout = B.ws(b_dcg[g]);
    q.gamma = B.th(poff(P_gs[g])); q.beta = B.th(poff(P_go[g]));
    q.dgamma = B.gr(poff(P_gs[g])); q.dbeta = B.gr(poff(P_go[g])); q.dbias = B.gr(poff(P_gb[g]));
    q.L = l.Ho * l.Co; q.Cn = l.Co; q.out_flen = dcg_flen[g]; q.out_off = dcg_off[g];
    snprintf(nm, sizeof nm, "wgrad_g%d", g);
    Op& w = B.op(OP_WGRAD, PH_BWD, nm);
    w.A = VA_g[g]; w.K = l.wn * l.Cip; w.N = l.s * l.Co;
    w.C = B.view(B.ws(b_dcg[g]), l.Hi, dcg_flen[g], l.s * l.Co, dcg_off[g], dcg_flen[g]);
    w.B = B.adw(A_gf[g]); w.ldb = ld_gf[g];
    // dgrad of the transposed conv = strided conv over the padded gradient.  With 8 channels a tap is half an
    // MMA K step: split the rows by parity, so that each half sees 16-element taps (2 positions) at a whole
    // number of taps per row step (2*s*Co/16) -- two GEMMs over interleaved views of the same buffers.
    const int halves = gen_parity_split(l.Co, l.s, l.Hi, use_umma) ? 2 : 1;
    for (int hf = 0; hf < halves; hf++) {
      if (halves == 1) snprintf(nm, sizeof nm, "dgrad_g%d", g); else snprintf(nm, sizeof nm, "dgrad_g%d_%s", g, hf ? "odd" : "even");
      Op& o = B.op(OP_GEMM, PH_BWD, nm);
      const int R = (halves == 1) ? l.Hi : (hf ? l.Hi / 2 : (l.Hi + 1) / 2);
      o.A = B.view(B.ws(b_dcg[g]), R, dcg_flen[g], halves * l.s * l.Co, hf * l.s * l.Co, dcg_flen[g]); o.K = l.k * l.Co;
      if (halves == 1) { o.tap_T = l.k; o.tap_C = l.Co; o.tap_s = l.s; }
      else { o.tap_C = 16; o.tap_T = cdiv(l.k * l.Co, 16); o.tap_s = 2 * l.s * l.Co / 16; }
      o.B = B.aw(A_gd[g]); o.ldb = ld_gd[g]; o.N = l.Cip;
      Ref dst = (g > 0) ? B.ws(b_dag[g - 1]) : B.ws(b_dhm);
      o.C = B.view(dst, R, l.Hi * l.Cip, halves * l.Cip, hf * l.Cip, l.Hi * l.Cip);
    }
  }
  View V_dhm = B.view(B.ws(b_dhm), 1, Nm, 0, 0, Nm);
  {  // merge backward; rows z.. of the weight gradient = per-speaker sums of dhm (IndexedSlices of
     // embedding_lookup with duplicates summed)
    Op& w = B.op(OP_WGRAD, PH_BWD, "wgrad_merge_z"); w.A = V_z; w.K = zk; w.C = V_dhm; w.N = Nm; w.B = B.adw(A_bz[0]); w.ldb = Nm;
    Op& o = B.op(OP_GEMM, PH_BWD, "dgrad_merge_z"); o.A = V_dhm; o.K = Nm; o.B = B.aw(A_bzd[0]); o.ldb = z; o.N = z;
    o.C = B.view(B.ws(b_dz), 1, z, 0, 0, z);
  }
  {  // sampler + KL backward -> (dmu | dlv), head-bias grads
    Op& o = B.op(OP_SAMPLE_BWD, PH_BWD, "sample_bwd");
    o.r0 = B.ws(b_dz); o.r1 = B.ws(p.buf_hz); o.r2 = B.ws(b_dhz); o.r3 = B.adw(A_bhb); o.i0 = z;
  }
  {
    View V_dhz = B.view(B.ws(b_dhz), 1, 2 * z, 0, 0, 2 * z);
    Op& w = B.op(OP_WGRAD, PH_BWD, "wgrad_heads"); w.A = V_af; w.K = flat; w.C = V_dhz; w.N = 2 * z; w.B = B.adw(A_bh); w.ldb = 2 * z;
    Op& o = B.op(OP_GEMM, PH_BWD, "dgrad_heads"); o.A = V_dhz; o.K = 2 * z; o.B = B.aw(A_hd); o.ldb = flat; o.N = flat;
    o.C = B.view(B.ws(b_dae[nE - 1]), 1, flat, 0, 0, flat);
  }
  for (int e = nE - 1; e >= 0; e--) {
    const EncL& l = E[e]; char nm[32];
    snprintf(nm, sizeof nm, "ln_bwd_e%d", e);
    Op& q = B.op(OP_LN_BWD, PH_BWD, nm);
    q.in = B.ws(b_dae[e]); q.xhat = B.ws(b_ce[e]); q.r0 = B.ws(b_me[e]); q.rstd = B.ws(b_re[e]); q.aout = B.ws(b_dce[e]);
    q.gamma = B.th(poff(P_es[e])); q.beta = B.th(poff(P_eo[e]));
    q.dgamma = B.gr(poff(P_es[e])); q.dbeta = B.gr(poff(P_eo[e])); q.dbias = B.gr(poff(P_eb[e]));
    q.L = l.Ho * l.Co; q.Cn = l.Co; q.out_flen = dce_flen[e]; q.out_off = dce_off[e];
    // first layer: no data gradient, so dc_e0 is read by the weight gradient alone -- one kernel keeps it in registers
    if (fuse && e == 0 && l.Ci == 1 && l.k <= 8 && e0_bwd_group(l.Ho * l.Co, l.Co) > 0) { q.fuse = FUSE_E0_BWD; p.bufs[b_dce[e]].elide = 1; }
    snprintf(nm, sizeof nm, "wgrad_e%d", e);
    Op& w = B.op(OP_WGRAD, PH_BWD, nm);
    w.A = VA_e[e]; w.K = l.k * l.Ci; w.a_scalar = (e == 0); w.N = l.Co;
    w.C = B.view(B.ws(b_dce[e]), l.Ho, dce_flen[e], l.Co, dce_off[e], dce_flen[e]);
    w.B = B.adw(A_we[e]); w.ldb = l.Co;
    if (e > 0) {
      snprintf(nm, sizeof nm, "dgrad_e%d", e);
      Op& o = B.op(OP_GEMM, PH_BWD, nm);
      int R = e_q1[e] - e_q0[e] + 1;
      o.A = B.view(B.ws(b_dce[e]), R, dce_flen[e], l.Co, (e_q0[e] - (e_wn[e] - 1) + e_pf[e]) * l.Co, dce_flen[e]);
      o.K = e_wn[e] * l.Co; o.B = B.aw(A_ed[e]); o.ldb = l.s * l.Ci; o.N = l.s * l.Ci;
      o.tap_T = e_wn[e]; o.tap_C = l.Co; o.tap_s = 1;
      o.C = B.view(B.ws(b_dae[e - 1]), R, l.Hi * l.Ci, l.s * l.Ci, (l.s * e_q0[e] - l.pl) * l.Ci, l.Hi * l.Ci, 1);
    }
  }
  {  // once per call, after all chunks: y-branch grads from the per-speaker sums, then unpack
    View V_emb = B.view(B.th(poff(P_emb)), 1, z, 0, 0, z);
    View V_dptab = B.view(B.adw(A_bz[0] + (int64_t)z * Nm), 1, Nm, 0, 0, Nm);   // rows z.. of the merge weight gradient
    Op& w = B.op(OP_WGRAD, PH_FINAL, "wgrad_merge_y"); w.rows_fixed = a.y_dim; w.A = V_emb; w.K = z; w.a_scalar = 1;
    w.C = V_dptab; w.N = Nm; w.B = B.adw(A_bz[1]); w.ldb = Nm;
    Op& o = B.op(OP_GEMM, PH_FINAL, "dgrad_emb"); o.rows_fixed = a.y_dim; o.A = V_dptab; o.K = Nm;
    o.B = B.aw(A_bzd[1]); o.ldb = z; o.N = z; o.C = B.view(B.gr(poff(P_emb)), 1, z, 0, 0, z);
    Op& c = B.op(OP_COLSUM, PH_FINAL, "colsum_dptab"); c.r0 = B.adw(A_bz[0] + (int64_t)z * Nm); c.r1 = B.adw(A_bm[0]); c.i0 = Nm; c.i1 = a.y_dim;
    if (fuse && a.y_dim <= 16) p.ops[p.ops.size() - 3].fuse = FUSE_SPK_BWD;      // wgrad_merge_y + dgrad_emb + colsum_dptab
    Op& u = B.op(OP_UNPACK, PH_FINAL, "unpack"); u.count = p.n_params;
  }

  // ---------------------------------------------------------------- tcgen05 routing
  // (F) ops: A = bf16 hi / lo planes read by TMA through the strided view (whole frames per <= 128-row
  // tile), B = K-major [N, kpad] bf16 hi / lo packs.  (W) ops: both operands are the split views
  // themselves (the same TMA boxes, consumed as MN-major operands).  bf16x3: hi.hi + hi.lo + lo.hi.
  p.aw16_off = p.arena_w;
  if (use_umma) {
    if (p.n_params > PACK_INDEX_MASK || p.arena_w > PACK_INDEX_MASK) return "too many parameters for the pack index encoding";
    const int64_t ptab_lo = A_bz[0] + (int64_t)z * Nm, ptab_hi = ptab_lo + (int64_t)a.y_dim * Nm;
    auto view_ok = [&](const View& v) {
      return v.split && !v.pred && v.off >= 0 && v.off % 8 == 0 && v.rs % 8 == 0 && v.fs % 8 == 0 && umma_row_tile(v.R) > 0;
    };
    auto a16_alloc = [&](int64_t n) {
      int64_t off = p.aw16_count; p.aw16_count = rup64(off + n, 64);
      p.pack16_src.resize(p.aw16_count, -1); return off;
    };
    for (Op& o : p.ops) {
      if (o.rows_fixed || o.a_scalar) continue;
      if (o.kind == OP_WGRAD) {
        if (!view_ok(o.A) || !view_ok(o.C) || o.A.R != o.C.R || o.K < 8 || o.N < 8) continue;
        o.umma = 1;
        continue;
      }
      if (o.kind != OP_GEMM || !view_ok(o.A) || o.K < 8 || o.N < 8 || o.B.space != SP_AW) continue;
      o.kpad = rup(o.K, 8);
      const int64_t sz = (int64_t)o.N * o.kpad;
      o.bu_hi = a16_alloc(sz);                 // the lo pack mirrors it half a region further on (set below)
      for (int n = 0; n < o.N; n++) for (int k = 0; k < o.K; k++) {
        const int64_t q = o.B.off + (int64_t)k * o.ldb + n;
        int32_t src = p.pack_src[q];
        if (src < 0) {
          // rows computed on the device at pack time (the per-speaker table): source = the fp32 pack itself
          if (!(q >= ptab_lo && q < ptab_hi)) continue;
          src = (int32_t)q | PACK16_FROM_ARENA;
        }
        p.pack16_src[o.bu_hi + (int64_t)n * o.kpad + k] = src;
      }
      o.umma = 1;
    }
    // bf16 region = [hi packs of every op | lo packs in the same layout]: one table entry (and one gather) per
    // element writes both planes
    const int64_t half = p.aw16_count;
    for (Op& o : p.ops) if (o.umma && o.kind == OP_GEMM) o.bu_lo = o.bu_hi + half;
    p.aw16_count = 2 * half;
    p.arena_w = rup64(p.aw16_off + (p.aw16_count + 1) / 2, 64);
    // fp32 packs nobody reads: the B operand of an op that always runs on the tensor path (its bf16 packs are gathered
    // from theta directly).  Few-tap ops that may fall back to the row kernel keep theirs.
    std::vector<uint8_t> need(p.aw16_off, 1);
    for (const Op& o : p.ops)
      if (o.umma && o.kind == OP_GEMM && !(o.K <= 64 && o.N <= 32))
        for (int64_t k = 0; k < o.K; k++) for (int n = 0; n < o.ldb; n++) {
          const int64_t q = o.B.off + k * o.ldb + n;
          if (q < (int64_t)need.size() && !(q >= ptab_lo && q < ptab_hi)) need[q] = 0;
        }
    for (const Op& o : p.ops)          // ... unless another op reads the same pack without the tensor path
      if (o.kind == OP_GEMM && !o.umma && o.B.space == SP_AW)
        for (int64_t k = 0; k < o.K; k++) for (int n = 0; n < o.ldb; n++) { const int64_t q = o.B.off + k * o.ldb + n; if (q < (int64_t)need.size()) need[q] = 1; }
    for (int64_t q = 0; q < p.aw16_off; q++) if (need[q]) p.pack_list.push_back((int32_t)q);     // (structural zeros included)
  }

  // ---------------------------------------------------------------- JSON
  std::ostringstream js;
  js << "{\"n_params\":" << p.n_params << ",\"arena_w\":" << p.arena_w << ",\"arena_dw\":" << p.arena_dw
     << ",\"aw16_off\":" << p.aw16_off << ",\"aw16_count\":" << p.aw16_count
     << ",\"z\":" << z << ",\"out_dim\":" << p.out_dim << ",\"xh_ld\":" << xhld << ",\"params\":[";
  for (size_t i = 0; i < p.params.size(); i++) {
    const Param& q = p.params[i];
    js << (i ? "," : "") << "{\"name\":\"" << q.name << "\",\"off\":" << q.off << ",\"size\":" << q.size << ",\"shape\":[";
    for (int d = 0; d < q.rank; d++) js << (d ? "," : "") << q.shape[d];
    js << "],\"fan_in\":" << q.fan_in << ",\"fan_out\":" << q.fan_out << ",\"init\":" << q.init << "}";
  }
  js << "],\"bufs\":[";
  for (size_t i = 0; i < p.bufs.size(); i++) {
    const Buf& q = p.bufs[i];
    js << (i ? "," : "") << "{\"name\":\"" << q.name << "\",\"per_frame\":" << q.per_frame << ",\"fixed\":" << q.fixed
       << ",\"train_only\":" << q.train_only << ",\"split\":" << q.split << ",\"elide\":" << q.elide << ",\"alias\":" << q.alias << "}";
  }
  js << "],\"ops\":[";
  for (size_t i = 0; i < p.ops.size(); i++) {
    const Op& o = p.ops[i];
    js << (i ? "," : "") << "{\"kind\":" << o.kind << ",\"phase\":" << o.phase << ",\"fuse\":" << o.fuse << ",\"name\":\"" << o.name << "\",";
    json_view(js, "A", o.A); js << ","; json_view(js, "C", o.C);
    js << ",\"K\":" << o.K << ",\"N\":" << o.N << ","; json_ref(js, "B", o.B);
    js << ",\"ldb\":" << o.ldb << ","; json_ref(js, "bias0", o.bias[0]); js << ","; json_ref(js, "bias1", o.bias[1]);
    js << ","; json_ref(js, "bias2", o.bias[2]); js << ",\"bias_mod\":" << o.bias_mod << ",";
    js << "\"rows_fixed\":" << o.rows_fixed
       << ",\"a_scalar\":" << o.a_scalar << ",\"umma\":" << o.umma << ",\"bu_hi\":" << o.bu_hi << ",\"bu_lo\":" << o.bu_lo
       << ",\"kpad\":" << o.kpad << ",\"tap\":[" << o.tap_T << "," << o.tap_C << "," << o.tap_s << "],";
    json_ref(js, "in", o.in); js << ","; json_ref(js, "xhat", o.xhat); js << ","; json_ref(js, "aout", o.aout); js << ",";
    json_ref(js, "rstd", o.rstd); js << ","; json_ref(js, "gamma", o.gamma); js << ","; json_ref(js, "beta", o.beta); js << ",";
    json_ref(js, "dgamma", o.dgamma); js << ","; json_ref(js, "dbeta", o.dbeta); js << ","; json_ref(js, "dbias", o.dbias);
    js << ",\"L\":" << o.L << ",\"Cn\":" << o.Cn << ",\"out_flen\":" << o.out_flen << ",\"out_off\":" << o.out_off << ",";
    json_ref(js, "r0", o.r0); js << ","; json_ref(js, "r1", o.r1); js << ","; json_ref(js, "r2", o.r2); js << ",";
    json_ref(js, "r3", o.r3);
    js << ",\"count\":" << o.count << ",\"per_frame_count\":" << o.per_frame_count << ",\"i0\":" << o.i0 << ",\"i1\":" << o.i1 << "}";
  }
  js << "]}";
  p.json = js.str();
  return "";
}

}  // namespace npvc
