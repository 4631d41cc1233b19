// Plan interpreter + C-ABI of libnpvc_b200.so (see include/npvc_b200.h).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges are emitted only with NPVC_NVTX=1 (ncu --nvtx --print-nvtx-rename kernel)

#include <cstdlib>
#include <utility>
#include <map>

#include "kernels.cuh"
#include "fused_e0.cuh"
#include "plan.h"
#include "umma_gemm.cuh"

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

using namespace npvc;

struct npvc_handle {
  Plan plan;
  int64_t max_chunk = 16384;
  int32_t* d_pack_src = nullptr;
  int32_t* d_pack16_src = nullptr;
  int32_t* d_pack_list = nullptr;
  int32_t* d_unpack_ptr = nullptr;
  int32_t* d_unpack_idx = nullptr;
  int32_t* d_heavy = nullptr; int n_heavy = 0;     // parameters with >= UNPACK_HEAVY packed-gradient entries
  bool tables_on_device = false;
  int64_t launches = 0;
  int64_t last_chunk = 0; bool last_train = false;
  int sm_count = 148;
  bool use_umma = true;
  std::string umma_allow;            // debug: comma-separated op names allowed on the tensor path ("" = all)
  PFN_tmapEncodeTiled encode = nullptr;
  struct TMaps { const void* a; const void* b; long long frames; int bn, rows_tile, sw; CUtensorMap tAh, tAl, tBh, tBl; };
  std::map<int, TMaps> tmaps;        // per-op tensor-map cache
  std::map<int, TMaps> tmaps_pair;   // same, CTA-pair launches (B boxes of BN / 2 rows)
  int fuse_ln_train = 0;             // NPVC_FUSE_LN_TRAIN=1: the Layernorm epilogue in training passes too (A/B comparisons; measured slower)
  bool attr_fwd_ln = false;
  int umma_pair = 1;                 // NPVC_PAIR=0: no cta_group::2 CTA pairs; 2: every BN >= 128 layer (A/B comparisons)
  std::string pair_ops;              // NPVC_PAIR_OPS: comma-separated op names for the pair form (overrides the shape rule)
  int wgrad_pair = -1;               // cta_group::2 form of the weight-gradient kernel: -1 (default) where it measured faster (launch_umma_wgrad),
                                     // NPVC_WGRAD_PAIR=0 never, 1 every N >= 128 layer, 2 the same with 256-column N tiles (A/B comparisons)
  bool attr_fwd = false, attr_pair = false, attr_wgrad = false, attr_wgrad_pair = false, attr_ln_bulk = false, attr_e0_bwd = false, attr_ln_reg = false;   // cudaFuncSetAttribute done (per handle = per device)
  int64_t umma_launches = 0;
  int ln_bulk = 1;                   // double-buffered bulk-copy Layernorm backward for frames > 2048 floats
  int wgrad_smem_kb = 225;           // shared-memory budget of the weight-gradient kernel
  long long prefetch_e0_min = 0, prefetch_min = 0;     // frames from which the Layernorm-backward kernels prefetch (PREFETCH_MIN_FRAMES; set in npvc_create)
  int umma_dual = 1;                 // NPVC_UMMA_DUAL=0: one MMA issuer warp in tap mode (A/B comparisons)
  int umma_bres = 1;                 // NPVC_UMMA_BRES=0: window mode re-loads the weight tiles with every stage (A/B comparisons)
  int overlap_wgrad = 1;             // NPVC_OVERLAP=0: weight gradients on the caller's stream (A/B comparisons, per-op profiling)
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pack = nullptr;
  bool pack_defer = false, pack_pending = false;   // training call: the pack ops after the first run beside the first layer (run_phase)
  int umma_groups = 4;               // epilogue groups of the forward kernel (<= accumulator sets)
  int umma_tap = 1;                  // NPVC_UMMA_TAP=0: conv-shaped layers through the overlapping-window boxes (A/B comparisons)
  int umma_merge = 1;                // NPVC_UMMA_MERGE=0: three MMAs per K step instead of two (A/B comparisons; see UmmaArgs::merge)
  int nvtx = 0;                      // NPVC_NVTX=1: an NVTX push / pop range named after the plan op around every launch
  bool profiling = false;
  struct Ev { int op; cudaEvent_t a, b; long long rows, frames; };
  std::vector<Ev> events;
  std::string profile_json;
};

static thread_local std::string g_err;
static int fail(int code, const std::string& m) { g_err = m; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(NPVC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

namespace {

// Every launch of the library goes through here: programmatic stream serialisation (PDL) lets the blocks of a kernel be
// scheduled while the previous kernel of the stream drains; the kernels begin with griddepcontrol.wait (kernels.cuh,
// pdl_prologue), so only launch latency and block scheduling overlap, never the work.  NPVC_PDL=0 switches it off.
int g_pdl = 1;
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);       // (errors surface through cudaGetLastError at the call sites)
}

struct Ctx {
  npvc_handle* h; float* ws; int64_t chunk_cap; bool train;
  const float* theta; float* grad; const float* x; const int64_t* y; const float* eps;
  int64_t n;          // frames in this chunk
  int64_t n_total;    // frames the loss means span
  cudaStream_t st;
  const StepState* state = nullptr;   // in-kernel sampler (eps == nullptr): device-resident {seed, draws, step}
  long long frame0 = 0;               // index of this chunk's first frame in the sampler's counter
};
inline float* shared_ws(const Ctx& c) { return c.ws; }
inline int tmap_key(const Ctx&, int op_index) { return op_index; }

float* resolve(const Ctx& c, const Ref& r) {
  const Plan& p = c.h->plan;
  switch (r.space) {
    case SP_WS: return (p.bufs[r.buf].per_frame == 0 ? shared_ws(c) : c.ws) + p.buf_offset(r.buf, c.chunk_cap, c.train);
    case SP_THETA: return const_cast<float*>(c.theta) + r.off;
    case SP_GRAD: return c.grad ? c.grad + r.off : nullptr;
    case SP_AW: return shared_ws(c) + r.off;
    case SP_ADW: return shared_ws(c) + p.buf_offset(p.buf_adw, c.chunk_cap, c.train) + r.off;
    case SP_USER:
      if (r.buf == U_X) return const_cast<float*>(c.x);
      if (r.buf == U_EPS) return const_cast<float*>(c.eps);
      return nullptr;
    default: return nullptr;
  }
}
DView dview(const Ctx& c, const View& v) {
  DView d; d.p = resolve(c, v.ref); d.fs = v.fs; d.R = v.R; d.rs = v.rs; d.off = v.off; d.flen = v.flen; d.pred = v.pred;
  d.split = v.split;
  return d;
}
bool view_vec_ok(const DView& d) {
  return !d.pred && ((reinterpret_cast<uintptr_t>(d.p) & 15) == 0) && (d.fs % 4 == 0) && (d.rs % 4 == 0) && (d.off % 4 == 0);
}

template <int BN>
void launch_gemm_bn(const GemmArgs& g, bool scalar, cudaStream_t st) {
  dim3 grid((unsigned)((g.rows + 127) / 128), (unsigned)((g.N + BN - 1) / BN));
  if (scalar) launch_k(gemm_view_kernel<BN, true>, dim3(grid), dim3(256), 0, st, g);
  else launch_k(gemm_view_kernel<BN, false>, dim3(grid), dim3(256), 0, st, g);
}
void launch_gemm(const GemmArgs& g, bool scalar, cudaStream_t st) {
  if (g.N >= 96) launch_gemm_bn<128>(g, scalar, st);
  else if (g.N >= 48) launch_gemm_bn<64>(g, scalar, st);
  else if (g.N >= 24) launch_gemm_bn<32>(g, scalar, st);
  else launch_gemm_bn<16>(g, scalar, st);
}

template <int BTK, int BTN>
void launch_wgrad_t(WgradArgs g, bool scalar, int sms, cudaStream_t st) {
  int tk = (g.K + BTK - 1) / BTK, tn = (g.N + BTN - 1) / BTN;
  g.tiles_n = tn;
  long long tiles = (long long)tk * tn;
  long long splits = (4LL * sms + tiles - 1) / tiles;
  long long max_splits = (g.rows + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long rps = (g.rows + splits - 1) / splits;
  rps = (rps + 15) / 16 * 16;
  splits = (g.rows + rps - 1) / rps;
  g.rows_per_split = rps;
  dim3 grid((unsigned)tiles, (unsigned)splits);
  if (scalar) launch_k(wgrad_view_kernel<BTK, BTN, true>, dim3(grid), dim3(256), 0, st, g);
  else launch_k(wgrad_view_kernel<BTK, BTN, false>, dim3(grid), dim3(256), 0, st, g);
}
template <int BTK>
void launch_wgrad_k(const WgradArgs& g, bool scalar, int sms, cudaStream_t st) {
  if (g.N > 64) launch_wgrad_t<BTK, 128>(g, scalar, sms, st);
  else if (g.N > 32) launch_wgrad_t<BTK, 64>(g, scalar, sms, st);
  else if (g.N > 16) launch_wgrad_t<BTK, 32>(g, scalar, sms, st);
  else launch_wgrad_t<BTK, 16>(g, scalar, sms, st);
}
void launch_wgrad(const WgradArgs& g, bool scalar, int sms, cudaStream_t st) {
  if (g.K > 64) launch_wgrad_k<128>(g, scalar, sms, st);
  else if (g.K > 16) launch_wgrad_k<64>(g, scalar, sms, st);
  else launch_wgrad_k<16>(g, scalar, sms, st);
}

// ---- tcgen05 path ---------------------------------------------------------------------------
// N tile of the window-mode forward kernel: the fewest padded columns among tile counts near N / cap.  cap = 256 is
// the MMA limit; bn_cap() lowers it to 128 for the short-K layers: 2 * 2 * BN <= 512 TMEM columns leave room for two
// accumulator sets, so the epilogue of a tile overlaps the next tile's mainloop (measured on a B200,
// profiles/r2a_switches.txt: E4 dgrad 0.099 -> 0.068 ms, E4 0.072 -> 0.064, E3 dgrad 0.078 -> 0.065, heads dgrad
// 0.049 -> 0.044; the long-K G3 forward and the 4104-column G3 dgrad were faster with wide tiles and keep them)
// Frame prefetch of the Layernorm-backward kernels (kernels.cuh ln_bwd_reg_kernel, fused_e0.cuh e0_bwd_kernel): the next
// frame's dy / c land in shared memory (cp.async) while the current one is computed.  Timed alone the kernels gain (first
// layer 0.153 -> 0.099 ms, the five register-resident layers -0.02 ms together), but the step does not: their 50-60 KB of
// shared memory per block keep them from sharing SMs with the side stream's weight-gradient GEMMs (cfg1: 0.955 ms with the
// prefetch, 0.895 ms without; cfg2, 60-step runs alternated on one box: 3.644 ms without, 3.644 / 3.663 ms with either
// one), so it is OFF unless NPVC_PREFETCH_MIN / NPVC_PREFETCH_E0_MIN (frames from which it applies) ask for it; the GPU
// switch-equivalence test runs it.
constexpr long long PREFETCH_MIN_FRAMES = 1LL << 62;
inline int bn_cap(const Op& o) { return ((o.K <= 1024 && o.N <= 1024) || o.K <= 256) ? 128 : 256; }
int pick_bn(int N, int cap, int* n_tiles) {
  if (N <= cap) { *n_tiles = 1; return (N + 15) / 16 * 16; }
  int best_bn = cap, best_t = (N + cap - 1) / cap; long long best_cost = (long long)best_bn * best_t + (cap < 256 ? 8LL * best_t : 0LL);
  const int t0 = (N + cap - 1) / cap;
  for (int t = t0; t <= t0 + 6; t++) {
    int bn = ((N + t - 1) / t + 15) / 16 * 16;
    if (bn > cap) continue;
    long long cost = (long long)bn * t + (cap < 256 ? 8LL * t : 0LL);     // (capped tiles: do not trade tile width for a few padded columns)
    if (cost < best_cost) { best_cost = cost; best_bn = bn; best_t = t; }
  }
  *n_tiles = best_t; return best_bn;
}

// rows of a view -> tiles of whole (frame, row-group) TMA boxes, <= row_target (<= 128) rows each
RowTiling make_tiling(int R, long long frames, int row_target) {
  RowTiling t;
  t.Rb = umma_row_tile(R); t.Ra = R / t.Rb;
  if (t.Ra == 1) { t.FB = row_target / t.Rb; if (t.FB < 1) t.FB = 1; t.Ab = 1; }
  else { t.FB = 1; t.Ab = row_target / t.Rb; if (t.Ab < 1) t.Ab = 1; if (t.Ab > t.Ra) t.Ab = t.Ra; }
  t.TA = (t.Ra + t.Ab - 1) / t.Ab;
  t.RbH = t.Rb;
  t.rows_tile = t.Rb * t.Ab * t.FB;
  t.frames = (int)frames; t.m_tiles = (int)(((frames + t.FB - 1) / t.FB) * t.TA);
  return t;
}

// 4-D tensor maps (k, row-in-group, row-group, frame) over the hi / lo planes of a split view
int make_view_maps(npvc_handle* h, const Ctx& c, const View& v, int extent, const RowTiling& rt, int box_inner, int sw_bytes,
                   CUtensorMap* hi, CUtensorMap* lo, const std::string& name) {
  uint16_t* base = reinterpret_cast<uint16_t*>(resolve(c, v.ref)) + v.off;
  const cuuint64_t fsb = (cuuint64_t)v.fs * 4;                     // frame stride in bytes (2 planes of fs bf16)
  cuuint64_t gd[4] = {(cuuint64_t)extent, (cuuint64_t)rt.Rb, (cuuint64_t)rt.Ra, (cuuint64_t)rt.frames};
  cuuint64_t gs[3] = {v.R == 1 ? fsb : (cuuint64_t)v.rs * 2, rt.Ra == 1 ? fsb : (cuuint64_t)rt.Rb * v.rs * 2, fsb};
  cuuint32_t bx[4] = {(cuuint32_t)box_inner, (cuuint32_t)rt.Rb, (cuuint32_t)rt.Ab, (cuuint32_t)rt.FB};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  for (int w = 0; w < 2; w++) {
    CUresult r = h->encode(w ? lo : hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base + (w ? v.fs : 0), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled(view) failed for " + name + " code " + std::to_string((int)r));
  }
  return NPVC_OK;
}

// Tap mode of the forward kernel (umma_gemm.cuh): eligible when the view is a conv window with 16 / 32 / 64
// channels per tap, one N tile, and the weights of all taps plus >= 2 activation stages fit shared memory.
constexpr int LN_EPI_SMEM = 32768;     // shared memory of the fused Layernorm epilogue (umma_gemm.cuh)

// Layernorm + lrelu in the forward kernel's epilogue: `ln` is the OP_LN_FWD that follows the GEMM `o` in the plan
// (Op::fuse == FUSE_LN_FWD).  Needs whole frames per M tile, whole rows per N tile and the dense row-major conv output
// the Layernorm op reads; anything else runs the two ops as two kernels.
// Used for inference passes: there the raw conv output is never stored and a whole pass per layer disappears (cfg3:
// 11.26 -> 10.44 ms).  In training the raw output must be kept for the backward, the saving is one read per layer, and
// the longer epilogue costs more than the separate Layernorm kernels did (measured: E1 0.121 -> 0.126 ms, E2 0.089 ->
// 0.104, G1 0.146 -> 0.166, profiles/r2g_ops_per_step.txt) -- training keeps the two kernels.
bool ln_epilogue_ok(const Ctx& c, const Op& o, const Op* ln, const RowTiling& rt, int n_tiles) {
  const Plan& p = c.h->plan;
  if (c.train && c.h->fuse_ln_train == 0) return false;
  if (!ln || ln->kind != OP_LN_FWD || rt.Ra != 1 || n_tiles != 1) return false;
  if (o.C.pred || o.C.split || o.C.off != 0 || o.C.rs != o.N || (o.C.fs % 4) || (o.N % 8) || o.N > 256) return false;
  if (ln->in.space != SP_WS || o.C.ref.space != SP_WS || ln->in.buf != o.C.ref.buf) return false;
  if (ln->L != o.A.R * o.N || (o.N % ln->Cn) || (ln->out_off % 8) || (ln->out_flen % 8)) return false;
  if (ln->aout.space != SP_WS || !p.bufs[ln->aout.buf].split) return false;
  return true;
}

struct TapGeom { bool ok; int sw, P, hr, BN, b_tile_al, stages; RowTiling rt; };
TapGeom tap_geometry(const Op& o, long long frames, int reserve = 0) {
  TapGeom t; memset(&t, 0, sizeof t);
  const int C = o.tap_C, T = o.tap_T, s = o.tap_s;
  if (T <= 0 || !(C == 16 || C == 32 || C == 64) || o.N > 256 || o.A.R < 2 || o.A.rs != s * C || o.K > T * C || o.K <= (T - 1) * C) return t;   // (a last tap may be partly beyond K: its weights are TMA zero fill)
  t.sw = 2 * C; t.P = s; t.hr = (T - 1) / s;
  t.BN = (o.N + 15) / 16 * 16;
  t.b_tile_al = (t.BN * t.sw + 1023) / 1024 * 1024;
  RowTiling& r = t.rt;
  r.Rb = umma_row_tile(o.A.R); if (r.Rb <= 0) return t;
  r.Ra = o.A.R / r.Rb; r.RbH = r.Rb + t.hr;
  if (r.RbH > 128) return t;
  if (r.Ra == 1) { r.FB = 128 / r.RbH; r.Ab = 1; } else { r.FB = 1; r.Ab = 128 / r.RbH; if (r.Ab > r.Ra) r.Ab = r.Ra; }
  r.TA = (r.Ra + r.Ab - 1) / r.Ab;
  r.rows_tile = r.RbH * r.Ab * r.FB;
  r.frames = (int)frames; r.m_tiles = (int)(((frames + r.FB - 1) / r.FB) * r.TA);
  const int bres = T * 2 * t.b_tile_al, stage = t.P * 2 * 128 * t.sw;
  t.stages = (225 * 1024 - 6144 - reserve - bres) / stage; if (t.stages > 8) t.stages = 8;
  t.ok = t.stages >= 2;
  return t;
}

void fill_ln_epilogue(Ctx& c, const Op& ln, UmmaArgs& g) {
  g.ln.on = 1; g.ln.store_c = c.train ? 1 : 0;               // inference never reads the raw conv output again
  g.ln.aout = resolve(c, ln.aout); g.ln.mean = resolve(c, ln.r0); g.ln.rstd = resolve(c, ln.rstd);
  g.ln.gamma = resolve(c, ln.gamma); g.ln.beta = resolve(c, ln.beta);
  g.ln.Cn = ln.Cn; g.ln.L = ln.L; g.ln.out_flen = ln.out_flen; g.ln.out_off = ln.out_off;
}

int launch_umma_tap(Ctx& c, const Op& o, int op_index, const TapGeom& tg_in, const Op* ln, bool* fused) {
  npvc_handle* h = c.h; cudaStream_t st = c.st;
  TapGeom tg = tg_in;
  bool fuse_ln = ln_epilogue_ok(c, o, ln, tg.rt, 1);
  if (fuse_ln) { const TapGeom t2 = tap_geometry(o, c.n, LN_EPI_SMEM); if (t2.ok) tg = t2; else fuse_ln = false; }
  const RowTiling& rt = tg.rt; const int BN = tg.BN, sw = tg.sw, C = o.tap_C;
  const void* a_base = resolve(c, o.A.ref);
  uint16_t* arena16 = reinterpret_cast<uint16_t*>(shared_ws(c) + h->plan.aw16_off);
  uint16_t* b_hi = arena16 + o.bu_hi; uint16_t* b_lo = arena16 + o.bu_lo;
  const int key = tmap_key(c, op_index);
  auto it = h->tmaps.find(key);
  if (it == h->tmaps.end() || it->second.a != a_base || it->second.b != b_hi || it->second.frames != rt.frames || it->second.bn != BN || it->second.sw != -sw) {
    npvc_handle::TMaps tm; tm.a = a_base; tm.b = b_hi; tm.frames = rt.frames; tm.bn = BN; tm.rows_tile = rt.rows_tile; tm.sw = -sw;   // (negative: tap-mode maps)
    // A: (column within the s*C wide position group, row-in-group incl. halo, row-group, frame); groups overlap by the halo
    uint16_t* base = reinterpret_cast<uint16_t*>(resolve(c, o.A.ref)) + o.A.off;
    const cuuint64_t fsb = (cuuint64_t)o.A.fs * 4;
    cuuint64_t gd[4] = {(cuuint64_t)(o.tap_s * C), (cuuint64_t)rt.RbH, (cuuint64_t)rt.Ra, (cuuint64_t)rt.frames};
    cuuint64_t gs[3] = {(cuuint64_t)o.A.rs * 2, rt.Ra == 1 ? fsb : (cuuint64_t)rt.Rb * o.A.rs * 2, fsb};
    cuuint32_t bx[4] = {(cuuint32_t)C, (cuuint32_t)rt.RbH, (cuuint32_t)rt.Ab, (cuuint32_t)rt.FB};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle swz = sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    for (int w = 0; w < 2; w++) {
      CUresult r = h->encode(w ? &tm.tAl : &tm.tAh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base + (w ? o.A.fs : 0), gd, gs, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled(tap A) failed for " + o.name + " code " + std::to_string((int)r));
    }
    cuuint64_t gdB[2] = {(cuuint64_t)o.kpad, (cuuint64_t)o.N};
    cuuint64_t gsB[1] = {(cuuint64_t)o.kpad * 2};
    cuuint32_t bxB[2] = {(cuuint32_t)C, (cuuint32_t)BN};
    for (int w = 0; w < 2; w++) {
      CUresult r = h->encode(w ? &tm.tBl : &tm.tBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w ? b_lo : b_hi, gdB, gsB, bxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled(tap B) failed for " + o.name + " code " + std::to_string((int)r));
    }
    h->tmaps[key] = tm; it = h->tmaps.find(key);
  }
  UmmaArgs g; memset(&g, 0, sizeof g);
  g.K = o.K; g.N = o.N; g.BN = BN; g.kblocks = 1; g.rt = rt; g.n_tiles = 1; g.sw = sw;
  g.tapT = o.tap_T; g.tapC = C; g.tapP = tg.P; g.b_tile_al = tg.b_tile_al;
  g.acc_sets = 512 / (2 * BN) >= 4 ? 4 : (512 / (2 * BN) >= 2 ? 2 : 1);
  int tc = 32; while (tc < g.acc_sets * 2 * BN) tc *= 2; g.tmem_cols = tc;
  g.stages = tg.stages;
  g.merge = (h->umma_merge && 2 * BN <= 256 && tg.b_tile_al == BN * sw) ? 1 : 0;     // (hi and lo weight tiles back to back)
  g.C = dview(c, o.C);
  g.bias0 = resolve(c, o.bias[0]); g.bias1 = resolve(c, o.bias[1]); g.bias2 = resolve(c, o.bias[2]); g.bias_mod = o.bias_mod;
  if (fuse_ln) { fill_ln_epilogue(c, *ln, g); if (fused) *fused = true; }
  if (!h->attr_fwd) {
    CUDA_TRY(cudaFuncSetAttribute(umma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->attr_fwd = true;
  }
  const size_t smem = (size_t)o.tap_T * 2 * tg.b_tile_al + (size_t)tg.stages * tg.P * 2 * 128 * sw + 1024 + 8 * (2 * tg.stages + 11) + 32 + 4096 + (fuse_ln ? LN_EPI_SMEM : 0);
  unsigned grid = (unsigned)(rt.m_tiles < h->sm_count ? rt.m_tiles : h->sm_count);
  // second MMA issuer warp (umma_gemm.cuh, tap mode): even / odd tiles of a CTA issued by two warps.  Measured at cfg2:
  // the narrow layers gain (G2 forward 0.145 -> 0.123 ms, its dgrads 0.095 -> 0.074 / 0.097 -> 0.091, G1 dgrad -0.004),
  // the 96-column ones with 2 accumulator sets lose (E2 dgrad 0.047 -> 0.064, G0 dgrad 0.070 -> 0.087: each issuer then
  // owns ONE set and ONE or two stages and stalls where a single issuer ran ahead) -- so: 4 sets and >= 4 stages
  const bool dual = h->umma_dual && g.acc_sets >= 4 && tg.stages >= 4 && rt.m_tiles > (long long)grid;
  const int issuers = dual ? 2 : 1;        // (four issuers = 21 warps of 96 registers: more than the SM's four 16 K-register partitions hold; not offered)
  const unsigned threads = 64 + 128 * (g.acc_sets < h->umma_groups ? g.acc_sets : h->umma_groups) + 32 * (issuers - 1);
  if (fuse_ln) {
    if (!h->attr_fwd_ln) { CUDA_TRY(cudaFuncSetAttribute(umma_fwd_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); h->attr_fwd_ln = true; }
    launch_k(umma_fwd_ln_kernel, dim3(grid), dim3(threads), smem, st, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
  } else
  launch_k(umma_fwd_kernel, dim3(grid), dim3(threads), smem, st, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
  h->launches++; h->umma_launches++;
  return NPVC_OK;
}

// CTA-pair form of the window-mode forward kernel (umma_gemm.cuh, PAIR): clusters of 2 CTAs, each staging its own
// 128-row A tile and half of the B tile per k-block for one 256 x BN cta_group::2 MMA.
// Default rule = the shapes measured faster on a B200 (profiles/r1i_pair_check_*.txt, 16,384 frames): wide N tiles
// (BN >= 128: the B fill dominates and halves), long reductions (>= 8 k-blocks: G3 forward 0.268 -> 0.189 ms,
// G3 dgrad 0.276 -> 0.253, E4 0.076 -> 0.071, merge dgrad 0.034 -> 0.030; the 3-k-block merge GEMM got slower)
// and enough M tiles to fill the 74 pairs (small batches keep the single-CTA form, which is bit-identical per frame).
bool pair_wanted(const npvc_handle* h, const Op& o, int BN, int m_tiles) {
  if (!h->umma_pair || m_tiles < 2 || o.K <= 32 || (BN & 15)) return false;
  if (!h->pair_ops.empty()) return ("," + h->pair_ops + ",").find("," + o.name + ",") != std::string::npos;
  const int bn_min = bn_cap(o) < 256 ? 96 : 128;          // (capped N tiles of 112 columns still pair)
  if (h->umma_pair >= 2) return BN >= bn_min;
  return BN >= bn_min && o.K > 7 * 64 && m_tiles >= 64;
}
int launch_umma_pair(Ctx& c, const Op& o, int op_index, int BN, int n_tiles, const RowTiling& rt) {
  npvc_handle* h = c.h; cudaStream_t st = c.st;
  const int sw = 128, bk = 64;
  const void* a_base = resolve(c, o.A.ref);
  uint16_t* arena16 = reinterpret_cast<uint16_t*>(shared_ws(c) + h->plan.aw16_off);
  uint16_t* b_hi = arena16 + o.bu_hi; uint16_t* b_lo = arena16 + o.bu_lo;
  const int key = tmap_key(c, op_index);
  auto it = h->tmaps_pair.find(key);
  if (it == h->tmaps_pair.end() || it->second.a != a_base || it->second.b != b_hi || it->second.frames != rt.frames || it->second.bn != BN) {
    npvc_handle::TMaps tm; tm.a = a_base; tm.b = b_hi; tm.frames = rt.frames; tm.bn = BN; tm.rows_tile = rt.rows_tile; tm.sw = sw;
    int rc = make_view_maps(h, c, o.A, o.K, rt, bk, sw, &tm.tAh, &tm.tAl, o.name); if (rc) return rc;
    cuuint64_t gdB[2] = {(cuuint64_t)o.kpad, (cuuint64_t)o.N};
    cuuint64_t gsB[1] = {(cuuint64_t)o.kpad * 2};
    cuuint32_t bxB[2] = {(cuuint32_t)bk, (cuuint32_t)(BN / 2)};         // each CTA of the pair loads half of the N tile
    cuuint32_t es[2] = {1, 1};
    for (int w = 0; w < 2; w++) {
      CUresult r = h->encode(w ? &tm.tBl : &tm.tBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w ? b_lo : b_hi, gdB, gsB, bxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled(pair B) failed for " + o.name + " code " + std::to_string((int)r));
    }
    h->tmaps_pair[key] = tm; it = h->tmaps_pair.find(key);
  }
  UmmaArgs g; memset(&g, 0, sizeof g);
  g.K = o.K; g.N = o.N; g.BN = BN; g.kblocks = (o.K + bk - 1) / bk; g.rt = rt; g.n_tiles = n_tiles; g.sw = sw;
  const int stage_bytes = 2 * 128 * sw + BN * sw;                        // A hi / lo + this CTA's half of B hi / lo
  g.acc_sets = 512 / (2 * BN) >= 4 ? 4 : (512 / (2 * BN) >= 2 ? 2 : 1);
  int tc = 32; while (tc < g.acc_sets * 2 * BN) tc *= 2; g.tmem_cols = tc;
  int stages = (225 * 1024 - 6144) / stage_bytes; if (stages > 10) stages = 10; if (stages < 1) stages = 1;
  g.stages = stages;
  g.C = dview(c, o.C);
  g.bias0 = resolve(c, o.bias[0]); g.bias1 = resolve(c, o.bias[1]); g.bias2 = resolve(c, o.bias[2]); g.bias_mod = o.bias_mod;
  if (!h->attr_pair) {
    CUDA_TRY(cudaFuncSetAttribute(umma_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->attr_pair = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 11) + 32 + 4096;
  const long long pair_tiles = (long long)((rt.m_tiles + 1) / 2) * n_tiles;
  const long long pairs = pair_tiles < h->sm_count / 2 ? pair_tiles : h->sm_count / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3((unsigned)(64 + 128 * (g.acc_sets < h->umma_groups ? g.acc_sets : h->umma_groups)));
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, umma_fwd_pair_kernel, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g));
  h->launches++; h->umma_launches++;
  return NPVC_OK;
}

int launch_umma(Ctx& c, const Op& o, int op_index, const Op* ln = nullptr, bool* fused = nullptr) {
  npvc_handle* h = c.h; cudaStream_t st = c.st;
  const long long frames = c.n;
  if (h->umma_tap) {
    const TapGeom tg = tap_geometry(o, frames);
    if (tg.ok) return launch_umma_tap(c, o, op_index, tg, ln, fused);
  }
  int n_tiles = 1; int BN = pick_bn(o.N, bn_cap(o), &n_tiles);
  const RowTiling rt = make_tiling(o.A.R, frames, 128);
  // small batches (the reference trains on 16 frames per step): a handful of M tiles times a few wide N tiles leaves most
  // SMs idle while 3 CTAs stream the whole weight matrix -- narrower N tiles spread it (per-element arithmetic is the
  // same for any tile width: the K order does not change, results stay bit-identical)
  // (not where the Layernorm epilogue applies: it needs the whole row in one N tile, and a frame's result must not depend
  // on the size of the batch it arrives in -- fused and separate Layernorm differ in the last bits)
  const bool ln_tile = ln_epilogue_ok(c, o, ln, rt, n_tiles);
  while (!ln_tile && (long long)rt.m_tiles * n_tiles < h->sm_count / 2 && BN > 16) {
    BN = (BN / 2 + 15) / 16 * 16; n_tiles = (o.N + BN - 1) / BN;
  }
  if (pair_wanted(h, o, BN, rt.m_tiles)) {
    const int rc = launch_umma_pair(c, o, op_index, BN, n_tiles, rt);
    if (rc == NPVC_OK || h->umma_pair >= 2 || !h->pair_ops.empty()) return rc;     // (an explicit request reports its failure)
    // the cluster launch was refused (a device / partition that cannot co-schedule two such CTAs): the single-CTA
    // form of the same kernel computes bit-identical results -- use it from now on
    cudaGetLastError(); h->umma_pair = 0;
  }
  // k-block: 64 bf16 (128-byte swizzled rows); 32 (64-byte rows) only when K itself is that short
  // (32-wide k-blocks for deeper pipelines measured 20 % slower: more TMA requests per byte)
  int sw = 128;
  if (o.K <= 32) sw = 64;
  const int bk = sw / 2;
  const void* a_base = resolve(c, o.A.ref);
  uint16_t* arena16 = reinterpret_cast<uint16_t*>(shared_ws(c) + h->plan.aw16_off);
  uint16_t* b_hi = arena16 + o.bu_hi; uint16_t* b_lo = arena16 + o.bu_lo;
  const int key = tmap_key(c, op_index);
  auto it = h->tmaps.find(key);
  if (it == h->tmaps.end() || it->second.a != a_base || it->second.b != b_hi || it->second.frames != frames || it->second.bn != BN || it->second.sw != sw) {
    npvc_handle::TMaps tm; tm.a = a_base; tm.b = b_hi; tm.frames = frames; tm.bn = BN; tm.rows_tile = rt.rows_tile; tm.sw = sw;
    int rc = make_view_maps(h, c, o.A, o.K, rt, bk, sw, &tm.tAh, &tm.tAl, o.name); if (rc) return rc;
    cuuint64_t gdB[2] = {(cuuint64_t)o.kpad, (cuuint64_t)o.N};
    cuuint64_t gsB[1] = {(cuuint64_t)o.kpad * 2};
    cuuint32_t bxB[2] = {(cuuint32_t)bk, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    for (int w = 0; w < 2; w++) {
      CUresult r = h->encode(w ? &tm.tBl : &tm.tBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w ? b_lo : b_hi, gdB, gsB, bxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed for " + o.name + " code " + std::to_string((int)r));
    }
    h->tmaps[key] = tm; it = h->tmaps.find(key);
  }
  UmmaArgs g; memset(&g, 0, sizeof g);
  g.K = o.K; g.N = o.N; g.BN = BN; g.kblocks = (o.K + bk - 1) / bk; g.rt = rt; g.n_tiles = n_tiles; g.sw = sw;
  int stage_bytes = 2 * 128 * sw + 2 * BN * sw;
  g.acc_sets = 512 / (2 * BN) >= 4 ? 4 : (512 / (2 * BN) >= 2 ? 2 : 1);   // accumulator ring in TMEM: the epilogue of tile i overlaps the mainloops of the next tiles
  int tc = 32; while (tc < g.acc_sets * 2 * BN) tc *= 2; g.tmem_cols = tc;
  bool fuse_ln = ln_epilogue_ok(c, o, ln, rt, n_tiles) && (225 * 1024 - 6144 - LN_EPI_SMEM) / stage_bytes >= 2;
  const int budget = 225 * 1024 - 6144 - (fuse_ln ? LN_EPI_SMEM : 0);
  // resident weights (launch_args.h, b_res): one N tile, CTAs that run several tiles, and the tiles of all k-blocks fit in
  // front of >= 3 activation stages
  const int bres_bytes = g.kblocks * 2 * BN * sw;
  if (h->umma_bres && n_tiles == 1 && rt.m_tiles >= 2 * h->sm_count && (budget - bres_bytes) / (2 * 128 * sw) >= 3) {
    g.b_res = 1; stage_bytes = 2 * 128 * sw;
  }
  int stages = (budget - (g.b_res ? bres_bytes : 0)) / stage_bytes; if (stages > 10) stages = 10; if (stages < 1) stages = 1;
  g.stages = stages;
  g.merge = (h->umma_merge && 2 * BN <= 256) ? 1 : 0;
  g.C = dview(c, o.C);
  g.bias0 = resolve(c, o.bias[0]); g.bias1 = resolve(c, o.bias[1]); g.bias2 = resolve(c, o.bias[2]); g.bias_mod = o.bias_mod;
  if (fuse_ln) { fill_ln_epilogue(c, *ln, g); if (fused) *fused = true; }
  if (!h->attr_fwd) {
    CUDA_TRY(cudaFuncSetAttribute(umma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->attr_fwd = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + (g.b_res ? bres_bytes : 0) + 1024 + 8 * (2 * stages + 11) + 32 + 4096 + (fuse_ln ? LN_EPI_SMEM : 0);   // + bias_s[4][256]
  long long total = rt.m_tiles * n_tiles;
  unsigned grid = (unsigned)(total < h->sm_count ? total : h->sm_count);
  if (fuse_ln) {
    if (!h->attr_fwd_ln) { CUDA_TRY(cudaFuncSetAttribute(umma_fwd_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); h->attr_fwd_ln = true; }
    launch_k(umma_fwd_ln_kernel, dim3(grid), dim3(64 + 128 * (g.acc_sets < h->umma_groups ? g.acc_sets : h->umma_groups)), smem, st, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
  } else
  launch_k(umma_fwd_kernel, dim3(grid), dim3(64 + 128 * (g.acc_sets < h->umma_groups ? g.acc_sets : h->umma_groups)), smem, st, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
  h->launches++; h->umma_launches++;
  return NPVC_OK;
}

// dB[K,N] += A_view^T . dC_view on the tensor cores: both operands are the TMA boxes of the split views,
// consumed MN-major; the reduction over row tiles is split across CTAs, partial tiles added with RED.ADD.
int launch_umma_wgrad(Ctx& c, const Op& o, int op_index) {
  npvc_handle* h = c.h; cudaStream_t st = c.st;
  const long long frames = c.n;
  // N tile: 16 / 32 columns in one 32B / 64B-swizzled box, else 64-column boxes (128B swizzle)
  int BN, n_tiles = 1, d_sw;
  // CTA-pair form (cta_group::2, 256 x BN MMAs, each CTA half of the dC columns): measured on the B200 (profiles/
  // r2q_wgrad_pair.txt) it wins only where both dimensions are large -- the 4104 x 513 gradient of the last generator layer
  // (0.284 -> 0.257 ms) -- and loses on 896 x 256 / 144 x 1672 (E4, merge).  NPVC_WGRAD_PAIR=0 off, 1 / 2 everywhere it fits.
  const bool pair_shape = o.N >= 128 && o.K > 128;                     // (>= 2 K tiles, whole 64-column boxes per CTA)
  const bool pair = pair_shape && (h->wgrad_pair > 0 || (h->wgrad_pair < 0 && h->umma_pair != 0 && o.K >= 2048 && o.N >= 512 && frames * o.A.R >= 4096));
  if (pair) {
    d_sw = 128;
    const int t128 = (o.N + 127) / 128, t256 = (o.N + 255) / 256;
    if ((h->wgrad_pair >= 2 && o.N > 128) || 256 * t256 <= 128 * t128) { BN = 256; n_tiles = t256; } else { BN = 128; n_tiles = t128; }
  }
  else if (o.N <= 16) { BN = 16; d_sw = 32; }
  else if (o.N <= 32) { BN = 32; d_sw = 64; }
  else {
    d_sw = 128; BN = 64; long long best = -1;
    for (int bn = 64; bn <= 256; bn += 64) {
      int t = (o.N + bn - 1) / bn; long long cost = (long long)bn * t;
      if (best < 0 || cost * 100 <= best * 103) { if (best < 0 || cost < best) best = cost; BN = bn; n_tiles = t; }
    }
  }
  const int m_tiles_k = (o.K + 127) / 128;
  // rows per stage: keep >= 3 stages in shared memory
  const int a_boxes = (!pair && o.K <= 64) ? 1 : 2;                // 64-column A boxes per stage and plane
  const int per_row = pair ? 512 + 2 * BN : 256 * a_boxes + 4 * BN;   // (pair: each CTA stages half of the dC columns)
  const int budget = h->wgrad_smem_kb * 1024;       // < 227 KB leaves shared memory for a co-resident Layernorm block (side-stream overlap)
  int row_target = ((budget - 5 * 1024) / 3 / per_row) / 16 * 16; if (row_target > 128) row_target = 128; if (row_target < 16) row_target = 16;
  const RowTiling rt = make_tiling(o.A.R, frames, row_target);
  const void* a_base = resolve(c, o.A.ref); const void* d_base = resolve(c, o.C.ref);
  const int key = tmap_key(c, op_index);
  auto it = h->tmaps.find(key);             // (the boxes do not depend on the pair form; BN and rows_tile are part of the check)
  if (it == h->tmaps.end() || it->second.a != a_base || it->second.b != d_base || it->second.frames != frames || it->second.bn != BN ||
      it->second.rows_tile != rt.rows_tile) {
    npvc_handle::TMaps tm; tm.a = a_base; tm.b = d_base; tm.frames = frames; tm.bn = BN; tm.rows_tile = rt.rows_tile; tm.sw = 128;
    int rc = make_view_maps(h, c, o.A, o.K, rt, 64, 128, &tm.tAh, &tm.tAl, o.name); if (rc) return rc;
    rc = make_view_maps(h, c, o.C, o.N, rt, d_sw / 2, d_sw, &tm.tBh, &tm.tBl, o.name); if (rc) return rc;
    h->tmaps[key] = tm; it = h->tmaps.find(key);
  }
  UmmaArgs g; memset(&g, 0, sizeof g);
  g.K = o.K; g.N = o.N; g.BN = BN; g.rt = rt; g.n_tiles = n_tiles; g.d_sw = d_sw;
  g.rows_al = (rt.rows_tile + 15) / 16 * 16;
  g.a_boxes = a_boxes;
  g.merge = (h->umma_merge && !pair && 2 * BN <= 256 && BN % (d_sw / 2) == 0) ? 1 : 0;     // (whole dC boxes: the lo plane's boxes follow the hi plane's)
  const int d_boxes = pair ? BN / 128 : (BN + d_sw / 2 - 1) / (d_sw / 2);      // per CTA
  const int d_region = (g.rows_al * d_sw + 1023) / 1024 * 1024;
  const int stage_bytes = 2 * (a_boxes * g.rows_al * 128) + 2 * d_boxes * d_region;
  int tc = 32; while (tc < 2 * BN) tc *= 2; g.tmem_cols = tc;
  int stages = (budget - 3072) / stage_bytes; if (stages > 8) stages = 8; if (stages < 1) stages = 1;
  g.stages = stages;
  // split the reduction so that tiles * S fills whole waves of SMs, each CTA keeping enough row
  // tiles to amortise its 128 x BN atomic epilogue
  const int grid_x = pair ? 2 * ((m_tiles_k + 1) / 2) : m_tiles_k;
  const long long tiles = (long long)grid_x * n_tiles;
  const long long min_tiles = (512 + rt.rows_tile - 1) / rt.rows_tile;     // >= 512 rows per CTA
  long long maxS = rt.m_tiles / min_tiles; if (maxS < 1) maxS = 1; if (maxS > 4096) maxS = 4096;
  const double slots = (double)h->sm_count;
  long long S = 1; double best = 1e30;
  for (long long cand = 1; cand <= maxS; cand++) {
    double waves = (double)(tiles * cand) / slots;
    if (waves > 8.0 && cand > 1) break;
    double cost = (waves < 1.0 ? 1.0 / waves : std::ceil(waves) / waves) + 0.01 * waves;
    if (cost < best - 1e-9) { best = cost; S = cand; }
  }
  g.tiles_per_split = (int)((rt.m_tiles + S - 1) / S);
  S = (rt.m_tiles + g.tiles_per_split - 1) / g.tiles_per_split;
  g.out = resolve(c, o.B); g.ld = o.ldb;
  if (!h->attr_wgrad) {
    CUDA_TRY(cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->attr_wgrad = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 1) + 16;
  dim3 grid((unsigned)grid_x, (unsigned)n_tiles, (unsigned)S);
  if (pair) {
    if (!h->attr_wgrad_pair) {
      CUDA_TRY(cudaFuncSetAttribute(umma_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      h->attr_wgrad_pair = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, umma_wgrad_pair_kernel, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
    if (le != cudaSuccess) {
      if (h->wgrad_pair > 0) return fail(NPVC_ERR_CUDA, std::string("cluster launch of the pair weight-gradient kernel: ") + cudaGetErrorString(le));
      // default rule and the cluster launch was refused (a device / partition that cannot co-schedule two such CTAs): the
      // single-CTA form from now on (as the forward kernel does)
      cudaGetLastError(); h->wgrad_pair = 0;
      return launch_umma_wgrad(c, o, op_index);
    }
    h->launches++; h->umma_launches++;
    return NPVC_OK;
  }
  launch_k(umma_wgrad_kernel, dim3(grid), dim3(192), smem, st, it->second.tAh, it->second.tAl, it->second.tBh, it->second.tBl, g);
  h->launches++; h->umma_launches++;
  return NPVC_OK;
}

// ---- fused first encoder layer (fused_e0.cuh): `o` and `nx` are the two plan ops the kernel stands for --------------
int launch_e0_fwd(Ctx& c, const Op& o, const Op& nx) {      // o: conv (OP_GEMM on the caller's frames), nx: its OP_LN_FWD
  npvc_handle* h = c.h; const Plan& p = h->plan;
  E0FwdArgs g;
  g.x = c.x; g.W = resolve(c, o.B); g.bias = resolve(c, o.bias[0]); g.gamma = resolve(c, nx.gamma); g.beta = resolve(c, nx.beta);
  g.c = c.train ? resolve(c, nx.in) : nullptr;             // inference never reads the raw conv output again
  g.mean = resolve(c, nx.r0); g.rstd = resolve(c, nx.rstd); g.aout = resolve(c, nx.aout);
  g.Hi = (int)o.A.fs; g.Ho = o.A.R; g.Co = o.N; g.k = o.K; g.s = o.A.rs; g.pl = -o.A.off;
  g.out_flen = nx.out_flen; g.out_off = nx.out_off; g.out_split = p.bufs[nx.aout.buf].split; g.frames = c.n;
  if (!g.x) return fail(NPVC_ERR_ARG, "frames (x) required");
  g.xp = e0_row_floats(g.Ho, g.s, g.pl, g.Hi);
  const int G = ln_group(nx.L, nx.Cn, nx.out_off, nx.out_flen);
  const int bt = E0_BLOCK(G), fpb = bt / G;
  const long long fbs = (c.n + fpb - 1) / fpb;
  long long blocks = (long long)h->sm_count * 2 * (768 / bt); if (blocks > fbs) blocks = fbs;
  const size_t sm = ((size_t)(E0_KT + 3) * g.Co + (size_t)3 * fpb * g.xp) * sizeof(float);      // (three row buffers: fused_e0.cuh)
  const bool v4 = nx.L / 8 > 3 * G;                       // units of 8 elements per thread: 3 or 4
#define NPVC_E0_FWD(GG) do { if (v4) launch_k(e0_fwd_kernel<GG, 4>, dim3((unsigned)blocks), dim3(bt), sm, c.st, g); else launch_k(e0_fwd_kernel<GG, 3>, dim3((unsigned)blocks), dim3(bt), sm, c.st, g); } while (0)
  if (G == 32) NPVC_E0_FWD(32); else if (G == 64) NPVC_E0_FWD(64); else if (G == 128) NPVC_E0_FWD(128); else NPVC_E0_FWD(256);
#undef NPVC_E0_FWD
  h->launches++;
  return NPVC_OK;
}
int launch_e0_bwd(Ctx& c, const Op& o, const Op& nx) {      // o: OP_LN_BWD of the first layer, nx: its OP_WGRAD
  npvc_handle* h = c.h;
  E0BwdArgs g;
  g.x = c.x; g.dy = resolve(c, o.in); g.cin = resolve(c, o.xhat); g.mean = resolve(c, o.r0); g.rstd = resolve(c, o.rstd);
  g.gamma = resolve(c, o.gamma); g.beta = resolve(c, o.beta);
  g.dW = resolve(c, nx.B); g.dgamma = resolve(c, o.dgamma); g.dbeta = resolve(c, o.dbeta); g.dbias = resolve(c, o.dbias);
  g.Hi = (int)nx.A.fs; g.Ho = nx.A.R; g.Co = nx.N; g.k = nx.K; g.s = nx.A.rs; g.pl = -nx.A.off; g.frames = c.n;
  if (!g.x) return fail(NPVC_ERR_ARG, "frames (x) required");
  if (nx.ldb != nx.N) return fail(NPVC_ERR_ARG, "fused first-layer backward: packed weight gradient must be dense");
  g.xp = e0_row_floats(g.Ho, g.s, g.pl, g.Hi);
  const int G = e0_bwd_group(o.L, o.Cn);
  const int bt = E0_BLOCK(G), fpb = bt / G;
  const long long fbs = (c.n + fpb - 1) / fpb;
  long long blocks = (long long)h->sm_count * (512 / bt); if (blocks > fbs) blocks = fbs;
  g.prefetch = c.n >= h->prefetch_e0_min ? 1 : 0;
  const size_t sm = e0_bwd_smem_floats(g.Co, fpb, g.xp, o.L, g.prefetch) * sizeof(float);      // (rows and the prefetched dy / c: fused_e0.cuh)
  if (sm > 110 * 1024) return fail(NPVC_ERR_ARG, "fused first-layer backward: frame too long for the staging buffers");
  const bool v4 = o.L / 4 > 3 * G;                        // units of 4 elements per thread: 3 or 4
  const bool set_attr = !h->attr_e0_bwd; h->attr_e0_bwd = true;      // (one instantiation per handle: the architecture is fixed)
#define NPVC_E0_BWD_I(GG, VV) do { if (set_attr) CUDA_TRY(cudaFuncSetAttribute(e0_bwd_kernel<GG, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); \
                                   launch_k(e0_bwd_kernel<GG, VV>, dim3((unsigned)blocks), dim3(bt), sm, c.st, g); } while (0)
#define NPVC_E0_BWD(GG) do { if (v4) NPVC_E0_BWD_I(GG, 4); else NPVC_E0_BWD_I(GG, 3); } while (0)
  if (G == 32) NPVC_E0_BWD(32); else if (G == 64) NPVC_E0_BWD(64); else if (G == 128) NPVC_E0_BWD(128); else NPVC_E0_BWD(256);
#undef NPVC_E0_BWD
#undef NPVC_E0_BWD_I
  h->launches++;
  return NPVC_OK;
}

// speaker branch backward (kernels.cuh, speaker_bwd_kernel): w = wgrad_merge_y, d = dgrad_emb, cs = colsum_dptab
int launch_speaker_bwd(Ctx& c, const Op& w, const Op& d, const Op& cs) {
  npvc_handle* h = c.h;
  const int S = (int)w.rows_fixed, Z = w.K, Nm = w.N;
  if (d.K != Nm || d.N != Z || cs.i0 != Nm || w.ldb != Nm || d.ldb != Z) return fail(NPVC_ERR_ARG, "speaker backward: unexpected plan shapes");
  const float* dP = resolve(c, w.C.ref) + w.C.off;                 // per-speaker sums of the merge gradient [S, Nm]
  const float* emb = resolve(c, w.A.ref) + w.A.off;
  const int nA = (Nm + 255) / 256, nB = (Nm + SPK_CH - 1) / SPK_CH;
  const size_t sm = (size_t)S * (Z > SPK_CH ? Z : SPK_CH) * sizeof(float);
  launch_k(speaker_bwd_kernel, dim3((unsigned)(nA + nB)), dim3(256), sm, c.st, dP, emb, (int)w.A.fs, resolve(c, d.B), resolve(c, w.B),
           resolve(c, d.C.ref) + d.C.off, (int)d.C.fs, resolve(c, cs.r1), S, Z, Nm, nA);
  h->launches++;
  return NPVC_OK;
}

bool umma_allowed(const npvc_handle* h, const Op& o) {
  if (!o.umma || !h->encode) return false;
  if (h->umma_allow.empty()) return true;
  return ("," + h->umma_allow + ",").find("," + o.name + ",") != std::string::npos;
}

// ln / fused: the OP_LN_FWD that follows a GEMM marked FUSE_LN_FWD; *fused = true when the launch covered both ops
int run_op(Ctx& c, const Op& o, int op_index, const Op* ln = nullptr, bool* fused = nullptr) {
  npvc_handle* h = c.h; const Plan& p = h->plan; cudaStream_t st = c.st;
  switch (o.kind) {
    case OP_PACK: {
      // tensor path: only the fp32 packs some op reads (the bf16 packs are gathered from theta directly); with the
      // debugging switch NPVC_UMMA_OPS any op may fall back to the CUDA-core kernels, which read all of them
      const bool compact = h->use_umma && h->umma_allow.empty() && h->d_pack_list != nullptr;
      long long n = compact ? (long long)p.pack_list.size() : p.aw16_off;
      if (compact) launch_k(pack_list_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, c.theta, h->d_pack_src, h->d_pack_list, c.ws, n);
      else launch_k(pack_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, c.theta, h->d_pack_src, c.ws, n);
      h->launches++; break;
    }
    case OP_PACK16: {
      long long n = p.aw16_count / 2;
      if (n > 0) {
        launch_k(pack16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, c.theta, c.ws, h->d_pack16_src, reinterpret_cast<uint16_t*>(c.ws + p.aw16_off), n);
        h->launches++;
      }
      break;
    }
    case OP_ZCAT: {
      if (!c.y) return fail(NPVC_ERR_ARG, "labels (y) required");
      long long n4 = c.n * ((o.i0 + o.i1) / 4);
      if (n4 > 0) {
        launch_k(zcat_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, resolve(c, o.r0), reinterpret_cast<const long long*>(c.y), resolve(c, o.r1),
                                                                o.i0, o.i1, c.n, p.bufs[o.r1.buf].split);
        h->launches++;
      }
      break;
    }
    case OP_UNPACK: {
      long long n = p.n_params;
      const float* adw = shared_ws(c) + p.buf_offset(p.buf_adw, c.chunk_cap, c.train);
      launch_k(unpack_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, adw, h->d_unpack_ptr, h->d_unpack_idx, c.grad, n);
      h->launches++;
      if (h->n_heavy > 0) {
        launch_k(unpack_heavy_kernel, dim3((unsigned)((h->n_heavy * 32LL + 255) / 256)), dim3(256), 0, st, adw, h->d_unpack_ptr, h->d_unpack_idx, h->d_heavy, h->n_heavy, c.grad);
        h->launches++;
      }
      break;
    }
    case OP_GEMM: {
      GemmArgs g; g.A = dview(c, o.A); g.K = o.K; g.B = resolve(c, o.B); g.ldb = o.ldb; g.N = o.N; g.C = dview(c, o.C);
      g.rows = o.rows_fixed ? o.rows_fixed : c.n * o.A.R;
      g.bias0 = resolve(c, o.bias[0]); g.bias1 = resolve(c, o.bias[1]); g.bias2 = resolve(c, o.bias[2]); g.bias_mod = o.bias_mod;
      if (g.rows <= 0) break;
      const bool row_shaped = !g.bias1 && !g.bias2 && o.K <= 64 && o.N <= 32 && !o.rows_fixed && !g.C.split && !(g.A.split && g.A.pred);
      // few-tap, few-channel layers over millions of rows: through the overlapping-window boxes of the generic tensor
      // kernel they cost more than the thread-per-row FFMA kernel (measured) -- tensor cores only in tap mode
      const bool tensor_ok = umma_allowed(h, o) && (!row_shaped || (h->umma_tap && tap_geometry(o, c.n).ok));
      if (tensor_ok) { int rc = launch_umma(c, o, op_index, ln, fused); if (rc) return rc; break; }
      if (o.rows_fixed && o.rows_fixed <= 16 && !g.bias0 && o.C.ref.space == SP_GRAD && !g.A.pred && g.A.R == 1 && g.C.R == 1 && o.K >= 256) {
        // few-row GEMM accumulated into the (zero-initialised) gradient buffer
        const int kchunk = 32, ks = (o.K + kchunk - 1) / kchunk;
        dim3 grid((unsigned)((o.N + 127) / 128), (unsigned)ks);
        launch_k(fewrows_gemm_kernel, dim3(grid), dim3(128), (size_t)o.rows_fixed * kchunk * sizeof(float), st, 
            g.A.p + g.A.off, (int)g.A.fs, (int)o.rows_fixed, o.K, g.B, o.ldb, o.N, g.C.p + g.C.off, (int)g.C.fs, kchunk);
        h->launches++; break;
      }
      if (o.rows_fixed && o.rows_fixed <= 16 && o.K <= 512 && !g.A.pred && g.A.R == 1 && g.C.R == 1 && !g.A.split && !g.C.split && !g.C.pred) {
        // few rows, short K, plain store (+ biases): the per-speaker table at pack time
        launch_k(fewrows_fwd_kernel, dim3((unsigned)((o.N + 255) / 256)), dim3(256), (size_t)o.rows_fixed * o.K * sizeof(float), st,
                 g.A.p + g.A.off, (int)g.A.fs, (int)o.rows_fixed, o.K, g.B, o.ldb, o.N, g.C.p + g.C.off, (int)g.C.fs, g.bias0, g.bias1, g.bias2, o.bias_mod);
        h->launches++; break;
      }
      if (row_shaped) {     // (independent of n: per-frame results must not depend on the batch size)
        RowGemmArgs rg; rg.A = g.A; rg.K = o.K; rg.B = g.B; rg.ldb = o.ldb; rg.N = o.N; rg.C = g.C; rg.rows = g.rows;
        rg.bias0 = g.bias0; rg.bias_mod = o.bias_mod;
        const bool sc = !view_vec_ok(g.A) || (o.K % 4 != 0);
        const unsigned b1 = (unsigned)((g.rows + 255) / 256), b2 = (unsigned)((g.rows + 511) / 512);
        (void)b2;   // 2 rows / thread measured slower for the 48x24 / 56x16 shapes (170 registers)
        if (sc && o.K <= 8 && o.N <= 16) launch_k(rowgemm_kernel<8, 16, true, 2>, dim3(b2), dim3(256), 0, st, rg);
        else if (!sc && o.K <= 48 && o.N <= 24) launch_k(rowgemm_kernel<48, 24, false, 1>, dim3(b1), dim3(256), 0, st, rg);
        else if (!sc && o.K <= 56 && o.N <= 16) launch_k(rowgemm_kernel<56, 16, false, 1>, dim3(b1), dim3(256), 0, st, rg);
        else if (!sc) launch_k(rowgemm_kernel<64, 32, false, 1>, dim3(b1), dim3(256), 0, st, rg);
        else launch_k(rowgemm_kernel<64, 32, true, 1>, dim3(b1), dim3(256), 0, st, rg);
        h->launches++; break;
      }
      launch_gemm(g, !view_vec_ok(g.A), st); h->launches++; break;
    }
    case OP_WGRAD: {
      WgradArgs g; g.A = dview(c, o.A); g.K = o.K; g.D = dview(c, o.C); g.N = o.N; g.out = resolve(c, o.B); g.ld = o.ldb;
      g.rows = o.rows_fixed ? o.rows_fixed : c.n * o.A.R; g.rows_per_split = 0; g.tiles_n = 0;
      if (g.rows <= 0) break;
      if (!view_vec_ok(g.D)) return fail(NPVC_ERR_ARG, "wgrad dC view must be 16B aligned: " + o.name);
      if (umma_allowed(h, o)) { int rc = launch_umma_wgrad(c, o, op_index); if (rc) return rc; break; }
      if (o.K <= 8 && o.N <= 16 && g.rows >= 4096) {
        const long long blocks = (long long)h->sm_count * 4;
        long long rpb = (g.rows + blocks - 1) / blocks; rpb = (rpb + 255) / 256 * 256;
        launch_k(wgrad_tiny_kernel<8, 16>, dim3((unsigned)((g.rows + rpb - 1) / rpb)), dim3(256), 0, st, g, rpb);
        h->launches++; break;
      }
      if (o.K > 8 && o.K <= 48 && o.N <= 24 && o.K % 4 == 0 && o.N % 4 == 0 && view_vec_ok(g.A) && g.rows >= 4096 && !g.A.split && !g.D.split) {
        const long long blocks = (long long)h->sm_count * 8;
        long long rpb = (g.rows + blocks - 1) / blocks; rpb = (rpb + 63) / 64 * 64;
        launch_k(wgrad_small_kernel<48, 24>, dim3((unsigned)((g.rows + rpb - 1) / rpb)), dim3(192), 0, st, g, rpb);
        h->launches++; break;
      }
      launch_wgrad(g, !view_vec_ok(g.A), h->sm_count, st); h->launches++; break;
    }
    case OP_LN_FWD: {
      LnFwdArgs g; g.in = resolve(c, o.in); g.mean = resolve(c, o.r0); g.aout = resolve(c, o.aout);
      g.rstd = resolve(c, o.rstd); g.gamma = resolve(c, o.gamma); g.beta = resolve(c, o.beta);
      g.L = o.L; g.Cn = o.Cn; g.out_flen = o.out_flen; g.out_off = o.out_off; g.frames = c.n;
      g.out_split = p.bufs[o.aout.buf].split;
      const int G = ln_group(o.L, o.Cn, o.out_off, o.out_flen);
      if (G) {
        const long long fbs = (c.n + 256 / G - 1) / (256 / G);
        long long blocks = (long long)h->sm_count * 8; if (blocks > fbs) blocks = fbs;
        if (G == 32) launch_k(ln_fwd_reg_kernel<32>, dim3((unsigned)blocks), dim3(256), 0, st, g);
        else if (G == 64) launch_k(ln_fwd_reg_kernel<64>, dim3((unsigned)blocks), dim3(256), 0, st, g);
        else if (G == 128) launch_k(ln_fwd_reg_kernel<128>, dim3((unsigned)blocks), dim3(256), 0, st, g);
        else launch_k(ln_fwd_reg_kernel<256>, dim3((unsigned)blocks), dim3(256), 0, st, g);
      } else launch_k(ln_fwd_kernel, dim3((unsigned)c.n), dim3(256), (size_t)o.L * sizeof(float), st, g);
      h->launches++; break;
    }
    case OP_LN_BWD: {
      LnBwdArgs g; g.prefetch = 0; g.dy = resolve(c, o.in); g.cin = resolve(c, o.xhat); g.mean = resolve(c, o.r0); g.rstd = resolve(c, o.rstd);
      g.gamma = resolve(c, o.gamma); g.beta = resolve(c, o.beta); g.dc = resolve(c, o.aout);
      g.dgamma = resolve(c, o.dgamma); g.dbeta = resolve(c, o.dbeta); g.dbias = resolve(c, o.dbias);
      g.L = o.L; g.Cn = o.Cn; g.out_flen = o.out_flen; g.out_off = o.out_off; g.frames = c.n;
      g.out_split = p.bufs[o.aout.buf].split;
      // frames of <= 2048 floats: register-resident kernel; larger frames keep more bytes in flight per SM
      // through the shared-memory kernel (measured)
      int G = ln_group(o.L, o.Cn, o.out_off, o.out_flen); if (G > 64) G = 0;
      if (G) {
        const long long fbs = (c.n + 256 / G - 1) / (256 / G);
        long long blocks = (long long)h->sm_count * 4; if (blocks > fbs) blocks = fbs;
        g.prefetch = c.n >= h->prefetch_min ? 1 : 0;
        const size_t sm = ((size_t)5 * o.Cn + (g.prefetch ? (size_t)(256 / G) * 2 * o.L : (size_t)0)) * sizeof(float);      // + landing zone (<= 64 KB)
        if (!h->attr_ln_reg) {
          CUDA_TRY(cudaFuncSetAttribute(ln_bwd_reg_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          CUDA_TRY(cudaFuncSetAttribute(ln_bwd_reg_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          h->attr_ln_reg = true;
        }
        if (G == 32) launch_k(ln_bwd_reg_kernel<32>, dim3((unsigned)blocks), dim3(256), sm, st, g);
        else launch_k(ln_bwd_reg_kernel<64>, dim3((unsigned)blocks), dim3(256), sm, st, g);
      } else if (ln_group(o.L, o.Cn, o.out_off, o.out_flen) && 2048 % o.Cn == 0 && h->ln_bulk &&
                 (size_t)(4 * o.L + 5 * o.Cn) * sizeof(float) <= 100 * 1024) {
        // large frames: double-buffered bulk-async frame stream (3 blocks / SM at L = 4104)
        const size_t sm = (size_t)(4 * o.L + 5 * o.Cn) * sizeof(float);
        if (!h->attr_ln_bulk) { CUDA_TRY(cudaFuncSetAttribute(ln_bwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); h->attr_ln_bulk = true; }
        int per_sm = (int)((220 * 1024) / (sm + 1024)); if (per_sm > 3) per_sm = 3; if (per_sm < 1) per_sm = 1;
        long long blocks = (long long)h->sm_count * per_sm; if (blocks > c.n) blocks = c.n;
        launch_k(ln_bwd_bulk_kernel, dim3((unsigned)blocks), dim3(256), sm, st, g);
      } else {
        long long blocks = (long long)h->sm_count * 8; if (blocks > c.n) blocks = c.n;
        launch_k(ln_bwd_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(2 * o.L + 3 * o.Cn) * sizeof(float), st, g);
      }
      h->launches++; break;
    }
    case OP_SAMPLE: {
      const int z = o.i0, fpb = 8;
      double* acc = reinterpret_cast<double*>(shared_ws(c) + p.buf_offset(p.buf_acc, c.chunk_cap, c.train));
      launch_k(sample_kl_kernel, dim3((unsigned)((c.n + fpb - 1) / fpb)), dim3(2 * z), 0, st, 
          resolve(c, o.r0), c.eps, c.state, c.frame0, resolve(c, o.r1), resolve(c, o.r2), resolve(c, o.r3), (c.eps || c.state) ? acc : nullptr, z, c.n, fpb);
      h->launches++; break;
    }
    case OP_SAMPLE_BWD: {
      const int z = o.i0, fpb = 16;
      launch_k(sample_bwd_kernel, dim3((unsigned)((c.n + fpb - 1) / fpb)), dim3(2 * z), 0, st, 
          resolve(c, o.r0), c.eps, c.state, c.frame0, resolve(c, o.r1), resolve(c, o.r2), resolve(c, o.r3), z, c.n, fpb, 1.0f / (float)c.n_total, p.bufs[o.r2.buf].split);
      h->launches++; break;
    }
    case OP_RECON: {
      double* acc = reinterpret_cast<double*>(shared_ws(c) + p.buf_offset(p.buf_acc, c.chunk_cap, c.train)) + 1;
      const int wpb = 8;
      launch_k(recon_kernel, dim3((unsigned)((c.n + wpb - 1) / wpb)), dim3(wpb * 32), 0, st, 
          c.x, resolve(c, o.r1), c.grad ? resolve(c, o.r2) : nullptr, c.grad ? resolve(c, o.r3) : nullptr, acc,
          o.i0, o.i1, 1, c.n, 1.0f / (float)c.n_total, p.bufs[o.r2.buf].split);
      h->launches++; break;
    }
    case OP_COLSUM: {
      launch_k(colsum_kernel, dim3((unsigned)((o.i0 + 255) / 256)), dim3(256), 0, st, resolve(c, o.r0), resolve(c, o.r1), o.i0, o.i1);
      h->launches++; break;
    }
    case OP_ZERO: {
      long long cnt = o.per_frame_count ? o.count * c.n : o.count;
      if (o.per_frame_count && o.i1 > 0 && o.r0.space == SP_WS && p.bufs[o.r0.buf].split && c.n > 0 &&
          o.i0 % 8 == 0 && o.i1 % 8 == 0 && o.count % 8 == 0) {
        // only the pads: a frame is [hi plane: count bf16][lo plane: count bf16]; the interior [i0, i0 + i1) of each plane is
        // written by the producing GEMM
        const long long per = 2LL * ((o.i0 >> 3) + ((o.count - o.i0 - o.i1) >> 3)), tot = per * c.n;
        if (tot > 0) {
          launch_k(zero_pads_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, reinterpret_cast<uint16_t*>(resolve(c, o.r0)), (int)o.count, o.i0, o.i1, (long long)c.n);
          h->launches++;
        }
        break;
      }
      CUDA_TRY(cudaMemsetAsync(resolve(c, o.r0), 0, (size_t)cnt * sizeof(float), st));
      break;
    }
    default: return fail(NPVC_ERR_ARG, "unknown op kind");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NPVC_ERR_CUDA, "launch " + o.name + ": " + cudaGetErrorString(e));
  return NPVC_OK;
}

int run_phase(Ctx& c, int phase) {
  npvc_handle* h = c.h;
  const std::vector<Op>& ops = h->plan.ops;
  // Backward: the weight gradients are off the critical path (nothing reads the packed gradient before the
  // final unpack, every activation / gradient buffer is written once per pass), so they are forked onto a
  // side stream behind an event and joined at the end of the phase: the L2-bound wgrad GEMMs overlap the
  // HBM-bound Layernorm backward kernels and the tails of the dgrad GEMMs.  Still ordered w.r.t. the
  // caller's stream (fork / join events only); nothing synchronises the device.
  const bool fork = (phase == PH_BWD) && h->overlap_wgrad && !h->profiling;
  // Packing inside a training call: the fused first layer reads only the fp32 pack (the phase's first op), so
  // the speaker table and the bf16 planes are forked onto the side stream and joined in front of the first op
  // that is not the fused first layer (below) - they run beside it instead of in front of it.
  const bool fork_pack = (phase == PH_PACK) && h->pack_defer && h->overlap_wgrad && !h->profiling;
  const cudaStream_t main_st = c.st;
  bool forked = false, first = true;
  if ((fork || fork_pack) && !h->side) {
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess) return fail(NPVC_ERR_CUDA, "cudaStreamCreate failed");
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming);
  }
  for (size_t i = 0; i < ops.size(); i++) {
    const Op& o = ops[i];
    if (o.phase != phase) continue;
    if (!c.grad && (o.kind == OP_UNPACK)) continue;
    npvc_handle::Ev ev{(int)i, nullptr, nullptr, 0, c.n};
    if (h->profiling) {
      cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
      ev.rows = (o.kind == OP_GEMM || o.kind == OP_WGRAD) ? (o.rows_fixed ? o.rows_fixed : c.n * o.A.R) : c.n;
      cudaEventRecord(ev.a, c.st);
    }
    if (h->pack_pending && phase != PH_PACK && o.fuse != FUSE_E0_FWD) {   // first reader of the forked packs
      cudaStreamWaitEvent(main_st, h->ev_pack, 0); h->pack_pending = false;
    }
    const bool on_side = (fork && o.kind == OP_WGRAD) || (fork_pack && !first);
    if (on_side && !(fork_pack && forked)) {         // everything issued so far on the caller's stream happens-before this op
      cudaEventRecord(h->ev_fork, main_st); cudaStreamWaitEvent(h->side, h->ev_fork, 0);
    }
    if (on_side) { c.st = h->side; forked = true; }
    first = false;
    if (h->nvtx) nvtxRangePushA(o.name.c_str());
    int rc;
    if (o.fuse == FUSE_E0_FWD) {                     // this op and the next one as one kernel (plan.h, Op::fuse)
      rc = launch_e0_fwd(c, o, ops[i + 1]); i++;
    } else if (o.fuse == FUSE_E0_BWD) {
      rc = launch_e0_bwd(c, o, ops[i + 1]); i++;
    } else if (o.fuse == FUSE_SPK_BWD) {
      rc = launch_speaker_bwd(c, o, ops[i + 1], ops[i + 2]); i += 2;
    } else if (o.fuse == FUSE_LN_FWD) {              // Layernorm in the GEMM's epilogue when the tiling allows it
      bool fused = false;
      rc = run_op(c, o, (int)i, &ops[i + 1], &fused);
      if (fused) i++;
    } else rc = run_op(c, o, (int)i);
    if (rc == NPVC_OK && (o.fuse == FUSE_E0_FWD || o.fuse == FUSE_E0_BWD || o.fuse == FUSE_SPK_BWD)) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) rc = fail(NPVC_ERR_CUDA, "launch " + o.name + " (fused): " + cudaGetErrorString(e)); }
    if (h->nvtx) nvtxRangePop();
    c.st = main_st;
    if (h->profiling) { cudaEventRecord(ev.b, c.st); h->events.push_back(ev); }
    if (rc) return rc;
  }
  if (forked && fork_pack) { cudaEventRecord(h->ev_pack, h->side); h->pack_pending = true; }
  else if (forked) { cudaEventRecord(h->ev_join, h->side); cudaStreamWaitEvent(main_st, h->ev_join, 0); }
  return NPVC_OK;
}

void free_tables(npvc_handle* h) {
  cudaFree(h->d_pack_list); h->d_pack_list = nullptr;
  cudaFree(h->d_pack_src); cudaFree(h->d_pack16_src); cudaFree(h->d_unpack_ptr); cudaFree(h->d_unpack_idx); cudaFree(h->d_heavy);   // (cudaFree(nullptr) is a no-op)
  h->d_pack_src = h->d_pack16_src = h->d_unpack_ptr = h->d_unpack_idx = h->d_heavy = nullptr;
  h->tables_on_device = false;
}

int upload_tables(npvc_handle* h);
int ensure_tables(npvc_handle* h) {
  if (h->tables_on_device) return NPVC_OK;
  const int rc = upload_tables(h);
  if (rc) free_tables(h);            // a failed upload leaves nothing behind: the next call starts over
  return rc;
}
int upload_tables(npvc_handle* h) {
  int dev = 0, cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) return fail(NPVC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
  const Plan& p = h->plan;
  if (h->use_umma) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t ee = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ee != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(NPVC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver (needed by the tcgen05 path)");
    h->encode = (PFN_tmapEncodeTiled)fn;
  }
  CUDA_TRY(cudaMalloc(&h->d_pack_src, (p.pack_src.size() + 1) * 4));
  CUDA_TRY(cudaMalloc(&h->d_pack16_src, (p.pack16_src.size() + 1) * 4));
  CUDA_TRY(cudaMemcpy(h->d_pack16_src, p.pack16_src.data(), p.pack16_src.size() * 4, cudaMemcpyHostToDevice));
  if (!p.pack_list.empty()) {
    CUDA_TRY(cudaMalloc(&h->d_pack_list, p.pack_list.size() * 4));
    CUDA_TRY(cudaMemcpy(h->d_pack_list, p.pack_list.data(), p.pack_list.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMalloc(&h->d_unpack_ptr, p.unpack_ptr.size() * 4));
  CUDA_TRY(cudaMalloc(&h->d_unpack_idx, (p.unpack_idx.size() + 1) * 4));
  CUDA_TRY(cudaMemcpy(h->d_pack_src, p.pack_src.data(), p.pack_src.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_unpack_ptr, p.unpack_ptr.data(), p.unpack_ptr.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_unpack_idx, p.unpack_idx.data(), p.unpack_idx.size() * 4, cudaMemcpyHostToDevice));
  {
    std::vector<int32_t> heavy;
    for (int64_t t = 0; t < p.n_params; t++) if (p.unpack_ptr[t + 1] - p.unpack_ptr[t] >= UNPACK_HEAVY) heavy.push_back((int32_t)t);
    h->n_heavy = (int)heavy.size();
    CUDA_TRY(cudaMalloc(&h->d_heavy, (heavy.size() + 1) * 4));
    if (!heavy.empty()) CUDA_TRY(cudaMemcpy(h->d_heavy, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice));
  }
  h->tables_on_device = true;
  return NPVC_OK;
}

int check_ws(npvc_handle* h, int64_t n, bool train, int64_t ws_bytes, const void* ws) {
  if (!h) return fail(NPVC_ERR_ARG, "null handle");
  if (!ws) return fail(NPVC_ERR_ARG, "null workspace");
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(NPVC_ERR_ARG, "workspace must be 256-byte aligned");
  int64_t need = npvc_workspace_bytes(h, n, train ? 1 : 0);
  if (ws_bytes < need) {
    char m[160]; snprintf(m, sizeof m, "workspace too small: %lld < %lld bytes", (long long)ws_bytes, (long long)need);
    return fail(NPVC_ERR_WORKSPACE, m);
  }
  return NPVC_OK;
}

}  // namespace

// --------------------------------------------------------------------------------------------
extern "C" {

const char* npvc_last_error(void) { return g_err.c_str(); }
const char* npvc_version(void) { return "npvc_b200 0.1 (sm_100a)"; }

int npvc_create(const npvc_arch* arch, int64_t max_chunk, npvc_handle** out) {
  if (!arch || !out) return fail(NPVC_ERR_ARG, "null argument");
  npvc_handle* h = new npvc_handle();
  const char* eu = getenv("NPVC_UMMA");            // "0" = CUDA-core GEMMs only (debug / A-B comparisons)
  h->use_umma = !(eu && eu[0] == '0');
  if (const char* tp = getenv("NPVC_UMMA_TAP")) h->umma_tap = atoi(tp);
  if (const char* ov = getenv("NPVC_OVERLAP")) h->overlap_wgrad = atoi(ov);
  { const char* pd = getenv("NPVC_PDL"); g_pdl = (pd && !atoi(pd)) ? 0 : 1; }      // (process-wide: the launch helper has no handle)
  if (const char* nv = getenv("NPVC_NVTX")) h->nvtx = atoi(nv);
  if (const char* mg = getenv("NPVC_UMMA_MERGE")) h->umma_merge = atoi(mg);
  if (const char* br = getenv("NPVC_UMMA_BRES")) h->umma_bres = atoi(br);
  if (const char* du = getenv("NPVC_UMMA_DUAL")) h->umma_dual = atoi(du);
  h->prefetch_min = PREFETCH_MIN_FRAMES; if (const char* pm = getenv("NPVC_PREFETCH_MIN")) h->prefetch_min = atoll(pm);
  h->prefetch_e0_min = PREFETCH_MIN_FRAMES; if (const char* pm = getenv("NPVC_PREFETCH_E0_MIN")) h->prefetch_e0_min = atoll(pm);
  if (const char* fl = getenv("NPVC_FUSE_LN_TRAIN")) h->fuse_ln_train = atoi(fl);
  if (const char* pr = getenv("NPVC_PAIR")) h->umma_pair = atoi(pr);
  if (const char* wp = getenv("NPVC_WGRAD_PAIR")) h->wgrad_pair = atoi(wp);
  if (const char* po = getenv("NPVC_PAIR_OPS")) { h->pair_ops = po; if (!h->pair_ops.empty() && !h->umma_pair) h->umma_pair = 1; }
  const char* ea = getenv("NPVC_UMMA_OPS");
  if (ea) h->umma_allow = ea;
  const char* fz = getenv("NPVC_FUSE");            // "0" = every plan op as its own kernel (A/B comparisons, fallback-path tests)
  std::string err = build_plan(*arch, h->plan, h->use_umma, !(fz && fz[0] == '0'));
  if (!err.empty()) { delete h; return fail(NPVC_ERR_ARG, "unsupported architecture: " + err); }
  if (max_chunk > 0) h->max_chunk = max_chunk;
  else if (const char* mc = getenv("NPVC_MAX_CHUNK")) { long v = atol(mc); if (v > 0) h->max_chunk = v; }
  // frames per internal pass: row counts, tile counts and TMA coordinates of a pass are 32-bit
  if (h->max_chunk > ((int64_t)1 << 20)) { delete h; return fail(NPVC_ERR_ARG, "max_chunk > 1048576 frames per pass is not supported (larger calls are chunked by the library)"); }
  *out = h;
  return NPVC_OK;
}

void npvc_destroy(npvc_handle* h) {
  if (!h) return;
  if (h->side) { cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); cudaEventDestroy(h->ev_pack); }
  if (h->tables_on_device) free_tables(h);
  delete h;
}

int64_t npvc_param_count(const npvc_handle* h) { return h ? h->plan.n_params : 0; }
int32_t npvc_param_tensors(const npvc_handle* h) { return h ? (int32_t)h->plan.params.size() : 0; }
int npvc_param_table(const npvc_handle* h, npvc_param_desc* out, int32_t max) {
  if (!h || !out) return fail(NPVC_ERR_ARG, "null argument");
  int n = (int)h->plan.params.size(); if (n > max) n = max;
  for (int i = 0; i < n; i++) {
    const Param& q = h->plan.params[i];
    memset(&out[i], 0, sizeof(npvc_param_desc));
    strncpy(out[i].name, q.name.c_str(), sizeof(out[i].name) - 1);
    out[i].offset = q.off; out[i].size = q.size; out[i].rank = q.rank;
    for (int d = 0; d < 4; d++) out[i].shape[d] = q.shape[d];
    out[i].fan_in = q.fan_in; out[i].fan_out = q.fan_out; out[i].init = q.init;
  }
  return NPVC_OK;
}

int64_t npvc_workspace_bytes(const npvc_handle* h, int64_t n, int32_t train) {
  if (!h || n < 0) return -1;
  int64_t chunk = n < h->max_chunk ? n : h->max_chunk;
  if (chunk < 1) chunk = 1;
  return h->plan.ws_floats(chunk, train != 0) * 4;
}

const char* npvc_plan_json(const npvc_handle* h) { return h ? h->plan.json.c_str() : ""; }

int64_t npvc_plan_table(const npvc_handle* h, const char* name, int32_t* out, int64_t max) {
  if (!h || !name) return -1;
  const std::vector<int32_t>* v = nullptr;
  if (!strcmp(name, "pack_src")) v = &h->plan.pack_src;
  else if (!strcmp(name, "unpack_ptr")) v = &h->plan.unpack_ptr;
  else if (!strcmp(name, "unpack_idx")) v = &h->plan.unpack_idx;
  else if (!strcmp(name, "pack16_src")) v = &h->plan.pack16_src;
  else if (!strcmp(name, "pack_list")) v = &h->plan.pack_list;
  if (!v) return -1;
  if (out) { int64_t n = (int64_t)v->size() < max ? (int64_t)v->size() : max; memcpy(out, v->data(), n * 4); }
  return (int64_t)v->size();
}

int64_t npvc_launch_count(const npvc_handle* h) { return h ? h->launches : 0; }

int npvc_profile_enable(npvc_handle* h, int32_t enable) {
  if (!h) return fail(NPVC_ERR_ARG, "null handle");
  h->profiling = enable != 0;
  return NPVC_OK;
}

const char* npvc_profile_json(npvc_handle* h) {
  if (!h) return "[]";
  const std::vector<Op>& ops = h->plan.ops;
  std::vector<double> ms(ops.size(), 0.0); std::vector<long long> calls(ops.size(), 0), rows(ops.size(), 0), frames(ops.size(), 0);
  for (auto& e : h->events) {
    cudaEventSynchronize(e.b);
    float t = 0.f; cudaEventElapsedTime(&t, e.a, e.b);
    ms[e.op] += t; calls[e.op]++; rows[e.op] += e.rows; frames[e.op] += e.frames;
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  h->events.clear();
  std::string js = "["; bool first = true; char buf[512];
  for (size_t i = 0; i < ops.size(); i++) {
    if (!calls[i]) continue;
    // algorithmic (compulsory) bytes of the op, summed over its calls: every operand buffer once
    const Op& o = ops[i]; const Plan& p = h->plan;
    auto frame_floats = [&](const Ref& r) -> double { return r.space == SP_WS ? (double)p.bufs[r.buf].per_frame : 0.0; };
    double per_frame = 0.0, fixed = 0.0;
    if (o.kind == OP_GEMM) { per_frame = frame_floats(o.A.ref) + (o.A.ref.space == SP_USER ? p.arch.in_h : 0) + frame_floats(o.C.ref); fixed = (double)o.K * o.N; }
    else if (o.kind == OP_WGRAD) { per_frame = frame_floats(o.A.ref) + (o.A.ref.space == SP_USER ? p.arch.in_h : 0) + frame_floats(o.C.ref); fixed = (double)o.K * o.N; }
    else if (o.kind == OP_LN_FWD) per_frame = o.L + o.out_flen;
    else if (o.kind == OP_LN_BWD) per_frame = 2.0 * o.L + o.out_flen;
    // fused first layer: x in, raw conv output + activation planes out (training) / dy, c and x in, nothing per frame out
    if (o.fuse == FUSE_E0_FWD) { const Op& ln = ops[i + 1]; per_frame = p.arch.in_h + (h->last_train ? ln.L : 0) + ln.out_flen; fixed = 0.0; }
    if (o.fuse == FUSE_E0_BWD) per_frame = 2.0 * o.L + p.arch.in_h;
    const double bytes = 4.0 * (per_frame * (double)frames[i] + fixed * (double)calls[i]);
    const bool tensor = o.umma && umma_allowed(h, o);
    snprintf(buf, sizeof buf, "%s{\"name\":\"%s\",\"kind\":%d,\"calls\":%lld,\"ms\":%.6f,\"rows\":%lld,\"K\":%d,\"N\":%d,\"tensor\":%d,\"bytes\":%.0f}",
             first ? "" : ",", o.name.c_str(), o.kind, calls[i], ms[i], rows[i], o.K, o.N, tensor ? 1 : 0, bytes);
    js += buf; first = false;
  }
  js += "]";
  h->profile_json = js;
  return h->profile_json.c_str();
}

int npvc_pack_weights(npvc_handle* h, const float* d_theta, void* d_ws, int64_t ws_bytes, void* stream) {
  int rc = check_ws(h, 1, false, ws_bytes, d_ws); if (rc) return rc;
  if (!d_theta) return fail(NPVC_ERR_ARG, "null theta");
  rc = ensure_tables(h); if (rc) return rc;
  Ctx c{h, (float*)d_ws, 1, false, d_theta, nullptr, nullptr, nullptr, nullptr, 0, 1, (cudaStream_t)stream};
  return run_phase(c, PH_PACK);
}

int npvc_encode(npvc_handle* h, const float* d_theta, const float* d_x, int64_t n, float* d_mu, float* d_lv,
                void* d_ws, int64_t ws_bytes, void* stream) {
  int rc = check_ws(h, n, false, ws_bytes, d_ws); if (rc) return rc;
  if (!d_theta || !d_x || n < 0) return fail(NPVC_ERR_ARG, "bad argument");
  rc = ensure_tables(h); if (rc) return rc;
  const Plan& p = h->plan; cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = n < h->max_chunk ? (n > 0 ? n : 1) : h->max_chunk; const int z = p.arch.z_dim;
  for (int64_t c0 = 0; c0 < n; c0 += cap) {
    int64_t m = n - c0 < cap ? n - c0 : cap;
    Ctx c{h, (float*)d_ws, cap, false, d_theta, nullptr, d_x + c0 * p.arch.in_h, nullptr, nullptr, m, n, st};
    rc = run_phase(c, PH_ENC); if (rc) return rc;
    rc = run_phase(c, PH_SAMPLE); if (rc) return rc;
    if (d_mu) CUDA_TRY(cudaMemcpyAsync(d_mu + c0 * z, c.ws + p.buf_offset(p.buf_mu, cap, false), (size_t)m * z * 4, cudaMemcpyDeviceToDevice, st));
    if (d_lv) CUDA_TRY(cudaMemcpyAsync(d_lv + c0 * z, c.ws + p.buf_offset(p.buf_lv, cap, false), (size_t)m * z * 4, cudaMemcpyDeviceToDevice, st));
  }
  h->last_chunk = cap; h->last_train = false;
  return NPVC_OK;
}

int npvc_sample(npvc_handle* h, const float* d_mu, const float* d_lv, const float* d_eps, int64_t n, float* d_z, void* stream) {
  if (!h || !d_mu || !d_lv || !d_eps || !d_z) return fail(NPVC_ERR_ARG, "null argument");
  int rc = ensure_tables(h); if (rc) return rc;
  long long tot = n * h->plan.arch.z_dim;
  if (tot > 0) {
    launch_k(sample_only_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_mu, d_lv, d_eps, d_z, tot);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int npvc_decode(npvc_handle* h, const float* d_theta, const float* d_z, const int64_t* d_y, int64_t n, float* d_xh,
                void* d_ws, int64_t ws_bytes, void* stream) {
  int rc = check_ws(h, n, false, ws_bytes, d_ws); if (rc) return rc;
  if (!d_theta || !d_z || !d_y || !d_xh || n < 0) return fail(NPVC_ERR_ARG, "bad argument");
  rc = ensure_tables(h); if (rc) return rc;
  const Plan& p = h->plan; cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = n < h->max_chunk ? (n > 0 ? n : 1) : h->max_chunk; const int z = p.arch.z_dim, H = p.arch.in_h;
  for (int64_t c0 = 0; c0 < n; c0 += cap) {
    int64_t m = n - c0 < cap ? n - c0 : cap;
    Ctx c{h, (float*)d_ws, cap, false, d_theta, nullptr, nullptr, d_y + c0, nullptr, m, n, st};
    CUDA_TRY(cudaMemcpyAsync(c.ws + p.buf_offset(p.buf_z, cap, false), d_z + c0 * z, (size_t)m * z * 4, cudaMemcpyDeviceToDevice, st));
    rc = run_phase(c, PH_DEC); if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(d_xh + c0 * H, (size_t)H * 4, c.ws + p.buf_offset(p.buf_xh, cap, false), (size_t)p.bufs[p.buf_xh].per_frame * 4, (size_t)H * 4, (size_t)m,
                               cudaMemcpyDeviceToDevice, st));      // (workspace rows at the padded pitch)
  }
  h->last_chunk = cap; h->last_train = false;
  return NPVC_OK;
}

static int loss_core(npvc_handle* h, const float* d_theta, const float* d_x, const int64_t* d_y, const float* d_eps,
                     StepState* d_state, int64_t frame_offset,
                     int64_t n, float* d_z, float* d_mu, float* d_lv, float* d_xh, float* d_losses, float* d_grad,
                     int32_t repack, void* d_ws, int64_t ws_bytes, void* stream) {
  int rc = check_ws(h, n, true, ws_bytes, d_ws); if (rc) return rc;
  if (!d_theta || !d_x || !d_y || (!d_eps && !d_state) || n < 1) return fail(NPVC_ERR_ARG, "bad argument");
  rc = ensure_tables(h); if (rc) return rc;
  const Plan& p = h->plan; cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = n < h->max_chunk ? n : h->max_chunk; const int z = p.arch.z_dim, H = p.arch.in_h;
  float* ws = (float*)d_ws;
  CUDA_TRY(cudaMemsetAsync(ws + p.buf_offset(p.buf_acc, cap, true), 0, 8 * 4, st));
  if (d_grad) {
    CUDA_TRY(cudaMemsetAsync(d_grad, 0, (size_t)p.n_params * 4, st));
    CUDA_TRY(cudaMemsetAsync(ws + p.buf_offset(p.buf_adw, cap, true), 0, (size_t)p.arena_dw * 4, st));
  }
  if (repack) {
    Ctx c{h, ws, cap, true, d_theta, nullptr, nullptr, nullptr, nullptr, 0, n, st};
    h->pack_defer = true; rc = run_phase(c, PH_PACK); h->pack_defer = false;
    if (rc) return rc;
  }
  for (int64_t c0 = 0; c0 < n; c0 += cap) {
    int64_t m = n - c0 < cap ? n - c0 : cap;
    float* wset = ws; const cudaStream_t cst = st;
    Ctx c{h, wset, cap, true, d_theta, d_grad, d_x + c0 * H, d_y + c0, d_eps ? d_eps + c0 * z : nullptr, m, n, cst};
    if (!d_eps) { c.state = d_state; c.frame0 = frame_offset + c0; }
    for (int ph : {PH_ENC, PH_SAMPLE, PH_DEC, PH_LOSS}) { rc = run_phase(c, ph); if (rc) return rc; }
    if (d_grad) { rc = run_phase(c, PH_BWD); if (rc) return rc; }
    struct { float* dst; int buf; int w; } outs[4] = {{d_z, p.buf_z, z}, {d_mu, p.buf_mu, z}, {d_lv, p.buf_lv, z}, {d_xh, p.buf_xh, H}};
    for (auto& o : outs)      // (2-D: the workspace rows of xh are at the padded pitch)
      if (o.dst) CUDA_TRY(cudaMemcpy2DAsync(o.dst + c0 * o.w, (size_t)o.w * 4, wset + p.buf_offset(o.buf, cap, true), (size_t)p.bufs[o.buf].per_frame * 4, (size_t)o.w * 4, (size_t)m,
                                            cudaMemcpyDeviceToDevice, cst));
  }
  if (d_grad) {
    Ctx c{h, ws, cap, true, d_theta, d_grad, nullptr, nullptr, nullptr, 0, n, st};
    rc = run_phase(c, PH_FINAL); if (rc) return rc;
  }
  if (d_losses || d_state) {
    launch_k(finalize_losses_kernel, dim3(1), dim3(32), 0, st, reinterpret_cast<const double*>(ws + p.buf_offset(p.buf_acc, cap, true)), d_losses, 1.0 / (double)n,
                                             d_state, d_grad ? 1 : 0);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  h->last_chunk = cap; h->last_train = true;
  return NPVC_OK;
}

int npvc_loss_fwd_bwd(npvc_handle* h, const float* d_theta, const float* d_x, const int64_t* d_y, const float* d_eps,
                      int64_t n, float* d_z, float* d_mu, float* d_lv, float* d_xh, float* d_losses, float* d_grad,
                      int32_t repack, void* d_ws, int64_t ws_bytes, void* stream) {
  if (!d_eps) return fail(NPVC_ERR_ARG, "bad argument (eps required: npvc_train_fwd_bwd draws it in-kernel)");
  return loss_core(h, d_theta, d_x, d_y, d_eps, nullptr, 0, n, d_z, d_mu, d_lv, d_xh, d_losses, d_grad, repack, d_ws, ws_bytes, stream);
}

int npvc_train_fwd_bwd(npvc_handle* h, const float* d_theta, const float* d_x, const int64_t* d_y, npvc_step_state* d_state,
                       int64_t frame_offset, int64_t n, float* d_z, float* d_mu, float* d_lv, float* d_xh, float* d_losses,
                       float* d_grad, int32_t repack, void* d_ws, int64_t ws_bytes, void* stream) {
  if (!d_state || frame_offset < 0) return fail(NPVC_ERR_ARG, "bad argument (device step state required)");
  return loss_core(h, d_theta, d_x, d_y, nullptr, reinterpret_cast<StepState*>(d_state), frame_offset, n, d_z, d_mu, d_lv, d_xh, d_losses,
                   d_grad, repack, d_ws, ws_bytes, stream);
}

int npvc_normal_draw(npvc_handle* h, const npvc_step_state* d_state, int64_t frame_offset, int64_t n, float* d_eps, void* stream) {
  if (!h || !d_state || !d_eps || n < 0 || frame_offset < 0) return fail(NPVC_ERR_ARG, "bad argument");
  int rc = ensure_tables(h); if (rc) return rc;
  const int z = h->plan.arch.z_dim; const long long tot = n * z;
  if (tot > 0) {
    launch_k(philox_normal_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const StepState*>(d_state), frame_offset, d_eps, z, n);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int npvc_adam_step(npvc_handle* h, float* d_theta, const float* d_grad, float* d_m, float* d_v, int64_t n_params,
                   int64_t step, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!h || !d_theta || !d_grad || !d_m || !d_v || step < 1) return fail(NPVC_ERR_ARG, "bad argument");
  int rc = ensure_tables(h); if (rc) return rc;
  double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) / (1.0 - std::pow((double)beta1, (double)step));
  launch_k(adam_kernel, dim3((unsigned)((n_params + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_theta, d_grad, d_m, d_v, n_params, (float)lr_t, beta1, beta2, eps, grad_scale, nullptr);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int npvc_adam_step_dev(npvc_handle* h, float* d_theta, const float* d_grad, float* d_m, float* d_v, int64_t n_params,
                       const npvc_step_state* d_state, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!h || !d_theta || !d_grad || !d_m || !d_v || !d_state) return fail(NPVC_ERR_ARG, "bad argument");
  int rc = ensure_tables(h); if (rc) return rc;
  launch_k(adam_kernel, dim3((unsigned)((n_params + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_theta, d_grad, d_m, d_v, n_params, lr, beta1, beta2, eps, grad_scale,
                                                                                    reinterpret_cast<const StepState*>(d_state));
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int npvc_tanhize_forward(npvc_handle* h, const float* d_x, const float* d_xmin, const float* d_xmax, int64_t n, int32_t dim, float* d_out, void* stream) {
  if (!h || !d_x || !d_xmin || !d_xmax || !d_out) return fail(NPVC_ERR_ARG, "null argument");
  int rc = ensure_tables(h); if (rc) return rc;
  long long tot = n * dim;
  if (tot > 0) { launch_k(tanhize_fwd_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_x, d_xmin, d_xmax, d_out, n, dim); h->launches++; }
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}
int npvc_tanhize_backward(npvc_handle* h, const float* d_x, const float* d_xmin, const float* d_xmax, int64_t n, int32_t dim, float* d_out, void* stream) {
  if (!h || !d_x || !d_xmin || !d_xmax || !d_out) return fail(NPVC_ERR_ARG, "null argument");
  int rc = ensure_tables(h); if (rc) return rc;
  long long tot = n * dim;
  if (tot > 0) { launch_k(tanhize_bwd_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_x, d_xmin, d_xmax, d_out, n, dim); h->launches++; }
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int npvc_unpack_records(npvc_handle* h, const float* d_records, int64_t n, int32_t rec_floats, int32_t sp_dim,
                        const float* d_xmin, const float* d_xmax, float* d_x, int64_t* d_y, void* stream) {
  if (!h || !d_records || !d_x || !d_y || sp_dim > rec_floats) return fail(NPVC_ERR_ARG, "bad argument");
  int rc = ensure_tables(h); if (rc) return rc;
  long long tot = n * sp_dim;
  if (tot > 0) {
    launch_k(unpack_records_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_records, n, rec_floats, sp_dim, d_xmin, d_xmax, d_x, reinterpret_cast<long long*>(d_y));
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return NPVC_OK;
}

int64_t npvc_debug_buffer(npvc_handle* h, const char* name, const void* d_ws, float* d_out, int64_t n, void* stream) {
  if (!h || !name || !d_ws) return -1;
  const Plan& p = h->plan;
  for (int i = 0; i < (int)p.bufs.size(); i++) {
    if (p.bufs[i].name != name) continue;
    if (p.bufs[i].train_only && !h->last_train) return -2;
    int64_t per = p.bufs[i].per_frame, cnt = p.bufs[i].fixed + per * (n < h->last_chunk ? n : h->last_chunk);
    if (d_out) cudaMemcpyAsync(d_out, (const float*)d_ws + p.buf_offset(i, h->last_chunk, h->last_train), (size_t)cnt * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return per ? per : p.bufs[i].fixed;
  }
  if (!strcmp(name, "arena_w")) {
    if (d_out) cudaMemcpyAsync(d_out, d_ws, (size_t)p.arena_w * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return p.arena_w;
  }
  return -1;
}

}  // extern "C"
