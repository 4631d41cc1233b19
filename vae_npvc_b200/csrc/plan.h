// Launch plan of the ConvVAE hot path: buffers, strided row views, operand packs and the op list.
// Pure host C++ (no CUDA): built once from the architecture in npvc_create(), interpreted by
// engine.cu on the GPU and by tests/plan_interp.py in numpy on the CPU.
//
// Every layer of the reference graph (model/vae.py:72-103) is one of three GEMM forms over
// channels-last, per-frame zero-padded activations (frames stay in the batch dimension, F5/F6):
//   (F) C[rows,N]  = A_view[rows,K] . B[K,N]          rows = (frame, position)
//   (W) dB[K,N]   += A_view[rows,K]^T . dC_view[rows,N]
// where A_view row (frame f, j) is the CONTIGUOUS window of K floats starting at
// f*frame_stride + j*row_stride + off of a padded channels-last buffer (an im2col that costs
// nothing: strided conv windows overlap in memory).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/npvc_b200.h"

namespace npvc {

enum Space { SP_NONE = 0, SP_WS = 1, SP_THETA = 2, SP_GRAD = 3, SP_AW = 4, SP_ADW = 5, SP_USER = 6 };
enum UserSlot { U_X = 0, U_Y = 1, U_EPS = 2 };
enum Phase { PH_PACK = 0, PH_ENC = 1, PH_SAMPLE = 2, PH_DEC = 3, PH_LOSS = 4, PH_BWD = 5, PH_FINAL = 6 };
enum OpKind {
  OP_GEMM = 0, OP_WGRAD = 1, OP_LN_FWD = 2, OP_LN_BWD = 3, OP_SAMPLE = 4, OP_SAMPLE_BWD = 5,
  OP_RECON = 6, OP_COLSUM = 8, OP_ZERO = 9, OP_PACK = 10, OP_UNPACK = 11,
  OP_PACK16 = 12,   // theta -> bf16 hi / lo operand packs (tensor path)
  OP_ZCAT = 13      // zs[f] = [z[f] (i0) | one-hot(y[f]) (i1)]  (r0 -> r1; fp32 or split planes)
};

// Op::fuse: the op and the NEXT op of the list are executed by one fused kernel (engine.cu); the plan keeps both
// ops (they remain the definition of the arithmetic; tests/plan_interp.py runs them one by one).
//   FUSE_E0_FWD  conv of the first encoder layer (one input channel, <= 8 taps) + its Layernorm / lrelu
//   FUSE_E0_BWD  Layernorm backward of the first encoder layer + its weight gradient: the gradient w.r.t. the
//                conv output is consumed in registers, its buffer is never materialised (Buf::elide)
//   FUSE_LN_FWD  a conv / transposed-conv GEMM + the Layernorm / lrelu that follows it: done in the GEMM kernel's
//                epilogue when a tile holds whole frames (decided per launch; otherwise the two ops run separately)
//   FUSE_SPK_BWD the three once-per-call ops of the speaker branch's backward (weight gradient of the per-speaker FC,
//                embedding gradient, bias column sums): this op and the next TWO as one kernel
enum Fuse { FUSE_NONE = 0, FUSE_E0_FWD = 1, FUSE_E0_BWD = 2, FUSE_LN_FWD = 3, FUSE_SPK_BWD = 4 };

struct Ref {
  int space = SP_NONE;
  int buf = -1;        // SP_WS: buffer index; SP_USER: slot
  int64_t off = 0;     // float offset (theta / grad / arena spaces)
};

struct View {
  Ref ref;
  int R = 1;           // rows per frame
  int64_t fs = 0;      // frame stride (floats)
  int rs = 0;          // row stride (floats)
  int off = 0;         // in-frame offset of column 0 of row 0 (may be negative)
  int flen = 0;        // valid in-frame range [0, flen) (used when pred)
  int pred = 0;        // predicate every element on the in-frame range
  int split = 0;       // the buffer holds bf16 hi / lo planes (see Buf::split); fs == the buffer's per_frame
};

struct Buf {
  std::string name;
  int64_t per_frame = 0;   // floats per frame
  int64_t fixed = 0;       // floats independent of n
  int train_only = 0;
  // split != 0: a frame's `per_frame` floats are stored as two bf16 planes [hi: per_frame][lo: per_frame]
  // with value = hi + lo (hi = bf16(v), lo = bf16(v - hi)): the tensor-core operand format.  The
  // tensor path multiplies such operands as hi.hi + hi.lo + lo.hi in fp32 (bf16x3).
  int split = 0;
  int elide = 0;           // no storage: every producer / consumer of the buffer runs inside a fused kernel (Op::fuse)
  int alias = -1;          // >= 0: the buffer has no storage of its own and lives at the start of buffer `alias` (its lifetime does
                           // not overlap the other tenants': the da_l gradients, each written by one dgrad GEMM and consumed by the
                           // Layernorm backward that follows it on the same stream)
};

struct Op {
  int kind = 0, phase = 0;
  int fuse = 0;          // Fuse
  std::string name;
  // GEMM / WGRAD
  View A, C;             // WGRAD: C is the dC view
  int K = 0, N = 0;
  Ref B; int ldb = 0;    // GEMM: B operand; WGRAD: output dB
  Ref bias[3]; int bias_mod = 1;
  int64_t rows_fixed = 0;        // >0: row count independent of n (A is not per-frame)
  int a_scalar = 0;              // A needs the scalar (unaligned / predicated) loader
  // tcgen05 path: OP_GEMM reads B as K-major [N, kpad] bf16 hi / lo packs (bf16-element offsets into the
  // arena16 region); OP_WGRAD reads both operands straight from the split activation / gradient views
  int umma = 0; int64_t bu_hi = 0, bu_lo = 0; int kpad = 0;
  // conv-shaped A view: K = tap_T taps x tap_C channels, consecutive rows tap_s positions apart (0: not a window view)
  int tap_T = 0, tap_C = 0, tap_s = 0;
  // LN_FWD / LN_BWD
  Ref in, xhat, aout, rstd, gamma, beta, dgamma, dbeta, dbias;
  int L = 0, Cn = 0, out_flen = 0, out_off = 0;
  // misc
  Ref r0, r1, r2, r3;
  int64_t count = 0; int per_frame_count = 0;
  int i0 = 0, i1 = 0;
};

struct Param {
  std::string name;
  int64_t off = 0, size = 0;
  int rank = 0; int shape[4] = {0, 0, 0, 0};
  int fan_in = 0, fan_out = 0, init = 0;
};

struct Plan {
  npvc_arch arch;
  std::vector<Param> params;
  int64_t n_params = 0;
  std::vector<Buf> bufs;
  std::vector<Op> ops;
  int64_t arena_w = 0;       // floats: packed operand matrices (fwd part first)
  int64_t arena_dw = 0;      // floats: packed weight gradients (mirror of the fwd part; ws buffer "arena_dw")
  int64_t aw16_off = 0;      // float offset of the bf16 pack region inside arena_w
  int64_t aw16_count = 0;    // bf16 elements in it
  std::vector<int32_t> pack_src;     // [aw16_off]  theta index or -1 (fp32 packs)
  std::vector<int32_t> pack16_src;   // [aw16_count / 2] theta index | mode << 29, or -1: entry i fills hi pack i and lo pack i + aw16_count / 2
  std::vector<int32_t> pack_list;    // tensor path: the positions of [0, aw16_off) that some op reads as fp32 (the rest is not packed)
  std::vector<int32_t> unpack_ptr;   // [n_params+1] CSR over theta
  std::vector<int32_t> unpack_idx;   // positions in arena_dw
  int buf_z = -1, buf_mu = -1, buf_lv = -1, buf_xh = -1, buf_acc = -1, buf_hz = -1, buf_adw = -1;
  int out_dim = 0;           // 513
  std::string json;

  int64_t ws_floats(int64_t chunk, bool train) const;
  int64_t buf_offset(int b, int64_t chunk, bool train) const;   // float offset inside ws
};

// pack16_src entries: -1 = zero; else source index | flags: the bf16 hi and the bf16 of the residual of theta[index]
// (or, with PACK16_FROM_ARENA, of the fp32 pack arena_w[index])
constexpr int32_t PACK16_FROM_ARENA = 1 << 29;
constexpr int32_t PACK_INDEX_MASK = (1 << 29) - 1;

// Returns empty string on success, else an error message.  use_umma: route GEMM-shaped ops to the
// tcgen05 kernels: their A operands become split (bf16 hi / lo) buffers, their B operands bf16 packs.
// fuse: mark the op pairs the engine runs as fused kernels (Op::fuse) and elide the buffers that only lived between them.
std::string build_plan(const npvc_arch& a, Plan& p, bool use_umma, bool fuse = true);

// Threads per frame of the register-resident Layernorm kernels, units of 8 elements (0: not applicable)
int ln_group(int L, int Cn, int out_off, int out_flen);
// Threads per frame of the fused first-layer backward kernel, units of 4 elements (0: not applicable)
int e0_bwd_group(int L, int Co);

// Rows of one frame per tensor-core tile (0: the view cannot be tiled); see plan.cpp.
int umma_row_tile(int R);

}  // namespace npvc
