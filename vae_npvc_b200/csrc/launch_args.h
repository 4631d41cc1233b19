// Plain (host + device) argument structs of the kernels: no device code, so host-only tools (the recording CUDA
// runtime stub of tests/cuda_stub) can decode what engine.cu launches.
#pragma once
#include <stdint.h>

namespace npvc {

// Strided row view (see plan.h): element (row, k) lives at
//   p + (row / R) * fs + (row % R) * rs + off + k,  valid (when pred) iff 0 <= (row%R)*rs+off+k < flen
// split != 0 (plan.h, Buf::split): the buffer holds bf16 hi / lo planes per frame -- element e of frame f
// is  hi[f*2*fs + e] + lo[f*2*fs + fs + e]  (bf16 units from p), the tensor-core operand format.
struct DView {
  float* p; long long fs; int R, rs, off, flen, pred, split;
};

// Tiling of a view's rows into <= 128-row tiles of whole (frame, row-group) boxes:
//   row-in-frame j = a * Rb + b  (b < Rb, a < Ra);  a tile = FB frames x Ab row-groups x Rb rows
//   (FB > 1 only when Ra == 1).  Local row r = (fl * Ab + al) * Rb + b.
struct RowTiling {
  int Rb, Ra, Ab, FB, TA;     // TA = ceil(Ra / Ab) tiles per frame block
  int RbH;                    // accumulator rows per row-group: Rb, or Rb + halo rows in tap mode (halo rows are discarded)
  int rows_tile;              // RbH * Ab * FB
  int frames, m_tiles;        // m_tiles = ceil(frames / FB) * TA
};

// Layernorm + lrelu fused into the forward kernel's epilogue (umma_gemm.cuh): applies when an M tile holds whole
// frames and one N tile holds whole rows.  The epilogue then owns every conv output of its frames: it forms the
// per-frame moments, stores the raw conv output only when the backward will need it, and writes the activation
// as zero-padded bf16 hi / lo planes -- the separate Layernorm pass (a read and a launch per layer) disappears.
struct LnEpi {
  int on;                // 0: plain GEMM epilogue
  int store_c;           // training: keep the raw conv output (C view) for the Layernorm backward
  float* aout;           // activation planes [frames][2][out_flen] bf16 (hi | lo)
  float* mean; float* rstd;
  const float* gamma; const float* beta;   // per channel; channel of column n = n % Cn
  int Cn, L, out_flen, out_off;
};

struct UmmaArgs {
  int K, N;              // logical GEMM sizes (wgrad: dB is [K, N])
  int BN;                // N tile
  int kblocks;           // (F): ceil(K / bk)
  int sw;                // (F): swizzle span = bytes of one operand row per k-block: 128 (bk = 64) or 64 (bk = 32)
  int stages;
  int tmem_cols;
  RowTiling rt;
  // (F)
  int n_tiles, acc_sets;
  // tap mode (conv-shaped views, tapT > 0): the tile's positions are loaded ONCE as tapP phase tiles of
  // [rows + halo][tapC] (no window overlap); tap t = tapP * m + ph multiplies phase tile ph shifted by m
  // rows (row-shifted K-major descriptor) with the resident weight tile of tap t.  sw = 2 * tapC bytes.
  int tapT, tapC, tapP;
  int b_tile_al;         // tap mode: bytes of one resident weight tile (BN * sw rounded up to 1024)
  // merge != 0 (single-CTA forms, 2 * BN <= 256): the hi and lo tiles of the B (wgrad: dC) operand lie back to back in
  // shared memory, so ONE MMA with N = 2 * BN computes  Ah.[Bh | Bl]  into the two adjacent accumulators (main | correction)
  // and a second adds  Al.Bh  to the correction: two MMAs per K step instead of three.  The few-tap layers are paced by the
  // MMA warp's instruction stream (a small-N MMA costs ~90 cycles whatever its N), so this is a third less of it.
  int merge;
  // b_res != 0 (window mode, single CTA, one N tile): the weight tiles of ALL k-blocks (hi, lo) are loaded once per CTA and
  // stay in shared memory in front of the ring, whose stages then hold the activation tiles only -- a persistent CTA's tiles
  // no longer re-fetch the same weights from L2 (layers whose weights fit beside >= 3 activation stages: G0, its dgrad, ...)
  int b_res;
  DView C;
  const float* bias0; const float* bias1; const float* bias2; int bias_mod;
  LnEpi ln;
  // (W)
  int d_sw;              // swizzle span (bytes) of the dC boxes: 128 / 64 / 32 -> 64 / 32 / 16 columns per box
  int rows_al;           // rows_tile rounded up to 16 (MMA K step)
  int tiles_per_split;
  int a_boxes;           // 64-column boxes of the A view per stage and plane: 2, or 1 when K <= 64 (the MMA's upper 64 rows then
                         // read the lo plane's box -- finite values that only reach accumulator rows k >= K, which are never added)
  float* out; int ld;
};

}  // namespace npvc
