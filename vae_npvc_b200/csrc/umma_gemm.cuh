// tcgen05 (5th-gen tensor core) view-GEMM for sm_100a:
//   C[rows,N] = A_view[rows,K] . B[K,N] (+ bias + table[label])      fp32 in, fp32 out
// computed as 3xTF32 (A = Ah + Al, B = Bh + Bl;  Ah.Bh + Al.Bh + Ah.Bl, fp32 accumulate in TMEM)
// so that the result holds the 1e-4 fp32 parity bar of the path (single-pass TF32 does not).
//
// Two kernels share the helpers below: umma_fwd_persistent_kernel (forward / dgrad form, TMA fed)
// and umma_wgrad_kernel (weight-gradient form, producer-warp fed).  Operand tiles are always
// K-major, 128-byte swizzled, BK = 32 fp32 per row; accumulators live in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace npvc {

struct UmmaArgs {
  int K, N;              // logical GEMM sizes (wgrad: dB is [K, N])
  int BN;                // N tile (multiple of 16, 16..256; wgrad: multiple of 32)
  int kblocks;           // forward: ceil(K / 32) reduction blocks
  int stages;            // smem pipeline depth
  int rows_tile;         // forward: valid rows per 128-row M tile (whole frames); wgrad: valid rows per 32-row block
  int FB;                // frames per tile / block
  long long rows;        // total rows
  int tmem_cols;         // power of two >= max(2*BN, 32): main + correction accumulators
  DView C;               // forward: output view
  const float* bias0; const float* bias1; const float* bias2; int bias_mod;
  const float* table; const long long* labels; int table_ld;
  // wgrad only: dB[K,N] += A_view^T . D_view over 32-row reduction blocks
  DView A, D;
  long long nblocks;     // total 32-row reduction blocks
  long long blocks_per_split;
  float* out; int ld;    // dB accumulated with atomics
};

namespace umma {

constexpr int BM = 128, BK = 32, A_TILE_BYTES = BM * BK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    long long t = clock64();
    if (t0 == 0) t0 = t;
    else if (t - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tile (no shared memory, no barrier): shortens the latency of the later TMA load
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart), sm_100 descriptor
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address            bits [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version = 1   bits [46,48)
  d |= (uint64_t)2 << 61;                         // layout = SWIZZLE_128B    bits [61,64)
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cvt_tf32(float x) {
  uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return u;
}

}  // namespace umma

// =============================================================================================
// Weight-gradient (W) kernel:  dB[K,N] += A_view[rows,K]^T . dC_view[rows,N]
//   warp 0      idle (barrier init only)
//   warp 1      TMEM allocator + MMA issuer (same K-major descriptors as the forward kernel)
//   warps 2-9   producers: 4x4 patches of the views -> tf32 hi/lo -> K-major operand tiles
//               (see below); warps 2-5 also run the RED.ADD epilogue
// MODE 1: software-pipelined producer registers, 1 CTA / SM (N tiles wider than 128).
// MODE 2: single register set, <= 102 registers so two CTAs co-reside (BN <= 128).
// =============================================================================================
template <int MODE>
__global__ void __launch_bounds__(320, MODE == 2 ? 2 : 1)
umma_wgrad_kernel(UmmaArgs g) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)g.BN * 128u;
  const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * b_tile_bytes;
  const uint32_t bar_base = sbase + (uint32_t)g.stages * stage_bytes;
  auto ready_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(g.stages + s); };
  const uint32_t accum_bar = bar_base + 8u * (uint32_t)(2 * g.stages);
  const uint32_t tmem_slot = accum_bar + 8u;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int n0 = tile_n * g.BN;
  const long long kb_begin = (long long)blockIdx.z * g.blocks_per_split;
  long long kb_end = kb_begin + g.blocks_per_split; if (kb_end > g.nblocks) kb_end = g.nblocks;
  if (kb_begin >= kb_end) return;                     // uniform per CTA
  const int nkb = (int)(kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.stages; s++) { mbar_init(ready_bar(s), 8); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // both operands K-major (MN-major tf32 operands return zeros on this part: tools/umma_mn_probe.cu)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nkb; i++) {
        const int s = i % g.stages; const uint32_t ph = (uint32_t)((i / g.stages) & 1);
        mbar_wait(ready_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = sbase + (uint32_t)s * stage_bytes;
        const uint64_t ah = make_sdesc(st), al = make_sdesc(st + A_TILE_BYTES);
        const uint64_t bh = make_sdesc(st + 2u * A_TILE_BYTES), bl = make_sdesc(st + 2u * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {               // UMMA_K = 8 tf32 = 32 bytes -> +2 in the (addr >> 4) field
          const uint64_t o = (uint64_t)(k4 * 2);
          const uint32_t first = (i > 0 || k4 > 0) ? 1u : 0u;
          mma_tf32(tmem_base, ah + o, bh + o, idesc, first);                       // main products
          mma_tf32(tmem_base + (uint32_t)g.BN, al + o, bh + o, idesc, first);      // corrections (separate accumulator:
          mma_tf32(tmem_base + (uint32_t)g.BN, ah + o, bl + o, idesc, 1u);         //  tensor-core fp32 accumulation truncates)
        }
        umma_commit(empty_bar(s));                  // frees the smem stage when these MMAs retire
      }
      umma_commit(accum_bar);                       // accumulator complete
    }
  } else if (warp >= 2) {
    // wgrad: these warps ARE the producers.  A task = a 4-row x 4-column patch of a view: four
    // coalesced LDG.128 (a warp reads 512 contiguous bytes of one view row), tf32 hi/lo split,
    // and one 16-byte store per column into the K-major 128B-swizzled operand tile.  The operand
    // row of view column (4*q + i) is PERMUTED to (i * quads + q): consecutive lanes then write
    // consecutive operand rows, whose (row & 7) swizzle phases differ -> conflict-free stores with
    // no register shuffling.  The epilogue applies the inverse permutation.
    const int pw = warp - 2;                        // 0..7: row-quad of this warp's A task
    const int pt = threadIdx.x - 64;                // 0..255 within the producer group
    const int NQ = g.BN >> 2;                       // column quads of the dC tile (multiple of 8)
    constexpr int DT = (MODE == 1) ? 2 : 1;         // dC tasks per thread
    const int ka = tile_m * 128 + 4 * lane;
    const bool a_ok = ka < g.K;
    // Row addressing is incremental (adds only): a task tracks (row, j = row % R, element offset
    // f*fs + j*rs) of its first row; blocks are visited in order, each 32 rows further on.
    const int q32 = 32 / g.A.R, r32 = 32 % g.A.R;                 // A and D views share R
    const long long a_wrap = g.A.fs - (long long)g.A.R * g.A.rs, d_wrap = g.D.fs - (long long)g.D.R * g.D.rs;
    const long long a_blk = (long long)q32 * g.A.fs + (long long)r32 * g.A.rs, d_blk = (long long)q32 * g.D.fs + (long long)r32 * g.D.rs;
    const float* a_base = g.A.p + g.A.off + ka;
    long long a_row = kb_begin * 32 + 4 * pw; int a_j; long long a_off;
    { const long long f = a_row / g.A.R; a_j = (int)(a_row - f * g.A.R); a_off = f * g.A.fs + (long long)a_j * g.A.rs; }
    const float* d_base[DT]; long long d_row[DT], d_off[DT]; int d_j[DT]; bool d_ok[DT];
#pragma unroll
    for (int d = 0; d < DT; d++) {
      const int tsk = pt + d * 256;
      const int rq = tsk / NQ, nq = tsk - rq * NQ;
      const int nn = n0 + 4 * nq;
      d_ok[d] = (tsk < 8 * NQ) && (nn < g.N);
      d_base[d] = g.D.p + g.D.off + nn;
      d_row[d] = kb_begin * 32 + 4 * rq;
      const long long f = d_row[d] / g.D.R; d_j[d] = (int)(d_row[d] - f * g.D.R);
      d_off[d] = f * g.D.fs + (long long)d_j[d] * g.D.rs;
    }
    auto load_block = [&](float4 (&va)[4], float4 (&vd)[DT][4]) {
      {
        long long off = a_off; int j = a_j;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_ok && a_row + e < g.rows) v = __ldg(reinterpret_cast<const float4*>(a_base + off));
          va[e] = v;
          off += g.A.rs; if (++j == g.A.R) { j = 0; off += a_wrap; }
        }
        a_row += 32; a_off += a_blk; a_j += r32; if (a_j >= g.A.R) { a_j -= g.A.R; a_off += a_wrap; }
      }
#pragma unroll
      for (int d = 0; d < DT; d++) {
        long long off = d_off[d]; int j = d_j[d];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (d_ok[d] && d_row[d] + e < g.rows) v = __ldg(reinterpret_cast<const float4*>(d_base[d] + off));
          vd[d][e] = v;
          off += g.D.rs; if (++j == g.D.R) { j = 0; off += d_wrap; }
        }
        d_row[d] += 32; d_off[d] += d_blk; d_j[d] += r32; if (d_j[d] >= g.D.R) { d_j[d] -= g.D.R; d_off[d] += d_wrap; }
      }
    };
    auto split_store = [&](uint8_t* hi_base, uint8_t* lo_base, int rho, int chunk, float x0, float x1, float x2, float x3) {
      const uint32_t off = (uint32_t)(rho >> 3) * 1024u + (uint32_t)(rho & 7) * 128u + (uint32_t)((chunk ^ (rho & 7)) * 16);
      uint4 h, l;
      h.x = cvt_tf32(x0); h.y = cvt_tf32(x1); h.z = cvt_tf32(x2); h.w = cvt_tf32(x3);
      l.x = cvt_tf32(x0 - __uint_as_float(h.x)); l.y = cvt_tf32(x1 - __uint_as_float(h.y));
      l.z = cvt_tf32(x2 - __uint_as_float(h.z)); l.w = cvt_tf32(x3 - __uint_as_float(h.w));
      *reinterpret_cast<uint4*>(hi_base + off) = h;
      *reinterpret_cast<uint4*>(lo_base + off) = l;
    };
    auto store_block = [&](int i, const float4 (&va)[4], const float4 (&vd)[DT][4]) {
      const int s = i % g.stages; const uint32_t ph = (uint32_t)((i / g.stages) & 1);
      mbar_wait(empty_bar(s), ph ^ 1u);
      uint8_t* stp = gen_base + (size_t)s * stage_bytes;
      // A: operand row of view column 4*lane + c is 32*c + lane; 16-byte chunk = row-quad pw
      split_store(stp, stp + A_TILE_BYTES, 0 * 32 + lane, pw, va[0].x, va[1].x, va[2].x, va[3].x);
      split_store(stp, stp + A_TILE_BYTES, 1 * 32 + lane, pw, va[0].y, va[1].y, va[2].y, va[3].y);
      split_store(stp, stp + A_TILE_BYTES, 2 * 32 + lane, pw, va[0].z, va[1].z, va[2].z, va[3].z);
      split_store(stp, stp + A_TILE_BYTES, 3 * 32 + lane, pw, va[0].w, va[1].w, va[2].w, va[3].w);
#pragma unroll
      for (int d = 0; d < DT; d++) {
        const int tsk = pt + d * 256;
        if (tsk < 8 * NQ) {
          const int rq = tsk / NQ, nq = tsk - rq * NQ;
          uint8_t* bh = stp + 2 * A_TILE_BYTES; uint8_t* bl = bh + b_tile_bytes;
          split_store(bh, bl, 0 * NQ + nq, rq, vd[d][0].x, vd[d][1].x, vd[d][2].x, vd[d][3].x);
          split_store(bh, bl, 1 * NQ + nq, rq, vd[d][0].y, vd[d][1].y, vd[d][2].y, vd[d][3].y);
          split_store(bh, bl, 2 * NQ + nq, rq, vd[d][0].z, vd[d][1].z, vd[d][2].z, vd[d][3].z);
          split_store(bh, bl, 3 * NQ + nq, rq, vd[d][0].w, vd[d][1].w, vd[d][2].w, vd[d][3].w);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar(s));
    };
    if (MODE == 1) {
      // software-pipelined: block i+1's loads are in flight while block i is split and stored
      float4 pa[4], pd[DT][4], qa[4], qd[DT][4];
      load_block(pa, pd);
      for (int i = 0; i < nkb; i += 2) {
        if (i + 1 < nkb) load_block(qa, qd);
        store_block(i, pa, pd);
        if (i + 1 < nkb) {
          if (i + 2 < nkb) load_block(pa, pd);
          store_block(i + 1, qa, qd);
        }
      }
    } else {
      // BN <= 128: the second CTA on the SM covers this CTA's load latency
      float4 pa[4], pd[DT][4];
      for (int i = 0; i < nkb; i++) { load_block(pa, pd); store_block(i, pa, pd); }
    }
    if (warp < 6) {
      // ---------------------------------------------------------------- epilogue (warps 2-5)
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int lq = warp & 3;                        // TMEM lane quarter this warp may access
      const int row_local = lq * 32 + lane;
      // accumulator row rho holds view column 4*(rho % 32) + rho / 32 (see the producers)
      const long long r = (long long)tile_m * 128 + 4 * (row_local & 31) + (row_local >> 5);
      const bool row_ok = r < g.K;
      float* cp = row_ok ? g.out + r * g.ld : nullptr;
      const int NQe = g.BN >> 2;
      for (int c0 = 0; c0 < g.BN; c0 += 16) {
        uint32_t v[16], w[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr) : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(taddr + (uint32_t)g.BN) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const int chi = c0 + e;                     // accumulator column chi holds view column 4*(chi % NQ) + chi / NQ
          const int n = n0 + 4 * (chi % NQe) + chi / NQe;
          if (n < g.N) atomicAdd(cp + n, __uint_as_float(v[e]) + __uint_as_float(w[e]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}

// =============================================================================================
// Persistent forward (F) kernel: one CTA per SM slot loops over (M, N) tiles.
//   warp 0      TMA producer, runs ahead across tile boundaries (the smem ring never drains)
//   warp 1      MMA issuer; accumulators double-buffered in TMEM when 4*BN <= 512 columns
//   warps 2-5   tf32 hi/lo converters of the landed A tiles
//   warps 6-9   epilogue (TMEM -> registers -> bias/table -> view store), overlapped with the
//               next tile's mainloop through the accf / acce barriers
// =============================================================================================
struct UmmaFwdArgs {
  UmmaArgs g;
  int m_tiles, n_tiles, acc_sets;
};

__global__ void __launch_bounds__(320, 1)
umma_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                           const __grid_constant__ CUtensorMap tmBl, UmmaFwdArgs pa) {
  using namespace umma;
  const UmmaArgs& g = pa.g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)g.BN * 128u;
  const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * b_tile_bytes;
  const uint32_t bar_base = sbase + (uint32_t)g.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (uint32_t)(g.stages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * g.stages + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (uint32_t)(3 * g.stages + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (uint32_t)(3 * g.stages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(3 * g.stages + 4);
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(gen_base + (tmem_slot - sbase) + 16);      // [256] effective bias of the current N tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = pa.m_tiles * pa.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
    for (int s = 0; s < g.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(ready_bar(s), 4); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; b++) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = (uint32_t)g.rows_tile * 128u + 2u * b_tile_bytes;
      uint32_t it = 0;
      // look-ahead iterator: the A tile of the k-block PF steps ahead is prefetched into L2 (A is
      // the operand that comes from DRAM; the weight tiles stay L2-resident)
      constexpr int PF = 6;
      int pt = blockIdx.x, pkb = 0;
      auto pf_step = [&]() {
        if (pt < total_tiles) {
          tma_prefetch_3d(&tmA, pkb * BK, 0, (pt / pa.n_tiles) * g.FB);
          if (++pkb == g.kblocks) { pkb = 0; pt += gridDim.x; }
        }
      };
      for (int i = 0; i < PF; i++) pf_step();
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tile_m = t / pa.n_tiles, n0 = (t % pa.n_tiles) * g.BN;
        for (int kb = 0; kb < g.kblocks; kb++, it++) {
          const int s = (int)(it % (uint32_t)g.stages); const uint32_t ph = (it / (uint32_t)g.stages) & 1u;
          pf_step();
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = sbase + (uint32_t)s * stage_bytes;
          mbar_expect_tx(full_bar(s), tx);
          tma_load_3d(st, &tmA, full_bar(s), kb * BK, 0, tile_m * g.FB);
          tma_load_2d(st + 2u * A_TILE_BYTES, &tmBh, full_bar(s), kb * BK, n0);
          tma_load_2d(st + 2u * A_TILE_BYTES + b_tile_bytes, &tmBl, full_bar(s), kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0; int lt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, lt++) {
        const int buf = lt % pa.acc_sets; const uint32_t aph = (uint32_t)((lt / pa.acc_sets) & 1);
        mbar_wait(acce_bar(buf), aph ^ 1u);                 // epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * g.BN);
        for (int kb = 0; kb < g.kblocks; kb++, it++) {
          const int s = (int)(it % (uint32_t)g.stages); const uint32_t ph = (it / (uint32_t)g.stages) & 1u;
          mbar_wait(ready_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + (uint32_t)s * stage_bytes;
          const uint64_t ah = make_sdesc(st), al = make_sdesc(st + A_TILE_BYTES);
          const uint64_t bh = make_sdesc(st + 2u * A_TILE_BYTES), bl = make_sdesc(st + 2u * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t o = (uint64_t)(k4 * 2);
            const uint32_t first = (kb > 0 || k4 > 0) ? 1u : 0u;
            mma_tf32(acc, ah + o, bh + o, idesc, first);                       // main products
            mma_tf32(acc + (uint32_t)g.BN, al + o, bh + o, idesc, first);      // corrections (see umma_gemm_kernel)
            mma_tf32(acc + (uint32_t)g.BN, ah + o, bl + o, idesc, 1u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(accf_bar(buf));
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ converters
    const int ct = threadIdx.x - 64;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      for (int kb = 0; kb < g.kblocks; kb++, it++) {
        const int s = (int)(it % (uint32_t)g.stages); const uint32_t ph = (it / (uint32_t)g.stages) & 1u;
        mbar_wait(full_bar(s), ph);
        float4* ahp = reinterpret_cast<float4*>(gen_base + (size_t)s * stage_bytes);
        uint4* alp = reinterpret_cast<uint4*>(gen_base + (size_t)s * stage_bytes + A_TILE_BYTES);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const int idx = ct + q * 128;
          float4 v = ahp[idx];
          uint4 h, l;
          h.x = cvt_tf32(v.x); h.y = cvt_tf32(v.y); h.z = cvt_tf32(v.z); h.w = cvt_tf32(v.w);
          l.x = cvt_tf32(v.x - __uint_as_float(h.x)); l.y = cvt_tf32(v.y - __uint_as_float(h.y));
          l.z = cvt_tf32(v.z - __uint_as_float(h.z)); l.w = cvt_tf32(v.w - __uint_as_float(h.w));
          reinterpret_cast<uint4*>(ahp)[idx] = h;
          alp[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_bar(s));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int lq = warp & 3;
    const int row_local = lq * 32 + lane;
    const int et = threadIdx.x - 192;               // 0..127 within the epilogue group
    int lt = 0, n0_staged = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, lt++) {
      const int tile_m = t / pa.n_tiles, n0 = (t % pa.n_tiles) * g.BN;
      if (g.bias0 && n0 != n0_staged) {             // (bias0 + bias1 + bias2)[n % bias_mod] for this tile's columns
        asm volatile("bar.sync 1, 128;" ::: "memory");          // previous tile's readers are done
        for (int c = et; c < g.BN; c += 128) {
          const int n = n0 + c; float b = 0.f;
          if (n < g.N) {
            const int bi = n % g.bias_mod;
            b = g.bias0[bi]; if (g.bias1) b += g.bias1[bi]; if (g.bias2) b += g.bias2[bi];
          }
          bias_s[c] = b;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        n0_staged = n0;
      }
      const int buf = lt % pa.acc_sets; const uint32_t aph = (uint32_t)((lt / pa.acc_sets) & 1);
      const long long r = (long long)tile_m * g.rows_tile + row_local;
      const bool row_ok = (row_local < g.rows_tile) && (r < g.rows);
      float* cp = nullptr; int inf = 0; const float* trow = nullptr;
      if (row_ok) {
        const long long f = r / g.C.R; const int j = (int)(r - f * g.C.R);
        inf = j * g.C.rs + g.C.off;
        cp = g.C.p + f * g.C.fs + inf;
        if (g.table) trow = g.table + (long long)g.labels[f] * g.table_ld;
      }
      mbar_wait(accf_bar(buf), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * g.BN) + ((uint32_t)(lq * 32) << 16);
      for (int c0 = 0; c0 < g.BN; c0 += 16) {
        uint32_t v[16], w[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(acc + (uint32_t)c0) : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(acc + (uint32_t)(g.BN + c0)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int nb = n0 + c0 + q * 4;
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int n = nb + e;
            float tt = __uint_as_float(v[q * 4 + e]) + __uint_as_float(w[q * 4 + e]);
            if (g.bias0) tt += bias_s[c0 + q * 4 + e];
            if (trow && n < g.N) tt += trow[n];
            o[e] = tt;
          }
          bool full = (nb + 4 <= g.N);
          if (g.C.pred) full = full && (inf + nb >= 0) && (inf + nb + 4 <= g.C.flen);
          if (full && ((reinterpret_cast<uintptr_t>(cp + nb) & 15) == 0)) {
            *reinterpret_cast<float4*>(cp + nb) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int n = nb + e;
              bool ok = n < g.N;
              if (g.C.pred) ok = ok && (inf + n >= 0) && (inf + n < g.C.flen);
              if (ok) cp[n] = o[e];
            }
          }
        }
      }
      // this accumulator set may be overwritten by the MMA warp now
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(buf));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}

}  // namespace npvc
