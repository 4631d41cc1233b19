// tcgen05 (5th-gen tensor core) view-GEMMs for sm_100a, bf16x3:
//   every operand is stored as bf16 hi / lo planes (value = hi + lo, see plan.h Buf::split) and a product
//   is  Ah.Bh  +  (Al.Bh + Ah.Bl)  with fp32 accumulation in TMEM -- ~2^-17 relative, which holds the 1e-4
//   fp32 parity bar of the path at the bf16 MMA rate (a single bf16 / tf32 pass does not hold it).
//
//   (F) umma_fwd_kernel    C[rows,N] = A_view[rows,K] . B[K,N] (+ bias)
//       A tiles: TMA boxes over the strided view (k, row-in-frame, row-group, frame) = im2col for free;
//       B tiles: K-major [N, kpad] bf16 packs.  Both K-major, 128B swizzle, BK = 64.
//   (W) umma_wgrad_kernel  dB[K,N] += A_view[rows,K]^T . D_view[rows,N]
//       the SAME row-major TMA boxes of both views, consumed as MN-major operands (rows = the MMA's
//       K dimension): no transposing producers.  Reduction over row tiles split across CTAs, RED.ADD out.
// Operand descriptors / box layouts were brought up with tools/umma_bf16_probe.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace npvc {

namespace umma {

constexpr int BM = 128, BK = 64, A_TILE_BYTES = BM * 128;

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// sm_100 shared-memory matrix descriptor.  sw = swizzle span in bytes (128 / 64 / 32).
//   K-major operand : rows of `sw` bytes, 8-row groups SBO = 8*sw apart (LBO unused)
//   MN-major operand: rows (= K index) of `sw` bytes, 8-row groups SBO = 8*sw apart, the next
//                     sw/2 MN elements live in the next box, LBO bytes further on
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sw) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                      // start address   bits [0,14)
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;                   // leading offset  bits [16,30)
  d |= (uint64_t)(((8u * sw) >> 4) & 0x3FFF) << 32;             // stride offset   bits [32,46)
  d |= (uint64_t)1 << 46;                                       // descriptor version = 1
  d |= (uint64_t)(sw == 128 ? 2 : (sw == 64 ? 4 : 6)) << 61;    // swizzle mode
  return d;
}
// kind::f16 with bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// instruction descriptor: D fp32, A/B bf16, M = 128, N = bn; mn_major: both operands MN-major
__device__ __forceinline__ uint32_t make_idesc(int bn, bool mn_major, int m = BM) {
  uint32_t d = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  if (mn_major) d |= (1u << 15) | (1u << 16);
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
// 32-byte global stores (STG.256, sm_100): an epilogue thread owns one accumulator ROW, so its stores are strided by the
// row pitch -- a 16-byte store leaves half of every 32-byte sector to a later instruction (the L1 does not merge them
// and the L2 sees two partial-sector writes); 32 bytes per thread fill whole sectors.
__device__ __forceinline__ void st_global_v8(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void st_global_v8f(float* p, const float* o) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
}
// M tile -> TMA coordinates (row-group a0, frame f0)
__device__ __forceinline__ void tile_coords(const RowTiling& rt, int mt, int& a0, int& f0) {
  const int fb = mt / rt.TA;
  a0 = (mt - fb * rt.TA) * rt.Ab; f0 = fb * rt.FB;
}
// Ring position (stage / accumulator set) and its mbarrier phase, advanced without divisions: the role loops
// of the few-tap layers are paced by their serial instruction streams.
struct RingPos {
  int idx; uint32_t phase; int n;
  __device__ __forceinline__ RingPos(int n_) : idx(0), phase(0u), n(n_) {}
  __device__ __forceinline__ void advance() { if (++idx == n) { idx = 0; phase ^= 1u; } }
};
// Tile iterator of one CTA (tiles blockIdx.x, blockIdx.x + gridDim.x, ...) in (frame block, row-group tile) form
struct TileIter {
  int fb, ta, dfb, dta, TA;
  __device__ __forceinline__ TileIter(const RowTiling& rt) : TA(rt.TA) {
    fb = (int)blockIdx.x / rt.TA; ta = (int)blockIdx.x - fb * rt.TA;
    dfb = (int)gridDim.x / rt.TA; dta = (int)gridDim.x - dfb * rt.TA;
  }
  __device__ __forceinline__ void advance() { fb += dfb; ta += dta; if (ta >= TA) { ta -= TA; fb++; } }
};
// One lane of a converged warp.  The role loops below are executed by ALL lanes of their warp (warp-uniform
// control flow and operands) and only the tcgen05 / TMA instruction itself sits under this predicate:
// issuing them from inside an `if (lane == 0)` region makes the compiler wrap every uniform-datapath
// instruction in an elect / branch loop (measured: ~2000 cycles per k-block of pure issue overhead).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// descriptor without the start address (added per stage: (addr & 0x3FFFF) >> 4)
__device__ __forceinline__ uint64_t sdesc_base(uint32_t lbo, uint32_t sw) { return make_sdesc(0u, lbo, sw); }
__device__ __forceinline__ uint64_t sdesc_at(uint64_t base, uint32_t saddr) { return base | (uint64_t)((saddr & 0x3FFFF) >> 4); }

// ---- CTA pair (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) run ONE 256 x BN MMA.  Each CTA
// stages its own 128 A rows and HALF of the B tile (BN / 2 rows); the leader (cluster rank 0) issues the MMAs,
// which read both CTAs' shared memory at the same offsets and write each CTA's 128 accumulator rows into its
// own TMEM.  Per CTA the B fill per MAC halves.  PTX forms as in CUTLASS (SM100_TMA_2SM_LOAD, SM100_MMA_*_2x1SM,
// umma_arrive_multicast_2x1SM, ClusterBarrier::arrive(cta_id)).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {      // same offset in CTA `rank` of the cluster
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a pair: data lands in the executing CTA, the transaction bytes are counted on `bar`, a
// shared::cluster address (the leader's full barrier)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// The MMAs of one tile in tap mode, taps and phases known at compile time: every descriptor is the stage's
// base descriptor plus a constant (the MMA warp's serial instruction stream paces the few-tap layers).
//   a0d / b0d: descriptors of phase tile 0 (hi plane) of the stage / of tap 0's resident weight tile (hi)
//   a_tile16 / b_tile16 / sw16: tile pitches and the row pitch in 16-byte units
template <int T, int P, bool MERGE>
__device__ __forceinline__ void issue_taps(uint32_t acc, uint32_t acc2, uint64_t a0d, uint64_t b0d, uint32_t a_tile16,
                                           uint32_t b_tile16, uint32_t sw16, int ksteps_c, uint32_t idesc, uint32_t idesc2) {
#pragma unroll
  for (int tp = 0; tp < T; tp++) {
    const int pz = tp % P, m = tp / P;                  // tap tp = P * m + pz: phase tile pz, shifted by m rows
    const uint64_t ah = a0d + (uint64_t)((uint32_t)(2 * pz) * a_tile16 + (uint32_t)m * sw16), al = ah + a_tile16;
    const uint64_t bh = b0d + (uint64_t)((uint32_t)(2 * tp) * b_tile16), bl = bh + b_tile16;
    for (int k4 = 0; k4 < ksteps_c; k4++) {             // UMMA_K = 16 bf16 = 32 bytes -> +2 in the (addr >> 4) field
      const uint64_t o = (uint64_t)(k4 * 2);
      const uint32_t first = (tp > 0 || k4 > 0) ? 1u : 0u;
      if constexpr (MERGE) {
        mma_bf16(acc, ah + o, bh + o, idesc2, first);   // Ah.[Bh | Bl] -> main | correction accumulators (N = 2 BN)
        mma_bf16(acc2, al + o, bh + o, idesc, 1u);      // + Al.Bh
      } else {
        mma_bf16(acc, ah + o, bh + o, idesc, first);      // main products
        mma_bf16(acc2, ah + o, bl + o, idesc, first);     // corrections: separate accumulator (same order as the merged form)
        mma_bf16(acc2, al + o, bh + o, idesc, 1u);
      }
    }
  }
}

}  // namespace umma

// =============================================================================================
// (F) persistent forward / dgrad kernel, 320 threads:
//   warp 0      TMA producer (A hi, A lo, B hi, B lo per 64-wide k-block), runs ahead across tiles
//   warp 1      TMEM allocator + MMA issuer; up to 4 accumulator sets in TMEM (512 / (2*BN) columns allow)
//   warps 2-9   two epilogue groups of 4 warps (TMEM -> registers -> bias -> fp32 or split store): group 0 takes
//               the CTA's even tiles, group 1 the odd ones, overlapped with the next tiles' mainloops through
//               the accf / acce barriers (small-K layers are bound by the per-tile epilogue latency)
// PAIR (window mode only; launched as clusters of 2 CTAs): the two CTAs of a cluster work on two adjacent M tiles
//   and the same N tile as ONE 256 x BN cta_group::2 MMA -- each CTA stages its own A tile and half of the B
//   tile, the leader (cluster rank 0) issues the MMAs and commits to both CTAs' barriers, every CTA drains its
//   own 128 accumulator rows.  B bytes per CTA and MAC halve, which buys pipeline stages for the wide-N layers.
// =============================================================================================
// LNF (single-CTA forms only): Layernorm + lrelu in the epilogue (launch_args.h, LnEpi) -- its own instantiation, so the
//   plain epilogue's code and registers are untouched by it.
template <bool PAIR, bool LNF = false>
__global__ void __maxnreg__(96)      // (up to 608 threads: 64 + 4 epilogue groups + the extra issuer warp; 96 registers as with 576)
umma_fwd_kernel_t(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, UmmaArgs g) {
  pdl_prologue();
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sw = (uint32_t)g.sw;
  const int bk = g.sw >> 1, ksteps = bk >> 4;
  const uint32_t a_tile_bytes = (uint32_t)BM * sw, b_tile_bytes = (uint32_t)(PAIR ? g.BN >> 1 : g.BN) * sw;   // (PAIR: this CTA's half of the B tile)
  const bool tap = !PAIR && g.tapT > 0;        // (the pair form exists in window mode only)
  // tap mode: [resident weights: tapT x (hi, lo) tiles][stages x tapP x (hi, lo) A tiles]
  // window mode with resident weights (g.b_res): [kblocks x (hi, lo) weight tiles][stages x (hi, lo) A tiles]
  const bool bres = !PAIR && !tap && g.b_res != 0;
  const uint32_t bres_bytes = tap ? (uint32_t)g.tapT * 2u * (uint32_t)g.b_tile_al : (bres ? (uint32_t)g.kblocks * 2u * b_tile_bytes : 0u);
  const uint32_t stage_bytes = tap ? (uint32_t)g.tapP * 2u * a_tile_bytes : 2u * a_tile_bytes + (bres ? 0u : 2u * b_tile_bytes);
  const uint32_t ring = sbase + bres_bytes;
  const uint32_t bar_base = ring + (uint32_t)g.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(g.stages + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * g.stages + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * g.stages + 4 + b); };
  const uint32_t bres_bar = bar_base + 8u * (uint32_t)(2 * g.stages + 8);
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * g.stages + 10);      // (keeps bias_s 16-byte aligned)
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* bias_all = reinterpret_cast<float*>(gen_base + (tmem_slot - sbase) + 16);    // [4][256] effective bias of each epilogue group's current N tile
  // fused Layernorm epilogue (g.ln.on; the launch adds 32 KB): per-column scale / offset, per-row moments, per-block sums
  float* gam_all = bias_all + 4 * 256;                                                // [4][256]
  float* bet_all = gam_all + 4 * 256;                                                 // [4][256]
  float4* ln_stats = reinterpret_cast<float4*>(bet_all + 4 * 256);                    // [4 groups][2][128] (shift, sum, sum of squares) of a row
  float2* ln_blk = reinterpret_cast<float2*>(ln_stats + 4 * 2 * 128);                 // [4 groups][2][128] sums over 8-row blocks of a frame

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a launch of 64 + 128 * groups + 32 * x threads carries x more MMA issuer warps behind the epilogue groups (tap mode)
  const int n_extra = (((int)blockDim.x - 64) & 127) >> 5;
  const int mma2_warp = n_extra > 0 ? (int)(blockDim.x >> 5) - n_extra : 1 << 20;       // first extra issuer warp
  // PAIR: a "tile" of the loops below is a pair tile (M tiles 2 * pm + rank of the two CTAs, one N tile); the pair
  // (cluster) index and count take the place of blockIdx.x / gridDim.x
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int total_tiles = PAIR ? ((g.rt.m_tiles + 1) >> 1) * g.n_tiles : g.rt.m_tiles * g.n_tiles;
#define NPVC_TILE0 (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define NPVC_TILE_STEP (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
    for (int s = 0; s < g.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // (PAIR: the leader's acce barriers collect the epilogue warps of both CTAs)
    for (int b = 0; b < 4; b++) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), PAIR ? 8 : 4); }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {       // the MMA warps of both CTAs allocate the same columns in both SMs' TMEM
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();        // barriers of both CTAs initialised before any remote arrive / TMA signal
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));

  if (warp == 0 && tap) {
    // ------------------------------------------------------------------ TMA producer, tap mode
    if (elect_one()) {                               // the weights of every tap, once per CTA
      mbar_expect_tx(bres_bar, (uint32_t)g.tapT * 2u * b_tile_bytes);
      for (int t = 0; t < g.tapT; t++) {
        tma_load_2d(sbase + (uint32_t)(2 * t) * (uint32_t)g.b_tile_al, &tmBh, bres_bar, t * g.tapC, 0);
        tma_load_2d(sbase + (uint32_t)(2 * t + 1) * (uint32_t)g.b_tile_al, &tmBl, bres_bar, t * g.tapC, 0);
      }
    }
    __syncwarp();
    const uint32_t tx = 2u * (uint32_t)g.tapP * (uint32_t)g.rt.rows_tile * sw;
    RingPos sp(g.stages); TileIter ti(g.rt);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, sp.advance(), ti.advance()) {
      const int a0 = ti.ta * g.rt.Ab, f0 = ti.fb * g.rt.FB;
      const int s = sp.idx;
      mbar_wait(empty_bar(s), sp.phase ^ 1u);
      const uint32_t st = ring + (uint32_t)s * stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), tx);
        for (int p = 0; p < g.tapP; p++) {             // phase p = columns [p * tapC, (p + 1) * tapC) of the s * tapC wide rows
          tma_load_4d(st + (uint32_t)(2 * p) * a_tile_bytes, &tmAh, full_bar(s), p * g.tapC, 0, a0, f0);
          tma_load_4d(st + (uint32_t)(2 * p + 1) * a_tile_bytes, &tmAl, full_bar(s), p * g.tapC, 0, a0, f0);
        }
      }
      __syncwarp();
    }
  } else if ((warp == 1 || warp >= mma2_warp) && tap) {
    // ------------------------------------------------------------------ MMA issuer(s), tap mode
    // The few MMAs of a tap-mode tile cost the issuing warp ~1000 cycles of its own instruction stream (uniform-datapath
    // descriptor arithmetic, barrier waits, commits: ~120 dependent instructions per tile) while the tensor pipe is busy
    // for ~150 -- measured: this warp never waits for data or accumulators, the epilogue warps wait for it
    // (profiles/r2z_ncu_full_top_ops.txt, convT_g2).  With the extra issuer warp of the launch (engine.cu launches one:
    // two issuers) the CTA's tiles are issued round-robin by the issuer warps side by side: stage and accumulator rings
    // are walked with that stride, every barrier still has one producer and one consumer per phase.
    const int nis = 1 + n_extra, me = warp == 1 ? 0 : 1 + (warp - mma2_warp);
    const uint32_t idesc = make_idesc(g.BN, false), idesc2 = make_idesc(2 * g.BN, false);
    const bool merge = g.merge != 0;
    const uint64_t dbase = sdesc_base(0, sw);
    const int ksteps_c = g.tapC >> 4;
    const uint64_t b0d = sdesc_at(dbase, sbase);
    mbar_wait(bres_bar, 0);
    RingPos sp(g.stages), ap(g.acc_sets);
    for (int i = 0; i < me; i++) { sp.advance(); ap.advance(); }
    for (int t = blockIdx.x + me * (int)gridDim.x; t < total_tiles; t += nis * (int)gridDim.x) {
      const int buf = ap.idx;
      mbar_wait(acce_bar(buf), ap.phase ^ 1u);
      const int s = sp.idx;
      mbar_wait(full_bar(s), sp.phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * g.BN), acc2 = acc + (uint32_t)g.BN;
      const uint32_t st = ring + (uint32_t)s * stage_bytes;
      const uint64_t a0d = sdesc_at(dbase, st);
      if (elect_one()) {
        const uint32_t a16 = a_tile_bytes >> 4, b16 = (uint32_t)g.b_tile_al >> 4, sw16 = sw >> 4;
        // the tap / phase combinations of the supported layer shapes, unrolled; anything else: generic loop
        if (g.tapT == 3 && g.tapP == 1) { if (merge) issue_taps<3, 1, true>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); else issue_taps<3, 1, false>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); }
        else if (g.tapT == 7 && g.tapP == 3) { if (merge) issue_taps<7, 3, true>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); else issue_taps<7, 3, false>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); }
        else if (g.tapT == 4 && g.tapP == 3) { if (merge) issue_taps<4, 3, true>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); else issue_taps<4, 3, false>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); }
        else if (g.tapT == 9 && g.tapP == 3) { if (merge) issue_taps<9, 3, true>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); else issue_taps<9, 3, false>(acc, acc2, a0d, b0d, a16, b16, sw16, ksteps_c, idesc, idesc2); }
        else {
          int pz = 0, m = 0;                              // tap tp = tapP * m + pz
          for (int tp = 0; tp < g.tapT; tp++) {
            const uint64_t ah = a0d + (uint64_t)((uint32_t)(2 * pz) * a16 + (uint32_t)m * sw16), al = ah + a16;   // row shift by m
            const uint64_t bh = b0d + (uint64_t)((uint32_t)(2 * tp) * b16), bl = bh + b16;
            for (int k4 = 0; k4 < ksteps_c; k4++) {
              const uint64_t o = (uint64_t)(k4 * 2);
              const uint32_t first = (tp > 0 || k4 > 0) ? 1u : 0u;
              if (merge) {
                mma_bf16(acc, ah + o, bh + o, idesc2, first);
                mma_bf16(acc2, al + o, bh + o, idesc, 1u);
              } else {
                mma_bf16(acc, ah + o, bh + o, idesc, first);
                mma_bf16(acc2, ah + o, bl + o, idesc, first);
                mma_bf16(acc2, al + o, bh + o, idesc, 1u);
              }
            }
            if (++pz == g.tapP) { pz = 0; m++; }
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(accf_bar(buf));
      }
      __syncwarp();
      for (int i = 0; i < nis; i++) { sp.advance(); ap.advance(); }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one lane issues)
    // (PAIR: the leader's full barrier counts the bytes of both CTAs' loads)
    const uint32_t tx = (PAIR ? 2u : 1u) * (2u * (uint32_t)g.rt.rows_tile * sw + (bres ? 0u : 2u * b_tile_bytes));
    if (bres) {                                      // the weights of every k-block, once per CTA (one N tile: n0 = 0)
      if (elect_one()) {
        mbar_expect_tx(bres_bar, (uint32_t)g.kblocks * 2u * b_tile_bytes);
        for (int kb = 0; kb < g.kblocks; kb++) {
          tma_load_2d(sbase + (uint32_t)(2 * kb) * b_tile_bytes, &tmBh, bres_bar, kb * bk, 0);
          tma_load_2d(sbase + (uint32_t)(2 * kb + 1) * b_tile_bytes, &tmBl, bres_bar, kb * bk, 0);
        }
      }
      __syncwarp();
    }
    RingPos sp(g.stages);
    for (int t = NPVC_TILE0; t < total_tiles; t += NPVC_TILE_STEP) {
      int mt = t / g.n_tiles; int n0 = (t - mt * g.n_tiles) * g.BN;
      if constexpr (PAIR) {
        mt = 2 * mt + (int)rank; if (mt >= g.rt.m_tiles) mt = g.rt.m_tiles - 1;     // odd tile count: the idle half re-reads the last tile (never stored)
        n0 += (int)rank * (g.BN >> 1);
      }
      int a0, f0; tile_coords(g.rt, mt, a0, f0);
      for (int kb = 0; kb < g.kblocks; kb++, sp.advance()) {
        const int s = sp.idx;
        mbar_wait(empty_bar(s), sp.phase ^ 1u);
        const uint32_t st = ring + (uint32_t)s * stage_bytes;
        if constexpr (PAIR) {
          const uint32_t lead_full = mapa_rank(full_bar(s), 0u);
          if (elect_one()) {
            if (rank == 0u) mbar_expect_tx(full_bar(s), tx);
            tma_load_4d_pair(st, &tmAh, lead_full, kb * bk, 0, a0, f0);
            tma_load_4d_pair(st + a_tile_bytes, &tmAl, lead_full, kb * bk, 0, a0, f0);
            tma_load_2d_pair(st + 2u * a_tile_bytes, &tmBh, lead_full, kb * bk, n0);
            tma_load_2d_pair(st + 2u * a_tile_bytes + b_tile_bytes, &tmBl, lead_full, kb * bk, n0);
          }
        } else if (elect_one()) {
          mbar_expect_tx(full_bar(s), tx);
          tma_load_4d(st, &tmAh, full_bar(s), kb * bk, 0, a0, f0);
          tma_load_4d(st + a_tile_bytes, &tmAl, full_bar(s), kb * bk, 0, a0, f0);
          if (!bres) {
            tma_load_2d(st + 2u * a_tile_bytes, &tmBh, full_bar(s), kb * bk, n0);
            tma_load_2d(st + 2u * a_tile_bytes + b_tile_bytes, &tmBl, full_bar(s), kb * bk, n0);
          }
        }
        __syncwarp();
      }
    }
    if constexpr (PAIR) {
      // producer tail: every stage this CTA filled has been consumed (its multicast commit has arrived here)
      // before the CTA may leave the cluster
      // (a slot that was never filled passes at once: the parity waited for is that of its "previous" phase)
      for (int i = 0; i < g.stages; i++, sp.advance()) mbar_wait(empty_bar(sp.idx), sp.phase ^ 1u);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one lane issues)
    const uint32_t idesc = make_idesc(g.BN, false, PAIR ? 2 * BM : BM);
    const uint32_t idesc_m = make_idesc(2 * g.BN, false, BM);          // (merge: single-CTA forms only)
    const bool merge_w = g.merge != 0;
    const uint64_t dbase = sdesc_base(0, sw);
    if (bres) mbar_wait(bres_bar, 0);
    RingPos sp(g.stages), ap(g.acc_sets);
    for (int t = (PAIR && rank != 0u) ? total_tiles : NPVC_TILE0; t < total_tiles; t += NPVC_TILE_STEP, ap.advance()) {   // (PAIR: the leader issues for both CTAs)
      const int buf = ap.idx;
      mbar_wait(acce_bar(buf), ap.phase ^ 1u);            // epilogue has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * g.BN), acc2 = acc + (uint32_t)g.BN;
      for (int kb = 0; kb < g.kblocks; kb++, sp.advance()) {
        const int s = sp.idx;
        mbar_wait(full_bar(s), sp.phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = ring + (uint32_t)s * stage_bytes;
        const uint64_t ah = sdesc_at(dbase, st), al = sdesc_at(dbase, st + a_tile_bytes);
        const uint32_t bt = bres ? sbase + (uint32_t)(2 * kb) * b_tile_bytes : st + 2u * a_tile_bytes;
        const uint64_t bh = sdesc_at(dbase, bt), bl = sdesc_at(dbase, bt + b_tile_bytes);
        if constexpr (PAIR) {
          if (elect_one()) {
            for (int k4 = 0; k4 < ksteps; k4++) {
              const uint64_t o = (uint64_t)(k4 * 2);
              const uint32_t first = (kb > 0 || k4 > 0) ? 1u : 0u;
              mma_bf16_pair(acc, ah + o, bh + o, idesc, first);
              mma_bf16_pair(acc2, ah + o, bl + o, idesc, first);      // (the accumulation order of the merged single-CTA form)
              mma_bf16_pair(acc2, al + o, bh + o, idesc, 1u);
            }
            umma_commit_pair(empty_bar(s));             // frees the stage in both CTAs
            if (kb == g.kblocks - 1) umma_commit_pair(accf_bar(buf));
          }
        } else if (elect_one()) {
          if (merge_w) {
            for (int k4 = 0; k4 < ksteps; k4++) {           // UMMA_K = 16 bf16 = 32 bytes -> +2 in the (addr >> 4) field
              const uint64_t o = (uint64_t)(k4 * 2);
              const uint32_t first = (kb > 0 || k4 > 0) ? 1u : 0u;
              mma_bf16(acc, ah + o, bh + o, idesc_m, first);     // Ah.[Bh | Bl] -> main | correction accumulators (N = 2 BN)
              mma_bf16(acc2, al + o, bh + o, idesc, 1u);         // + Al.Bh
            }
          } else {
            for (int k4 = 0; k4 < ksteps; k4++) {
              const uint64_t o = (uint64_t)(k4 * 2);
              const uint32_t first = (kb > 0 || k4 > 0) ? 1u : 0u;
              mma_bf16(acc, ah + o, bh + o, idesc, first);         // main products
              mma_bf16(acc2, ah + o, bl + o, idesc, first);        // corrections: separate accumulator
              mma_bf16(acc2, al + o, bh + o, idesc, 1u);           //  (tensor-core fp32 accumulation truncates)
            }
          }
          umma_commit(empty_bar(s));                  // frees the smem stage when these MMAs retire
          if (kb == g.kblocks - 1) umma_commit(accf_bar(buf));
        }
        __syncwarp();
      }
    }
  } else if (warp >= mma2_warp) {
    // (extra issuer warps outside tap mode: nothing to do)
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int lq = warp & 3;                        // TMEM lane quarter this warp may access
    const int row_local = lq * 32 + lane;
    const int eg = (warp - 2) >> 2;                  // epilogue group: tiles with (local tile index & 1) == eg
    const int et = (threadIdx.x - 64) & 127;        // 0..127 within the epilogue group
    float* bias_s = bias_all + eg * 256;             // [4][256]
    // local row -> (frame-in-tile, row-group-in-tile, row-in-group)
    const int grp = row_local / g.rt.RbH, b_in = row_local - grp * g.rt.RbH;
    const int fl = grp / g.rt.Ab, al = grp - fl * g.rt.Ab;
    // Each accumulator set (and its barriers) must be served by ONE group, phase after phase (parity waits):
    // the launch gives 1, 2 or 4 groups, a divisor of the number of sets; local tile lt belongs to group lt % groups.
    const int egmask = (int)((blockDim.x - 64) >> 7) - 1;
    int lt = 0, n0_staged = -1;
    int ln_par = 0;                                       // fused Layernorm: which half of the double-buffered row / block sums
    RingPos ap(g.acc_sets); TileIter ti(g.rt);           // (ti: the n_tiles == 1 fast path; tap mode always)
    for (int t = NPVC_TILE0; t < total_tiles; t += NPVC_TILE_STEP, lt++, ap.advance(), ti.advance()) {
      if ((lt & egmask) != eg) continue;
      int mt = t, n0 = 0;
      if (PAIR || g.n_tiles > 1) { mt = t / g.n_tiles; n0 = (t - mt * g.n_tiles) * g.BN; }
      bool tile_ok = true;
      if constexpr (PAIR) { mt = 2 * mt + (int)rank; tile_ok = mt < g.rt.m_tiles; if (!tile_ok) mt = g.rt.m_tiles - 1; }
      if (n0 != n0_staged) {                        // (bias0 + bias1 + bias2)[n % bias_mod] for this tile's columns (0 without bias)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");          // previous tile's readers are done
        for (int c = et; c < g.BN; c += 128) {
          const int n = n0 + c; float b = 0.f;
          if (g.bias0 && n < g.N) {
            const int bi = n % g.bias_mod;
            b = g.bias0[bi]; if (g.bias1) b += g.bias1[bi]; if (g.bias2) b += g.bias2[bi];
          }
          bias_s[c] = b;
          if constexpr (LNF) {
            const int ch = n < g.N ? n % g.ln.Cn : 0;
            gam_all[eg * 256 + c] = g.ln.gamma[ch]; bet_all[eg * 256 + c] = g.ln.beta[ch];
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        n0_staged = n0;
      }
      const int buf = ap.idx; const uint32_t aph = ap.phase;
      int a0, f0;
      if (PAIR || g.n_tiles > 1) tile_coords(g.rt, mt, a0, f0); else { a0 = ti.ta * g.rt.Ab; f0 = ti.fb * g.rt.FB; }
      const long long f = f0 + fl; const int a = a0 + al;
      const bool row_ok = (!PAIR || tile_ok) && (row_local < g.rt.rows_tile) && (b_in < g.rt.Rb) && (f < g.rt.frames) && (a < g.rt.Ra);
      float* cp = nullptr; uint16_t* chp = nullptr;
      int n_lo = 0, n_hi = 0; bool al16 = false, al32 = false;      // this row's valid columns [n_lo, n_hi); 16 / 32-byte aligned chunks
      if (row_ok) {
        const int j = a * g.rt.Rb + b_in;
        const int inf = j * g.C.rs + g.C.off;
        cp = g.C.p + f * g.C.fs + inf;
        chp = reinterpret_cast<uint16_t*>(g.C.p) + f * 2 * g.C.fs + inf;
        n_hi = g.N;
        if (g.C.pred) { n_lo = inf < 0 ? -inf : 0; if (g.C.flen - inf < n_hi) n_hi = g.C.flen - inf; }
        al16 = g.C.split ? ((((inf + n0) & 7) == 0) && ((g.C.fs & 7) == 0)) : ((reinterpret_cast<uintptr_t>(cp + n0) & 15) == 0);
        al32 = g.C.split ? ((((inf + n0) & 15) == 0) && ((g.C.fs & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.C.p) & 31) == 0))
                         : ((reinterpret_cast<uintptr_t>(cp + n0) & 31) == 0);
      }
      mbar_wait(accf_bar(buf), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * g.BN) + ((uint32_t)(lq * 32) << 16);
      if constexpr (LNF) {
        // ---------------------------------------------------------------- conv + bias + Layernorm + lrelu (util/layers.py:10-66)
        // The tile holds whole frames (Ra == 1, one N tile): row = position, columns = the row's channels.  Every value is
        // read from TMEM twice: sweep 1 forms the row's shifted sums (shift = the row's first value: no cancellation) and
        // keeps the raw conv output for the backward; the rows of a frame are combined in ROW ORDER (8-row blocks, then the
        // blocks), so a frame's statistics do not depend on where in a tile it sits; sweep 2 normalises, applies the
        // per-channel scale / offset and lrelu and writes the zero-padded bf16 hi / lo planes the next layer reads.
        const bool rok = row_ok;
        const int N = g.N; const float fN = (float)N, invL = 1.0f / (float)g.ln.L;
        const float* gam_s = gam_all + eg * 256; const float* bet_s = bet_all + eg * 256;
        float4* st_s = ln_stats + (eg * 2 + ln_par) * 128; float2* bk_s = ln_blk + (eg * 2 + ln_par) * 128; ln_par ^= 1;
        float k0 = 0.f, p1 = 0.f, q1 = 0.f;
        auto stat16 = [&](const uint32_t (&v)[16], const uint32_t (&w)[16], int c0) {
          if (c0 >= N) return;
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; e++) o[e] = __uint_as_float(v[e]) + __uint_as_float(w[e]) + bias_s[c0 + e];
          if (c0 == 0) k0 = o[0];
#pragma unroll
          for (int e = 0; e < 16; e++) {
            if (c0 + e < N) { const float d = o[e] - k0; p1 += d; q1 = fmaf(d, d, q1); }
          }
          if (rok && g.ln.store_c) {
            if (c0 + 16 <= N && ((reinterpret_cast<uintptr_t>(cp + c0) & 31) == 0)) { st_global_v8f(cp + c0, o); st_global_v8f(cp + c0 + 8, o + 8); }
            else {
#pragma unroll
              for (int q = 0; q < 4; q++)
                if (c0 + 4 * q < N) *reinterpret_cast<float4*>(cp + c0 + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
          }
        };
        for (int c0 = 0; c0 < g.BN; c0 += 16) {             // (16 columns per TMEM round trip: the register budget of 576 threads)
          uint32_t v0[16], w0[16];
          tmem_ld16(acc + (uint32_t)c0, v0);
          tmem_ld16(acc + (uint32_t)(g.BN + c0), w0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          stat16(v0, w0, c0);
        }
        st_s[row_local] = make_float4(k0, p1, q1, 0.f);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        const int nblk = (g.rt.Rb + 7) >> 3;
        if (rok && (b_in & 7) == 0) {                     // sums of this 8-row block about the frame's shift (its first row's)
          const float K0 = st_s[row_local - b_in].x;
          float P = 0.f, Q = 0.f;
#pragma unroll
          for (int r = 0; r < 8; r++) {
            if (b_in + r < g.rt.Rb) {
              const float4 s4 = st_s[row_local + r];
              const float d = s4.x - K0;                  // sum (x - K0) = sum (x - k) + N d;  sum (x - K0)^2 = sum (x - k)^2 + 2 d sum (x - k) + N d^2
              P += fmaf(fN, d, s4.y); Q += fmaf(d, fmaf(fN, d, 2.0f * s4.y), s4.z);
            }
          }
          bk_s[grp * nblk + (b_in >> 3)] = make_float2(P, Q);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        float mean = 0.f, rs = 0.f;
        if (rok) {
          const float K0 = st_s[row_local - b_in].x;
          float P = 0.f, Q = 0.f;
          for (int b = 0; b < nblk; b++) { const float2 t2 = bk_s[grp * nblk + b]; P += t2.x; Q += t2.y; }
          const float m = P * invL;
          mean = K0 + m;
          rs = rsqrtf(fmaxf(fmaf(-m, m, Q * invL), 0.f) + NPVC_LN_EPS);
        }
        uint16_t* ahp = reinterpret_cast<uint16_t*>(g.ln.aout) + f * 2 * g.ln.out_flen + g.ln.out_off + b_in * N;
        const bool a32 = ((reinterpret_cast<uintptr_t>(ahp) & 31) == 0) && ((g.ln.out_flen & 15) == 0);     // 32-byte plane stores
        auto norm16 = [&](const uint32_t (&v)[16], const uint32_t (&w)[16], int c0) {
          if (!rok || c0 >= N) return;
          uint4 hh[2], ll[2];
#pragma unroll
          for (int h8 = 0; h8 < 2; h8++) {
            const int cc = c0 + 8 * h8;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const float val = __uint_as_float(v[8 * h8 + e]) + __uint_as_float(w[8 * h8 + e]) + bias_s[cc + e];
              o[e] = lrelu_f(fmaf((val - mean) * rs, gam_s[cc + e], bet_s[cc + e]));
            }
            hh[h8].x = split_pack2(o[0], o[1], ll[h8].x); hh[h8].y = split_pack2(o[2], o[3], ll[h8].y);
            hh[h8].z = split_pack2(o[4], o[5], ll[h8].z); hh[h8].w = split_pack2(o[6], o[7], ll[h8].w);
          }
          if (c0 + 16 <= N && a32) { st_global_v8(ahp + c0, hh[0], hh[1]); st_global_v8(ahp + g.ln.out_flen + c0, ll[0], ll[1]); }
          else {
#pragma unroll
            for (int h8 = 0; h8 < 2; h8++)
              if (c0 + 8 * h8 < N) { *reinterpret_cast<uint4*>(ahp + c0 + 8 * h8) = hh[h8]; *reinterpret_cast<uint4*>(ahp + g.ln.out_flen + c0 + 8 * h8) = ll[h8]; }
          }
        };
        for (int c0 = 0; c0 < g.BN; c0 += 16) {
          uint32_t v0[16], w0[16];
          tmem_ld16(acc + (uint32_t)c0, v0);
          tmem_ld16(acc + (uint32_t)(g.BN + c0), w0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          norm16(v0, w0, c0);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(buf));          // the accumulator set may be overwritten now
        if (rok && (b_in == 0 || b_in == g.rt.Rb - 1)) {    // frame statistics and the zero pads around the frame
          uint16_t* fh = reinterpret_cast<uint16_t*>(g.ln.aout) + f * 2 * g.ln.out_flen;
          const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
          if (b_in == 0) {
            g.ln.mean[f] = mean; g.ln.rstd[f] = rs;
            for (int e = 0; e < g.ln.out_off; e += 8) { *reinterpret_cast<uint4*>(fh + e) = z4; *reinterpret_cast<uint4*>(fh + g.ln.out_flen + e) = z4; }
          }
          if (b_in == g.rt.Rb - 1)
            for (int e = g.ln.out_off + g.ln.L; e < g.ln.out_flen; e += 8) { *reinterpret_cast<uint4*>(fh + e) = z4; *reinterpret_cast<uint4*>(fh + g.ln.out_flen + e) = z4; }
        }
        continue;
      }
      // 16 accumulator columns -> bias -> store (whole aligned chunk / aligned groups of 4 / single elements)
      auto emit16 = [&](const uint32_t (&v)[16], const uint32_t (&w)[16], int c0) {
        const int nb = n0 + c0;
        if (!row_ok || nb >= n_hi || nb + 16 <= n_lo) return;
        float o[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);
          o[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + __uint_as_float(w[4 * q + 0]) + b4.x;
          o[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + __uint_as_float(w[4 * q + 1]) + b4.y;
          o[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + __uint_as_float(w[4 * q + 2]) + b4.z;
          o[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + __uint_as_float(w[4 * q + 3]) + b4.w;
        }
        if (al32 && nb >= n_lo && nb + 16 <= n_hi) {                        // the common case: whole 32-byte aligned chunk
          if (g.C.split) {
            uint4 hh[2], ll[2];
#pragma unroll
            for (int h8 = 0; h8 < 2; h8++) {
              hh[h8].x = split_pack2(o[8 * h8 + 0], o[8 * h8 + 1], ll[h8].x); hh[h8].y = split_pack2(o[8 * h8 + 2], o[8 * h8 + 3], ll[h8].y);
              hh[h8].z = split_pack2(o[8 * h8 + 4], o[8 * h8 + 5], ll[h8].z); hh[h8].w = split_pack2(o[8 * h8 + 6], o[8 * h8 + 7], ll[h8].w);
            }
            st_global_v8(chp + nb, hh[0], hh[1]); st_global_v8(chp + g.C.fs + nb, ll[0], ll[1]);
          } else {
            st_global_v8f(cp + nb, o); st_global_v8f(cp + nb + 8, o + 8);
          }
        } else if (al16 && nb >= n_lo && nb + 16 <= n_hi) {                 // whole 16-byte aligned chunk
          if (g.C.split) {
#pragma unroll
            for (int h8 = 0; h8 < 2; h8++) {
              uint4 hh, ll;
              hh.x = split_pack2(o[8 * h8 + 0], o[8 * h8 + 1], ll.x); hh.y = split_pack2(o[8 * h8 + 2], o[8 * h8 + 3], ll.y);
              hh.z = split_pack2(o[8 * h8 + 4], o[8 * h8 + 5], ll.z); hh.w = split_pack2(o[8 * h8 + 6], o[8 * h8 + 7], ll.w);
              *reinterpret_cast<uint4*>(chp + nb + 8 * h8) = hh; *reinterpret_cast<uint4*>(chp + g.C.fs + nb + 8 * h8) = ll;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
              *reinterpret_cast<float4*>(cp + nb + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        } else {                                                              // edges: N tail, predicated range, odd alignment
#pragma unroll
          for (int h8 = 0; h8 < 2; h8++) {
            // a whole aligned 8-column half (N = 24, 40, 88 ... end in one): one 32-byte fp32 store / one 16-byte store per plane,
            // or nothing at all when the half lies beyond N -- without the per-column checks below
            const int n8 = nb + 8 * h8;
            if (n8 >= n_hi || n8 + 8 <= n_lo) continue;
            if (n8 >= n_lo && n8 + 8 <= n_hi && (g.C.split ? al16 : al32)) {
              if (g.C.split) {
                uint4 hh, ll;
                hh.x = split_pack2(o[8 * h8 + 0], o[8 * h8 + 1], ll.x); hh.y = split_pack2(o[8 * h8 + 2], o[8 * h8 + 3], ll.y);
                hh.z = split_pack2(o[8 * h8 + 4], o[8 * h8 + 5], ll.z); hh.w = split_pack2(o[8 * h8 + 6], o[8 * h8 + 7], ll.w);
                *reinterpret_cast<uint4*>(chp + n8) = hh; *reinterpret_cast<uint4*>(chp + g.C.fs + n8) = ll;
              } else st_global_v8f(cp + n8, o + 8 * h8);
              continue;
            }
#pragma unroll
            for (int q = 2 * h8; q < 2 * h8 + 2; q++) {
              const int n4 = nb + 4 * q;
              if (al16 && n4 >= n_lo && n4 + 4 <= n_hi) {
                if (g.C.split) split_st4(chp + n4, g.C.fs, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
                else *reinterpret_cast<float4*>(cp + n4) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const int n = n4 + e;
                  if (n >= n_lo && n < n_hi) { if (g.C.split) split_st1(chp + n, g.C.fs, o[4 * q + e]); else cp[n] = o[4 * q + e]; }
                }
              }
            }
          }
        }
      };
      for (int c0 = 0; c0 < g.BN; c0 += 32) {            // two 16-column chunks per TMEM round trip
        uint32_t v0[16], w0[16], v1[16], w1[16];
        const bool two = c0 + 16 < g.BN;
        tmem_ld16(acc + (uint32_t)c0, v0);
        tmem_ld16(acc + (uint32_t)(g.BN + c0), w0);
        if (two) { tmem_ld16(acc + (uint32_t)(c0 + 16), v1); tmem_ld16(acc + (uint32_t)(g.BN + c0 + 16), w1); }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        emit16(v0, w0, c0);
        if (two) emit16(v1, w1, c0 + 16);
      }
      // this accumulator set may be overwritten by the MMA warp now
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if constexpr (PAIR) { if (lane == 0) mbar_arrive_cluster(mapa_rank(acce_bar(buf), 0u)); }   // the leader's barrier
      else if (lane == 0) mbar_arrive(acce_bar(buf));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();        // both CTAs have drained their accumulators and seen every commit
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}
// the single-CTA form (every layer) and the CTA-pair form (opt-in for the wide dense layers, engine.cu)
#undef NPVC_TILE0
#undef NPVC_TILE_STEP
#define umma_fwd_kernel umma_fwd_kernel_t<false, false>
#define umma_fwd_ln_kernel umma_fwd_kernel_t<false, true>
#define umma_fwd_pair_kernel umma_fwd_kernel_t<true, false>

// =============================================================================================
// (W) weight-gradient kernel, 192 threads, grid = (K tiles of 128, N tiles, row-tile splits):
//   warp 0      TMA producer: per row tile 2 x 2 boxes of the A view (64 columns each, hi / lo) and
//               BN / box-width boxes of the dC view -- the same boxes the forward kernel loads, here
//               read by the MMA as MN-major operands (16 view rows per K step)
//   warp 1      TMEM allocator + MMA issuer
//   warps 2-5   RED.ADD epilogue after the last row tile of this CTA's split
// =============================================================================================
// PAIR (clusters of 2 CTAs along x = two adjacent K tiles, same N tile and split): one 256 x BN cta_group::2 MMA per
//   K step; each CTA stages its own 128 A columns and HALF of the dC columns (BN / 2, whole boxes), the leader
//   issues and commits to both CTAs, each CTA adds its own 128 accumulator rows.  Opt-in (engine.cu) until measured.
template <bool PAIR>
__global__ void __launch_bounds__(192, 1)
umma_wgrad_kernel_t(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDl, UmmaArgs g) {
  pdl_prologue();
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int dw = g.d_sw >> 1;                                   // dC columns per box
  const int d_boxes = PAIR ? (g.BN / dw) >> 1 : (g.BN + dw - 1) / dw;      // dC boxes THIS CTA stages (PAIR: its half of the N tile)
  const uint32_t a_region = (uint32_t)g.rows_al * 128u;         // one 64-column A box (rows_al % 16 == 0 -> 1024-aligned)
  const uint32_t d_region = ((uint32_t)g.rows_al * (uint32_t)g.d_sw + 1023u) & ~1023u;
  const uint32_t a_plane = (uint32_t)g.a_boxes * a_region, d_plane = (uint32_t)d_boxes * d_region;
  const uint32_t stage_bytes = 2u * a_plane + 2u * d_plane;
  const uint32_t bar_base = sbase + (uint32_t)g.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(g.stages + s); };
  const uint32_t accum_bar = bar_base + 8u * (uint32_t)(2 * g.stages);
  const uint32_t tmem_slot = accum_bar + 8u;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * g.BN;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                       // (== blockIdx.x & 1)
  const int t_begin = (int)blockIdx.z * g.tiles_per_split;
  int t_end = t_begin + g.tiles_per_split; if (t_end > g.rt.m_tiles) t_end = g.rt.m_tiles;
  if (t_begin >= t_end) return;                       // uniform per CTA
  const int ntl = t_end - t_begin;

  // rows of a box region beyond the TMA box are never written: they must read as zero
  if (g.rows_al > g.rt.rows_tile) {
    for (uint32_t i = threadIdx.x * 16u; i < (uint32_t)g.stages * stage_bytes; i += blockDim.x * 16u)
      *reinterpret_cast<uint4*>(gen_base + i) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmDh); prefetch_tmap(&tmDl);
    for (int s = 0; s < g.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one lane issues)
    const uint32_t tx = (PAIR ? 2u : 1u) * 2u * ((uint32_t)g.a_boxes * (uint32_t)g.rt.rows_tile * 128u + (uint32_t)d_boxes * (uint32_t)g.rt.rows_tile * (uint32_t)g.d_sw);
    RingPos sp(g.stages);
    if constexpr (PAIR) {
      // the K tile past the last one (odd tile count) re-reads the last tile; its rows are never added (k >= K)
      int m0l = m0; if (m0l >= g.K) m0l -= BM;
      const int n0l = n0 + (int)rank * (g.BN >> 1);
      for (int i = 0; i < ntl; i++, sp.advance()) {
        const int s = sp.idx;
        int a0, f0; tile_coords(g.rt, t_begin + i, a0, f0);
        mbar_wait(empty_bar(s), sp.phase ^ 1u);
        const uint32_t st = sbase + (uint32_t)s * stage_bytes;
        const uint32_t lead_full = mapa_rank(full_bar(s), 0u);
        if (elect_one()) {
          if (rank == 0u) mbar_expect_tx(full_bar(s), tx);
#pragma unroll
          for (int b = 0; b < 2; b++) {
            tma_load_4d_pair(st + (uint32_t)b * a_region, &tmAh, lead_full, m0l + 64 * b, 0, a0, f0);
            tma_load_4d_pair(st + a_plane + (uint32_t)b * a_region, &tmAl, lead_full, m0l + 64 * b, 0, a0, f0);
          }
          for (int b = 0; b < d_boxes; b++) {
            tma_load_4d_pair(st + 2u * a_plane + (uint32_t)b * d_region, &tmDh, lead_full, n0l + dw * b, 0, a0, f0);
            tma_load_4d_pair(st + 2u * a_plane + d_plane + (uint32_t)b * d_region, &tmDl, lead_full, n0l + dw * b, 0, a0, f0);
          }
        }
        __syncwarp();
      }
      for (int i = 0; i < g.stages; i++, sp.advance()) mbar_wait(empty_bar(sp.idx), sp.phase ^ 1u);   // producer tail (see the forward kernel)
    } else
    for (int i = 0; i < ntl; i++, sp.advance()) {
      const int s = sp.idx;
      int a0, f0; tile_coords(g.rt, t_begin + i, a0, f0);
      mbar_wait(empty_bar(s), sp.phase ^ 1u);
      const uint32_t st = sbase + (uint32_t)s * stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), tx);
        for (int b = 0; b < g.a_boxes; b++) {
          tma_load_4d(st + (uint32_t)b * a_region, &tmAh, full_bar(s), m0 + 64 * b, 0, a0, f0);
          tma_load_4d(st + a_plane + (uint32_t)b * a_region, &tmAl, full_bar(s), m0 + 64 * b, 0, a0, f0);
        }
        for (int b = 0; b < d_boxes; b++) {
          tma_load_4d(st + 2u * a_plane + (uint32_t)b * d_region, &tmDh, full_bar(s), n0 + dw * b, 0, a0, f0);
          tma_load_4d(st + 2u * a_plane + d_plane + (uint32_t)b * d_region, &tmDl, full_bar(s), n0 + dw * b, 0, a0, f0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one lane issues)
    const uint32_t idesc = make_idesc(g.BN, true, PAIR ? 2 * BM : BM);
    const uint32_t idesc_m = make_idesc(2 * g.BN, true, BM);           // (merge: single-CTA form only)
    const int ksteps = g.rows_al >> 4;
    const uint64_t abase = sdesc_base(a_region, 128), dbase = sdesc_base(d_region, (uint32_t)g.d_sw);
    const uint32_t acc2 = tmem_base + (uint32_t)g.BN;
    RingPos sp(g.stages);
    for (int i = (PAIR && rank != 0u) ? ntl : 0; i < ntl; i++, sp.advance()) {      // (PAIR: the leader issues for both CTAs)
      const int s = sp.idx;
      mbar_wait(full_bar(s), sp.phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st = sbase + (uint32_t)s * stage_bytes;
      const uint64_t ah0 = sdesc_at(abase, st), al0 = sdesc_at(abase, st + a_plane);
      const uint64_t dh0 = sdesc_at(dbase, st + 2u * a_plane), dl0 = sdesc_at(dbase, st + 2u * a_plane + d_plane);
      const uint64_t astep = (uint64_t)((16u * 128u) >> 4), dstep = (uint64_t)((16u * (uint32_t)g.d_sw) >> 4);   // 16 view rows per MMA
      if constexpr (PAIR) {
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t ah = ah0 + astep * ks, al = al0 + astep * ks, dh = dh0 + dstep * ks, dl = dl0 + dstep * ks;
            const uint32_t first = (i > 0 || ks > 0) ? 1u : 0u;
            mma_bf16_pair(tmem_base, ah, dh, idesc, first);
            mma_bf16_pair(acc2, ah, dl, idesc, first);
            mma_bf16_pair(acc2, al, dh, idesc, 1u);
          }
          umma_commit_pair(empty_bar(s));
          if (i == ntl - 1) umma_commit_pair(accum_bar);
        }
      } else if (elect_one()) {
        if (g.merge) {
          for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t ah = ah0 + astep * ks, al = al0 + astep * ks, dh = dh0 + dstep * ks;
            const uint32_t first = (i > 0 || ks > 0) ? 1u : 0u;
            mma_bf16(tmem_base, ah, dh, idesc_m, first);      // Ah^T.[Dh | Dl] -> main | correction accumulators (N = 2 BN)
            mma_bf16(acc2, al, dh, idesc, 1u);                // + Al^T.Dh
          }
        } else {
          for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t ah = ah0 + astep * ks, al = al0 + astep * ks, dh = dh0 + dstep * ks, dl = dl0 + dstep * ks;
            const uint32_t first = (i > 0 || ks > 0) ? 1u : 0u;
            mma_bf16(tmem_base, ah, dh, idesc, first);          // main products
            mma_bf16(acc2, ah, dl, idesc, first);               // corrections
            mma_bf16(acc2, al, dh, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));
        if (i == ntl - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2-5)
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lq = warp & 3;
    const int k = m0 + lq * 32 + lane;              // accumulator row = view column of A = row of dB
    const bool row_ok = k < g.K;
    float* cp = row_ok ? g.out + (long long)k * g.ld : nullptr;
    // Coalesced form: a thread owns one accumulator ROW (one TMEM lane), so adding straight from the registers sends
    // 32 four-byte REDs to 32 different lines per instruction.  The warp's 32 x BN block goes through shared memory
    // instead (the pipeline stages are idle: every MMA has retired) and leaves as 16-byte vector REDs along the rows
    // of dB: whole sectors per request, 8 x fewer L2 atomic operations.
    const int spitch = g.BN + 4;                    // floats; (BN + 4) % 32 == 4 keeps the 16-byte row writes conflict-free
    const bool vec = ((g.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0) && ((n0 & 3) == 0) &&
                     ((size_t)BM * spitch * sizeof(float) <= (size_t)g.stages * stage_bytes);
    if (vec) {
      float* stg = reinterpret_cast<float*>(gen_base) + (size_t)(lq * 32) * spitch;       // this warp's 32 rows
      for (int c0 = 0; c0 < g.BN; c0 += 16) {
        uint32_t v[16], w[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
        tmem_ld16(taddr, v);
        tmem_ld16(taddr + (uint32_t)g.BN, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; q++)
          *reinterpret_cast<float4*>(stg + (size_t)lane * spitch + c0 + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]) + __uint_as_float(w[4 * q]), __uint_as_float(v[4 * q + 1]) + __uint_as_float(w[4 * q + 1]),
                          __uint_as_float(v[4 * q + 2]) + __uint_as_float(w[4 * q + 2]), __uint_as_float(v[4 * q + 3]) + __uint_as_float(w[4 * q + 3]));
      }
      __syncwarp();
      const int kmax = g.K - (m0 + lq * 32);          // rows of this warp inside dB
      for (int r = 0; r < 32 && r < kmax; r++) {
        float* orow = g.out + (long long)(m0 + lq * 32 + r) * g.ld + n0;
        for (int c = 4 * lane; c < g.BN; c += 128) {
          const int n = n0 + c;
          if (n >= g.N) break;
          const float4 t = *reinterpret_cast<const float4*>(stg + (size_t)r * spitch + c);
          if (n + 4 <= g.N) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
          } else {
            atomicAdd(orow + c, t.x);
            if (n + 1 < g.N) atomicAdd(orow + c + 1, t.y);
            if (n + 2 < g.N) atomicAdd(orow + c + 2, t.z);
          }
        }
      }
    } else
    for (int c0 = 0; c0 < g.BN; c0 += 16) {
      uint32_t v[16], w[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
      tmem_ld16(taddr, v);
      tmem_ld16(taddr + (uint32_t)g.BN, w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const int n = n0 + c0 + e;
        if (n < g.N) atomicAdd(cp + n, __uint_as_float(v[e]) + __uint_as_float(w[e]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}
#define umma_wgrad_kernel umma_wgrad_kernel_t<false>
#define umma_wgrad_pair_kernel umma_wgrad_kernel_t<true>

}  // namespace npvc
