// Bring-up probe (not part of the product): how fast can one SM fill shared-memory operand tiles of
// 128 rows x 128 bytes (x 2 planes) from global memory -- through TMA boxes vs through cp.async (LDGSTS)
// issued by 4 producer warps -- for the two access patterns of the engine:
//   dense : rows 8208 bytes apart (one row per frame, G3-like), 64 k-blocks per tile
//   conv  : rows 96 bytes apart (overlapping 224-byte windows, E1-like), 2 k-blocks per tile
// Prints GB/s and cycles per 128-byte row for each (pattern, loader).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encode g_enc;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

struct P { long long tiles; int kblocks; long long row_pitch_el; long long tile_pitch_el; long long plane_el; int stages; };

// mode 0: TMA (2 boxes per stage: hi, lo); mode 1: cp.async by 128 producer threads
template <int MODE>
__global__ void __launch_bounds__(192) probe(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmL,
                                             const uint16_t* base, P p, unsigned long long* sink) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = 2 * 16384;
  const uint32_t bar_base = sbase + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar(s)), "r"(MODE == 0 ? 1u : 128u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty_bar(s)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {                       // consumer: waits for full stages and frees them
    if (lane == 0) {
      uint32_t it = 0; unsigned long long acc = 0;
      for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x)
        for (int kb = 0; kb < p.kblocks; kb++, it++) {
          const int s = it % p.stages; const uint32_t ph = (it / p.stages) & 1u;
          mbar_wait(full_bar(s), ph);
          acc += *reinterpret_cast<volatile uint32_t*>(raw + (sbase - smem_u32(raw)) + s * stage_bytes + 64 * (it & 63));
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar(s)) : "memory");
        }
      sink[blockIdx.x] = acc;
    }
  } else if (MODE == 0) {
    if (warp == 0 && lane == 0) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x)
        for (int kb = 0; kb < p.kblocks; kb++, it++) {
          const int s = it % p.stages; const uint32_t ph = (it / p.stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = sbase + s * stage_bytes;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(s)), "r"(stage_bytes) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(st), "l"(&tmH), "r"(full_bar(s)), "r"(kb * 64), "r"((int)(t * 128)) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(st + 16384), "l"(&tmL), "r"(full_bar(s)), "r"(kb * 64), "r"((int)(t * 128)) : "memory");
        }
    }
  } else if (warp < 4) {
    // 128 producer threads: thread -> chunk c = tid % 8 of rows tid / 8 + 16 i (i < 8), both planes; the copy of
    // stage it is committed as one group and published (wait_group -> proxy fence -> arrive) LAG stages later
    const int tid = threadIdx.x, c = tid & 7, r0 = tid >> 3;
    constexpr int LAG = 2;
    uint32_t it = 0; uint32_t pub = 0;
    auto publish = [&](uint32_t j) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(j % p.stages)) : "memory");
    };
    for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const uint16_t* tb = base + t * p.tile_pitch_el;
      for (int kb = 0; kb < p.kblocks; kb++, it++) {
        const int s = it % p.stages; const uint32_t ph = (it / p.stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = sbase + s * stage_bytes;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int r = r0 + 16 * i;
          const uint16_t* src = tb + r * p.row_pitch_el + kb * 64 + c * 8;
          const uint32_t dst = st + r * 128 + ((c ^ (r & 7)) << 4);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16384), "l"(src + p.plane_el) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= LAG) { asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory"); publish(pub++); }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    while (pub < it) publish(pub++);
  }
}

static int enc2d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t pitch_bytes) {
  cuuint64_t gd[2] = {d0, d1}; cuuint64_t gs[1] = {pitch_bytes}; cuuint32_t bx[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  return 0;
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode\n"); return 1; }
  g_enc = (PFN_encode)fn;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t bytes = (size_t)1 << 30;
  uint16_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  unsigned long long* sink; cudaMalloc(&sink, 8 * 1024);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Pat { const char* name; long long rows; int K; long long row_pitch; int kblocks; };
  // dense: 16384 rows of K = 4104 (pitch 2 * 4104: hi then lo plane per row); conv: 933888 rows, K = 112, pitch 48 (overlap)
  Pat pats[2] = {{"dense", 16384, 4104, 2 * 4104, 64}, {"conv", 933888, 112, 48, 2}};
  for (auto& pt : pats) {
    P p; p.tiles = pt.rows / 128; p.kblocks = pt.kblocks; p.row_pitch_el = pt.row_pitch; p.tile_pitch_el = 128 * pt.row_pitch;
    p.plane_el = (pt.K == 4104) ? 4104 : 64 * 1024 * 1024; p.stages = 5;
    CUtensorMap tH, tL;
    if (enc2d(&tH, buf, pt.K, pt.rows, pt.row_pitch * 2) || enc2d(&tL, buf + p.plane_el, pt.K, pt.rows, pt.row_pitch * 2)) return 1;
    for (int mode = 0; mode < 2; mode++) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      const size_t smem = (size_t)p.stages * 32768 + 2048;
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a);
        if (mode == 0) probe<0><<<sms, 192, smem>>>(tH, tL, buf, p, sink); else probe<1><<<sms, 192, smem>>>(tH, tL, buf, p, sink);
        cudaEventRecord(b);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { printf("%s mode %d: %s\n", pt.name, mode, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      const double fill = (double)p.tiles * p.kblocks * 32768.0;
      const double rows = (double)p.tiles * p.kblocks * 256.0;
      printf("%-5s %-8s: %.3f ms, smem fill %.2f TB/s, %.2f cycles/row/SM @1.9GHz (%.1f B/clk/SM)\n", pt.name, mode ? "cp.async" : "TMA", best,
             fill / best / 1e9, best * 1e-3 * 1.9e9 / (rows / sms), fill / sms / (best * 1e-3 * 1.9e9));
    }
  }
  return 0;
}
