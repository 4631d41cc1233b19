// Standalone bring-up probe (not part of the product): tcgen05.mma kind::f16 (bf16 in, fp32 accumulate)
// fed by real TMA loads, in the two operand forms the engine uses:
//   F form: C[128, N]   = A[128 rows, K] . B[N, K]^T          both operands K-major (box = 64 k x rows)
//   W form: dB[Kw, Nw]  = A[rows, Kw]^T . D[rows, Nw]         both operands MN-major: the SAME row-major
//           boxes, read through MN-major descriptors (rows = the MMA's K dimension)
// Prints max error vs a CPU reference per configuration.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  int wform;            // 0: F form (K-major), 1: W form (MN-major)
  int N;                // MMA N
  int ksteps;           // number of K=16 MMAs
  uint32_t a_box_bytes; // bytes of one A box in smem; boxes of A: a_boxes
  int a_boxes;
  uint32_t b_box_bytes; int b_boxes;
  int a_sw, b_sw;       // swizzle span in bytes (128 / 64 / 32)
  uint32_t a_region, b_region;   // smem stride between boxes (>= box bytes, multiple of 1024)
  int row_shift;        // F form: the A descriptor starts row_shift rows into the loaded tile (tap of a stride-1 window)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int sw) {
  uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw == 128 ? 2 : sw == 64 ? 4 : 6) << 61;
  return d;
}

__global__ void probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, Cfg c) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, tbar; __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  const uint32_t a_s = sbase, b_s = sbase + c.a_region * c.a_boxes;
  // zero everything first (rows the boxes do not cover must be zero)
  for (uint32_t i = tid * 16; i < c.a_region * c.a_boxes + c.b_region * c.b_boxes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(raw + (sbase - smem_u32(raw)) + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar)), "r"(c.a_box_bytes * c.a_boxes + c.b_box_bytes * c.b_boxes) : "memory");
    for (int b = 0; b < c.a_boxes; b++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(a_s + b * c.a_region), "l"(&tmA), "r"(smem_u32(&tbar)), "r"(c.wform ? b * (c.a_sw / 2) : 0), "r"(0) : "memory");
    for (int b = 0; b < c.b_boxes; b++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(b_s + b * c.b_region), "l"(&tmB), "r"(smem_u32(&tbar)), "r"(c.wform ? b * (c.b_sw / 2) : 0), "r"(0) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&tbar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (c.wform) idesc |= (1u << 15) | (1u << 16);
    for (int ks = 0; ks < c.ksteps; ks++) {
      uint64_t ad, bd;
      if (!c.wform) {   // K-major: 16 bf16 = 32 bytes along the swizzled row per k-step
        ad = make_desc(a_s + ks * 32 + c.row_shift * c.a_sw, 0, 8 * c.a_sw, c.a_sw);
        bd = make_desc(b_s + ks * 32, 0, 8 * c.b_sw, c.b_sw);
      } else {          // MN-major: 16 rows of the box per k-step; LBO = next MN block (box), SBO = next 8-row group
        ad = make_desc(a_s + ks * 16 * c.a_sw, c.a_region, 8 * c.a_sw, c.a_sw);
        bd = make_desc(b_s + ks * 16 * c.b_sw, c.b_region, 8 * c.b_sw, c.b_sw);
      }
      uint32_t acc = ks > 0;
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  { uint32_t done = 0; long long t0 = clock64();
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) __trap();
    } }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t v[16];
    uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; e++) D[(warp * 32 + lane) * c.N + c0 + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

static PFN_encode g_enc;
static CUtensorMapSwizzle swz(int b) { return b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B; }
static int enc2d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t pitch_bytes, uint32_t b0, uint32_t b1, int sw) {
  cuuint64_t gd[2] = {d0, d1}; cuuint64_t gs[1] = {pitch_bytes}; cuuint32_t bx[2] = {b0, b1}; cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(sw),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  return 0;
}

static float rnd() { return (rand() % 2001 - 1000) / 1000.f; }

// W form: A[rows][Kw] (pitch lda), D[rows][Nw] (pitch ldd); dB[128][N] (rows of dB beyond Kw must be 0)
static int run_w(int rows, int Kw, int Nw, int N, int a_sw, int b_sw, int lda, int ldd) {
  std::vector<__nv_bfloat16> A((size_t)rows * lda), Dm((size_t)rows * ldd);
  for (auto& v : A) v = __float2bfloat16(rnd());
  for (auto& v : Dm) v = __float2bfloat16(rnd());
  std::vector<float> R(128 * N, 0.f), O(128 * N);
  for (int m = 0; m < Kw && m < 128; m++) for (int n = 0; n < Nw && n < N; n++) {
    double s = 0; for (int r = 0; r < rows; r++) s += (double)__bfloat162float(A[(size_t)r * lda + m]) * __bfloat162float(Dm[(size_t)r * ldd + n]);
    R[m * N + n] = (float)s;
  }
  __nv_bfloat16 *dA, *dD; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dD, Dm.size() * 2); cudaMalloc(&dO, O.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dD, Dm.data(), Dm.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, O.size() * 4);
  CUtensorMap tA, tB;
  const int a_el = a_sw / 2, b_el = b_sw / 2;
  if (enc2d(&tA, dA, Kw, rows, (uint64_t)lda * 2, a_el, rows, a_sw)) return 1;
  if (enc2d(&tB, dD, Nw, rows, (uint64_t)ldd * 2, b_el, rows, b_sw)) return 1;
  Cfg c; c.wform = 1; c.N = N; c.ksteps = (rows + 15) / 16; c.row_shift = 0;
  c.a_boxes = 128 / a_el; c.b_boxes = (N + b_el - 1) / b_el;
  c.a_box_bytes = rows * a_sw; c.b_box_bytes = rows * b_sw;
  c.a_sw = a_sw; c.b_sw = b_sw;
  c.a_region = ((uint32_t)(c.ksteps * 16 * a_sw) + 1023u) & ~1023u; c.b_region = ((uint32_t)(c.ksteps * 16 * b_sw) + 1023u) & ~1023u;
  size_t smem = (size_t)c.a_region * c.a_boxes + (size_t)c.b_region * c.b_boxes + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, smem>>>(tA, tB, dO, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("W rows=%d Kw=%d Nw=%d: CUDA error %s\n", rows, Kw, Nw, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0, mx = 0;
  for (size_t i = 0; i < O.size(); i++) { double d = fabs((double)O[i] - R[i]); if (!(d <= err)) err = d; mx = fmax(mx, fabs(R[i])); }
  printf("W form rows=%3d Kw=%3d Nw=%3d N=%3d a_sw=%3d b_sw=%3d: max|err| %.3e (max|ref| %.3e)\n", rows, Kw, Nw, N, a_sw, b_sw, err, mx);
  cudaFree(dA); cudaFree(dD); cudaFree(dO);
  return 0;
}

// F form: A[rows<=128][K], B[N][K]; shift: output row m uses A row m + shift (rows - shift valid outputs)
static int run_f(int rows, int K, int N, int sw, int shift = 0) {
  const int kel = sw / 2;    // k elements per box row
  std::vector<__nv_bfloat16> A((size_t)rows * K), B((size_t)N * K);
  for (auto& v : A) v = __float2bfloat16(rnd());
  for (auto& v : B) v = __float2bfloat16(rnd());
  const int Kc = K < kel ? K : kel;
  std::vector<float> R(128 * N, 0.f), O(128 * N);
  for (int m = 0; m + shift < rows; m++) for (int n = 0; n < N; n++) {
    double s = 0; for (int k = 0; k < Kc; k++) s += (double)__bfloat162float(A[(size_t)(m + shift) * K + k]) * __bfloat162float(B[(size_t)n * K + k]);
    R[m * N + n] = (float)s;
  }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, O.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, O.size() * 4);
  CUtensorMap tA, tB;
  if (enc2d(&tA, dA, K, rows, (uint64_t)K * 2, kel, rows, sw)) return 1;
  if (enc2d(&tB, dB, K, N, (uint64_t)K * 2, kel, N, sw)) return 1;
  Cfg c; c.wform = 0; c.N = N; c.ksteps = kel / 16; c.a_boxes = 1; c.b_boxes = 1; c.row_shift = shift;
  c.a_box_bytes = rows * sw; c.b_box_bytes = N * sw; c.a_sw = c.b_sw = sw;
  c.a_region = 128 * sw; c.b_region = ((uint32_t)(N * sw) + 1023u) & ~1023u;
  size_t smem = (size_t)c.a_region + c.b_region + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, smem>>>(tA, tB, dO, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("F rows=%d K=%d N=%d: CUDA error %s\n", rows, K, N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0, mx = 0;
  for (size_t i = 0; i < (size_t)(rows - shift) * N; i++) { double d = fabs((double)O[i] - R[i]); if (!(d <= err)) err = d; mx = fmax(mx, fabs(R[i])); }
  printf("F form rows=%3d K=%3d N=%3d sw=%3d shift=%d: max|err| %.3e (max|ref| %.3e)\n", rows, K, N, sw, shift, err, mx);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return 0;
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode\n"); return 1; }
  g_enc = (PFN_encode)fn;
  srand(1);
  int rc = 0;
  rc |= run_f(128, 64, 64, 128);
  rc |= run_f(114, 48, 32, 128);       // rows < 128 (zeroed tail), K < 64 (TMA zero fill)
  rc |= run_f(128, 32, 64, 64);        // 64-byte swizzle, 32-element k-block
  for (int sh = 1; sh <= 3; sh++) { rc |= run_f(128, 64, 64, 128, sh); rc |= run_f(120, 32, 48, 64, sh); rc |= run_f(118, 16, 32, 32, sh); }
  rc |= run_f(128, 64, 64, 128, 9);
  rc |= run_w(64, 128, 64, 64, 128, 128, 128, 64);
  rc |= run_w(128, 128, 128, 128, 128, 128, 136, 136);
  rc |= run_w(114, 48, 24, 32, 128, 64, 48, 24);    // G2-like: Kw=48 (second A box fully OOB), Nw=24 in a 64-byte box
  rc |= run_w(114, 112, 32, 32, 128, 64, 112, 32);  // E1-like
  rc |= run_w(96, 96, 16, 16, 128, 32, 96, 16);     // 32-byte swizzle D
  rc |= run_w(128, 128, 192, 192, 128, 128, 128, 520);
  return rc;
}
