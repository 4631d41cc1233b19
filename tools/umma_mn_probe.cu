// Standalone probe: tcgen05.mma kind::tf32 with MN-major (and K-major) smem operands written by
// threads in the canonical 128B-swizzled layouts; prints max error vs a CPU reference for a set
// of descriptor variants.  Bring-up tool only (not part of the product).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int a_mn, b_mn; uint32_t lbo, sbo, kstep_bytes; int N; };

__global__ void probe(const float* A /*[128][32] (m,k)*/, const float* B /*[N][32] (n,k)*/, float* D /*[128][N]*/, Cfg c) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (sbase - smem_u32(raw));
  float* sa = reinterpret_cast<float*>(gen);              // 16 KB
  float* sb = reinterpret_cast<float*>(gen + 16384);      // N*128 B
  __shared__ uint64_t bar; __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  // fill operand tiles
  for (int i = tid; i < 128 * 32; i += blockDim.x) {
    int m = i / 32, k = i % 32; uint32_t off;
    if (c.a_mn) off = (m / 32) * 4096 + (k / 8) * 1024 + (k % 8) * 128 + ((((m % 32) / 4) ^ (k % 8)) * 16) + (m % 4) * 4;
    else off = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) * 16) + (k % 4) * 4;
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sa) + off) = A[i];
  }
  for (int i = tid; i < c.N * 32; i += blockDim.x) {
    int n = i / 32, k = i % 32; uint32_t off;
    if (c.b_mn) off = (n / 32) * 4096 + (k / 8) * 1024 + (k % 8) * 128 + ((((n % 32) / 4) ^ (k % 8)) * 16) + (n % 4) * 4;
    else off = (n / 8) * 1024 + (n % 8) * 128 + (((k / 4) ^ (n % 8)) * 16) + (k % 4) * 4;
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sb) + off) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
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
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (c.a_mn) idesc |= 1u << 15;
    if (c.b_mn) idesc |= 1u << 16;
    auto desc = [&](uint32_t addr, int mn) {
      uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
      if (mn) d |= (uint64_t)(c.lbo >> 4) << 16;
      d |= (uint64_t)((mn ? c.sbo : 1024u) >> 4) << 32;
      d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; };
    for (int k4 = 0; k4 < 4; k4++) {
      uint64_t ad = desc(sbase, c.a_mn) + (uint64_t)((c.a_mn ? c.kstep_bytes : 32u) >> 4) * k4;
      uint64_t bd = desc(sbase + 16384, c.b_mn) + (uint64_t)((c.b_mn ? c.kstep_bytes : 32u) >> 4) * k4;
      uint32_t acc = k4 > 0;
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
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

static float tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &b, 4); return r; }

int main() {
  const int N = 64;
  std::vector<float> A(128 * 32), B(N * 32), D(128 * N), R(128 * N);
  srand(1);
  for (auto& v : A) v = tf32((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tf32((rand() % 2001 - 1000) / 1000.f);
  for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) { double s = 0; for (int k = 0; k < 32; k++) s += (double)A[m * 32 + k] * B[n * 32 + k]; R[m * N + n] = (float)s; }
  float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  Cfg cfgs[] = {
    {0, 0, 0, 1024, 32, N},          // K-major both (known good)
    {1, 0, 4096, 1024, 1024, N},     // A MN-major, current convention
    {0, 1, 4096, 1024, 1024, N},     // B MN-major
    {1, 1, 4096, 1024, 1024, N},     // both MN-major (wgrad config)
    {1, 1, 1024, 4096, 1024, N},     // swapped LBO/SBO
    {1, 1, 4096, 1024, 4096, N},     // k-step advances by LBO
    {1, 1, 1024, 4096, 4096, N},
  };
  for (auto& c : cfgs) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, 16384 + N * 128 + 2048>>>(dA, dB, dD, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u: CUDA error %s\n", c.a_mn, c.b_mn, c.lbo, c.sbo, c.kstep_bytes, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); i++) { err = fmax(err, fabs(D[i] - R[i])); mx = fmax(mx, fabs(R[i])); nz += D[i] != 0.f; }
    printf("cfg a_mn=%d b_mn=%d lbo=%4u sbo=%4u kstep=%4u: max|err| %.3e (max|ref| %.3e) nonzero %d/%zu\n", c.a_mn, c.b_mn, c.lbo, c.sbo, c.kstep_bytes, err, mx, nz, D.size());
  }
  return 0;
}
