// CUDA-core kernels of the ConvVAE hot path (sm_100a).  The tensor-core (tcgen05) GEMM lives in
// umma_gemm.cuh; everything here is fp32 FFMA / bandwidth-shaped work.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch_args.h"

namespace npvc {

// ---- bf16 hi / lo planes ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t split_pack2(float a, float b, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 4 consecutive elements (8-byte aligned) of a split frame: hi at `hi`, lo `lo_delta` bf16 further on
__device__ __forceinline__ float4 split_ld4(const uint16_t* hi, long long lo_delta) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi), l = *reinterpret_cast<const uint2*>(hi + lo_delta);
  float4 v;
  v.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  v.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  v.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  v.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return v;
}
__device__ __forceinline__ float split_ld1(const uint16_t* hi, long long lo_delta) {
  return __uint_as_float((uint32_t)hi[0] << 16) + __uint_as_float((uint32_t)hi[lo_delta] << 16);
}
__device__ __forceinline__ void split_st4(uint16_t* hi, long long lo_delta, float4 v) {
  uint2 h, l;
  h.x = split_pack2(v.x, v.y, l.x); h.y = split_pack2(v.z, v.w, l.y);
  *reinterpret_cast<uint2*>(hi) = h; *reinterpret_cast<uint2*>(hi + lo_delta) = l;
}
__device__ __forceinline__ void split_st1(uint16_t* hi, long long lo_delta, float v) {
  uint32_t l; const uint32_t h = split_pack2(v, 0.f, l);
  hi[0] = (uint16_t)(h & 0xffffu); hi[lo_delta] = (uint16_t)(l & 0xffffu);
}
// 4 consecutive elements of a view row starting at in-frame element index e (fp32 or split storage)
__device__ __forceinline__ float4 view_ld4(const DView& v, long long f, int e) {
  if (!v.split) return *reinterpret_cast<const float4*>(v.p + f * v.fs + e);
  return split_ld4(reinterpret_cast<const uint16_t*>(v.p) + f * 2 * v.fs + e, v.fs);
}
__device__ __forceinline__ float view_ld1(const DView& v, long long f, int e) {
  if (!v.split) return v.p[f * v.fs + e];
  return split_ld1(reinterpret_cast<const uint16_t*>(v.p) + f * 2 * v.fs + e, v.fs);
}

// Programmatic dependent launch: the engine launches every kernel with programmatic stream serialisation, so a kernel's
// blocks may be scheduled while the previous kernel of the stream is still draining.  Every kernel therefore BEGINS with
// griddepcontrol.wait (returns once the previous kernel has completed and its writes are visible: nothing of a kernel's
// work overlaps its producer, only the launch latency and block scheduling do) and then lets its own successor start
// the same way.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float lrelu_f(float x) { return fmaxf(x, 0.02f * x); }

#define NPVC_LN_EPS 1e-5f
// util/layers.py:7: EPSILON = 1e-6 (fp32), used as exp(0) + EPSILON
#define NPVC_ONE_PLUS_EPS (1.0f + 1e-6f)
#define NPVC_LOG_2PI 1.8378770664093453f

// =============================================================================================
// (F) C[rows,N] = A_view[rows,K] . B[K,N] + bias   -- 128 x BN tile, BK = 16
// =============================================================================================
struct GemmArgs {
  DView A; int K; const float* B; int ldb; int N; DView C; long long rows;
  const float* bias0; const float* bias1; const float* bias2; int bias_mod;
};

template <int BN, bool ASCALAR>
__global__ void __launch_bounds__(256) gemm_view_kernel(GemmArgs g) {
  pdl_prologue();
  constexpr int BM = 128, BK = 16, TN = BN / 16, LDA = BM + 4;
  constexpr int NB4 = (4 * BN + 255) / 256;            // float4 B loads per thread
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A loader: thread -> rows (lrow, lrow+64), k-quad kq
  const int lrow = tid >> 2, kq = tid & 3;
  const float* aptr[2]; int ainf[2]; long long afr[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    long long r = m0 + lrow + 64 * h;
    if (r < g.rows) {
      long long f = r / g.A.R; int j = (int)(r - f * g.A.R);
      ainf[h] = j * g.A.rs + g.A.off; afr[h] = f;
      aptr[h] = g.A.p + f * g.A.fs + ainf[h];
    } else { aptr[h] = nullptr; ainf[h] = 0; afr[h] = 0; }
  }
  float4 ra[2]; float4 rb[NB4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int k = k0 + kq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (aptr[h] != nullptr) {
        if (!ASCALAR) {
          if (k < g.K) {
            v = view_ld4(g.A, afr[h], ainf[h] + k);
            if (k + 3 >= g.K) { if (k + 1 >= g.K) v.y = 0.f; if (k + 2 >= g.K) v.z = 0.f; v.w = 0.f; }
          }
        } else {
          float t[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            int kk = k + i; bool ok = kk < g.K;
            if (g.A.pred) { int q = ainf[h] + kk; ok = ok && q >= 0 && q < g.A.flen; }
            t[i] = ok ? view_ld1(g.A, afr[h], ainf[h] + kk) : 0.f;
          }
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      ra[h] = v;
    }
#pragma unroll
    for (int it = 0; it < NB4; it++) {
      int i = tid + it * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < 4 * BN) {
        int kk = i / (BN / 4), nq = i % (BN / 4);
        int k = k0 + kk, n = n0 + nq * 4;
        if (k < g.K && n < g.N) v = *reinterpret_cast<const float4*>(g.B + (long long)k * g.ldb + n);
      }
      rb[it] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      As[buf][kq * 4 + 0][lrow + 64 * h] = ra[h].x; As[buf][kq * 4 + 1][lrow + 64 * h] = ra[h].y;
      As[buf][kq * 4 + 2][lrow + 64 * h] = ra[h].z; As[buf][kq * 4 + 3][lrow + 64 * h] = ra[h].w;
    }
#pragma unroll
    for (int it = 0; it < NB4; it++) {
      int i = tid + it * 256;
      if (i < 4 * BN) { int kk = i / (BN / 4), nq = i % (BN / 4); *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[it]; }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int nk = (g.K + BK - 1) / BK;
  load_tiles(0); store_tiles(0); __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TN == 8) {
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
      } else if (TN == 4) {
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        b[0] = b0.x; b[1 % TN] = b0.y; b[2 % TN] = b0.z; b[3 % TN] = b0.w;
      } else if (TN == 2) {
        float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
        b[0] = b0.x; b[1 % TN] = b0.y;
      } else {
        b[0] = Bs[buf][kk][tx];
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) { store_tiles(buf ^ 1); __syncthreads(); }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    const long long r = m0 + m;
    if (r >= g.rows) continue;
    const long long f = r / g.C.R; const int j = (int)(r - f * g.C.R);
    const int inf = j * g.C.rs + g.C.off;
    float* cp = g.C.p + f * g.C.fs + inf;
    constexpr int NG = (TN == 8) ? 2 : 1;            // column groups
    constexpr int GW = (TN == 8) ? 4 : TN;           // group width
#pragma unroll
    for (int gi = 0; gi < NG; gi++) {
      const int cbase = n0 + ((TN == 8) ? (gi * (BN / 2) + tx * 4) : (tx * TN));
      float v[GW];
#pragma unroll
      for (int j2 = 0; j2 < GW; j2++) {
        const int n = cbase + j2;
        float t = acc[i][gi * GW + j2];
        if (n < g.N) {
          const int bi = n % g.bias_mod;
          if (g.bias0) t += g.bias0[bi];
          if (g.bias1) t += g.bias1[bi];
          if (g.bias2) t += g.bias2[bi];
        }
        v[j2] = t;
      }
      bool full = (cbase + GW <= g.N);
      if (g.C.pred) full = full && (inf + cbase >= 0) && (inf + cbase + GW <= g.C.flen);
      if (g.C.split) {
        uint16_t* hp = reinterpret_cast<uint16_t*>(g.C.p) + f * 2 * g.C.fs + inf;
#pragma unroll
        for (int j2 = 0; j2 < GW; j2++) {
          const int n = cbase + j2;
          bool ok = n < g.N;
          if (g.C.pred) ok = ok && (inf + n >= 0) && (inf + n < g.C.flen);
          if (ok) split_st1(hp + n, g.C.fs, v[j2]);
        }
      } else if (GW == 4 && full && ((reinterpret_cast<uintptr_t>(cp + cbase) & 15) == 0)) {
        *reinterpret_cast<float4*>(cp + cbase) = make_float4(v[0], v[1 % GW], v[2 % GW], v[3 % GW]);
      } else {
#pragma unroll
        for (int j2 = 0; j2 < GW; j2++) {
          const int n = cbase + j2;
          bool ok = n < g.N;
          if (g.C.pred) ok = ok && (inf + n >= 0) && (inf + n < g.C.flen);
          if (ok) cp[n] = v[j2];
        }
      }
    }
  }
}

// =============================================================================================
// (W) dB[K,N] += A_view[rows,K]^T . D_view[rows,N]   -- BTK x BTN tile, split over row ranges
// =============================================================================================
struct WgradArgs {
  DView A; int K; DView D; int N; float* out; int ld; long long rows; long long rows_per_split; int tiles_n;
};

template <int BTK, int BTN, bool ASCALAR>
__global__ void __launch_bounds__(256) wgrad_view_kernel(WgradArgs g) {
  pdl_prologue();
  constexpr int BR = 16, TK = BTK / 16, TNN = BTN / 16;
  constexpr int NA4 = (BR * BTK / 4 + 255) / 256, ND4 = (BR * BTN / 4 + 255) / 256;
  __shared__ __align__(16) float As[2][BR][BTK];
  __shared__ __align__(16) float Ds[2][BR][BTN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tile_k = blockIdx.x / g.tiles_n, tile_n = blockIdx.x % g.tiles_n;
  const int kt0 = tile_k * BTK, nt0 = tile_n * BTN;
  const long long rbeg = (long long)blockIdx.y * g.rows_per_split;
  long long rend = rbeg + g.rows_per_split; if (rend > g.rows) rend = g.rows;
  if (rbeg >= rend) return;

  float4 ra[NA4], rd[ND4];
  auto load_tiles = [&](long long r0) {
#pragma unroll
    for (int it = 0; it < NA4; it++) {
      int i = tid + it * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < BR * BTK / 4) {
        int rr = i / (BTK / 4), q = i % (BTK / 4);
        long long r = r0 + rr; int k = kt0 + q * 4;
        if (r < rend && k < g.K) {
          long long f = r / g.A.R; int j = (int)(r - f * g.A.R);
          int inf = j * g.A.rs + g.A.off;
          if (!ASCALAR) {
            v = view_ld4(g.A, f, inf + k);
          } else {
            float t[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              int kk = k + e; bool ok = kk < g.K;
              if (g.A.pred) { int qq = inf + kk; ok = ok && qq >= 0 && qq < g.A.flen; }
              t[e] = ok ? view_ld1(g.A, f, inf + kk) : 0.f;
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
      ra[it] = v;
    }
#pragma unroll
    for (int it = 0; it < ND4; it++) {
      int i = tid + it * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < BR * BTN / 4) {
        int rr = i / (BTN / 4), q = i % (BTN / 4);
        long long r = r0 + rr; int n = nt0 + q * 4;
        if (r < rend && n < g.N) {
          long long f = r / g.D.R; int j = (int)(r - f * g.D.R);
          v = view_ld4(g.D, f, j * g.D.rs + g.D.off + n);
        }
      }
      rd[it] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int it = 0; it < NA4; it++) {
      int i = tid + it * 256;
      if (i < BR * BTK / 4) { int rr = i / (BTK / 4), q = i % (BTK / 4); *reinterpret_cast<float4*>(&As[buf][rr][q * 4]) = ra[it]; }
    }
#pragma unroll
    for (int it = 0; it < ND4; it++) {
      int i = tid + it * 256;
      if (i < BR * BTN / 4) { int rr = i / (BTN / 4), q = i % (BTN / 4); *reinterpret_cast<float4*>(&Ds[buf][rr][q * 4]) = rd[it]; }
    }
  };

  float acc[TK][TNN];
#pragma unroll
  for (int i = 0; i < TK; i++)
#pragma unroll
    for (int j = 0; j < TNN; j++) acc[i][j] = 0.f;

  const long long nit = (rend - rbeg + BR - 1) / BR;
  load_tiles(rbeg); store_tiles(0); __syncthreads();
  for (long long itr = 0; itr < nit; itr++) {
    const int buf = (int)(itr & 1);
    if (itr + 1 < nit) load_tiles(rbeg + (itr + 1) * BR);
#pragma unroll
    for (int rr = 0; rr < BR; rr++) {
      float a[TK], d[TNN];
      if (TK == 8) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][rr][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][rr][BTK / 2 + ty * 4]);
        a[0] = a0.x; a[1 % TK] = a0.y; a[2 % TK] = a0.z; a[3 % TK] = a0.w; a[4 % TK] = a1.x; a[5 % TK] = a1.y; a[6 % TK] = a1.z; a[7 % TK] = a1.w;
      } else if (TK == 4) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][rr][ty * 4]);
        a[0] = a0.x; a[1 % TK] = a0.y; a[2 % TK] = a0.z; a[3 % TK] = a0.w;
      } else if (TK == 2) {
        float2 a0 = *reinterpret_cast<const float2*>(&As[buf][rr][ty * 2]);
        a[0] = a0.x; a[1 % TK] = a0.y;
      } else { a[0] = As[buf][rr][ty]; }
      if (TNN == 8) {
        float4 d0 = *reinterpret_cast<const float4*>(&Ds[buf][rr][tx * 4]);
        float4 d1 = *reinterpret_cast<const float4*>(&Ds[buf][rr][BTN / 2 + tx * 4]);
        d[0] = d0.x; d[1 % TNN] = d0.y; d[2 % TNN] = d0.z; d[3 % TNN] = d0.w; d[4 % TNN] = d1.x; d[5 % TNN] = d1.y; d[6 % TNN] = d1.z; d[7 % TNN] = d1.w;
      } else if (TNN == 4) {
        float4 d0 = *reinterpret_cast<const float4*>(&Ds[buf][rr][tx * 4]);
        d[0] = d0.x; d[1 % TNN] = d0.y; d[2 % TNN] = d0.z; d[3 % TNN] = d0.w;
      } else if (TNN == 2) {
        float2 d0 = *reinterpret_cast<const float2*>(&Ds[buf][rr][tx * 2]);
        d[0] = d0.x; d[1 % TNN] = d0.y;
      } else { d[0] = Ds[buf][rr][tx]; }
#pragma unroll
      for (int i = 0; i < TK; i++)
#pragma unroll
        for (int j = 0; j < TNN; j++) acc[i][j] = fmaf(a[i], d[j], acc[i][j]);
    }
    if (itr + 1 < nit) { store_tiles(buf ^ 1); __syncthreads(); }
  }
#pragma unroll
  for (int i = 0; i < TK; i++) {
    int kl = (TK == 8) ? ((i < 4) ? ty * 4 + i : BTK / 2 + ty * 4 + (i - 4)) : (ty * TK + i);
    int k = kt0 + kl;
    if (k >= g.K) continue;
#pragma unroll
    for (int j = 0; j < TNN; j++) {
      int nl = (TNN == 8) ? ((j < 4) ? tx * 4 + j : BTN / 2 + tx * 4 + (j - 4)) : (tx * TNN + j);
      int n = nt0 + nl;
      if (n < g.N) atomicAdd(g.out + (long long)k * g.ld + n, acc[i][j]);
    }
  }
}

// =============================================================================================
// (F) for small K, N and millions of rows (E0, G2 fwd / dgrad): one output row per thread.
// The row's K-float window is loaded once into registers (coalesced across the warp: consecutive
// rows are consecutive / overlapping windows), weights are broadcast from shared memory as float4,
// the N outputs leave as one contiguous run.  Bandwidth-shaped: ~0.4 KB per row.
// =============================================================================================
struct RowGemmArgs {
  DView A; int K; const float* B; int ldb; int N; DView C; long long rows;
  const float* bias0; int bias_mod;
};

template <int KMAX, int NMAX, bool ASCALAR, int ROWS>
__global__ void __launch_bounds__(256) rowgemm_kernel(RowGemmArgs g) {
  pdl_prologue();
  // ROWS output rows per thread (rows r, r + 256, ... of the block's slab): every weight fetched
  // from shared memory feeds ROWS FMAs, which moves the kernel from LDS-issue-bound to FMA-bound
  __shared__ __align__(16) float Ws[KMAX * NMAX];
  __shared__ float bs[NMAX];
  const int tid = threadIdx.x;
  for (int i = tid; i < KMAX * NMAX; i += 256) {
    const int k = i / NMAX, n = i % NMAX;
    Ws[i] = (k < g.K && n < g.N) ? g.B[(long long)k * g.ldb + n] : 0.f;
  }
  if (tid < NMAX) bs[tid] = (g.bias0 && tid < g.N) ? g.bias0[tid % g.bias_mod] : 0.f;
  __syncthreads();
  const long long rbase = (long long)blockIdx.x * (256 * ROWS) + tid;
  float x[ROWS][KMAX];
#pragma unroll
  for (int q = 0; q < ROWS; q++) {
    const long long r = rbase + q * 256;
    const bool rok = r < g.rows;
    const long long rr = rok ? r : 0;
    const long long f = rr / g.A.R; const int j = (int)(rr - f * g.A.R);
    const int inf = j * g.A.rs + g.A.off;
    const float* ap = g.A.p + f * g.A.fs + inf;
    if (!ASCALAR) {
#pragma unroll
      for (int k4 = 0; k4 < KMAX / 4; k4++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok && k4 * 4 < g.K) v = view_ld4(g.A, f, inf + k4 * 4);
        x[q][k4 * 4 + 0] = v.x; x[q][k4 * 4 + 1] = v.y; x[q][k4 * 4 + 2] = v.z; x[q][k4 * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        bool ok = rok && k < g.K;
        if (g.A.pred) { const int qq = inf + k; ok = ok && qq >= 0 && qq < g.A.flen; }
        x[q][k] = ok ? ap[k] : 0.f;
      }
    }
  }
  float acc[ROWS][NMAX];
#pragma unroll
  for (int q = 0; q < ROWS; q++)
#pragma unroll
    for (int n = 0; n < NMAX; n++) acc[q][n] = bs[n];
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    if (k < g.K) {
#pragma unroll
      for (int n4 = 0; n4 < NMAX / 4; n4++) {
        const float4 w = *reinterpret_cast<const float4*>(&Ws[k * NMAX + n4 * 4]);
#pragma unroll
        for (int q = 0; q < ROWS; q++) {
          const float xv = x[q][k];
          acc[q][n4 * 4 + 0] = fmaf(xv, w.x, acc[q][n4 * 4 + 0]); acc[q][n4 * 4 + 1] = fmaf(xv, w.y, acc[q][n4 * 4 + 1]);
          acc[q][n4 * 4 + 2] = fmaf(xv, w.z, acc[q][n4 * 4 + 2]); acc[q][n4 * 4 + 3] = fmaf(xv, w.w, acc[q][n4 * 4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < ROWS; q++) {
    const long long r = rbase + q * 256;
    if (r >= g.rows) continue;
    const long long f = r / g.C.R; const int j = (int)(r - f * g.C.R);
    const int inf = j * g.C.rs + g.C.off;
    float* cp = g.C.p + f * g.C.fs + inf;
    const bool vec = !g.C.pred && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) && (g.N % 4 == 0);
    if (vec) {
#pragma unroll
      for (int n4 = 0; n4 < NMAX / 4; n4++)
        if (n4 * 4 < g.N) *reinterpret_cast<float4*>(cp + n4 * 4) = make_float4(acc[q][n4 * 4], acc[q][n4 * 4 + 1], acc[q][n4 * 4 + 2], acc[q][n4 * 4 + 3]);
    } else {
#pragma unroll
      for (int n = 0; n < NMAX; n++) {
        bool ok = n < g.N;
        if (g.C.pred) ok = ok && (inf + n >= 0) && (inf + n < g.C.flen);
        if (ok) cp[n] = acc[q][n];
      }
    }
  }
}

// C[R,N] += A[R,K] . B[K,N] for a handful of rows (R <= 16): K split across blocks, atomics out
// (the per-speaker embedding gradient: 10 x 1596 x 128).
__global__ void __launch_bounds__(128) fewrows_gemm_kernel(const float* A, int lda, int R, int K, const float* B, int ldb, int N,
                                                           float* C, int ldc, int kchunk) {
  pdl_prologue();
  extern __shared__ float As[];                    // [R][kchunk]
  const int k0 = blockIdx.y * kchunk;
  const int kn = min(kchunk, K - k0);
  for (int i = threadIdx.x; i < R * kchunk; i += blockDim.x) {
    const int r = i / kchunk, k = i % kchunk;
    As[i] = (k < kn) ? A[(long long)r * lda + k0 + k] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; r++) acc[r] = 0.f;
  for (int k = 0; k < kn; k++) {
    const float b = B[(long long)(k0 + k) * ldb + n];
#pragma unroll
    for (int r = 0; r < 16; r++) if (r < R) acc[r] = fmaf(As[r * kchunk + k], b, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 16; r++) if (r < R) atomicAdd(&C[(long long)r * ldc + n], acc[r]);
}

// C[R,N] = A[R,K] . B[K,N] + biases for a handful of rows (R <= 16) and a short K: one thread per output column keeps the
// R sums in registers, A sits in shared memory, B is read once, coalesced (the per-speaker table of the merge layer:
// 10 x 128 x 1672 at pack time -- the tiled GEMM spent 33 us of latency on it).
__global__ void __launch_bounds__(256) fewrows_fwd_kernel(const float* A, int lda, int R, int K, const float* B, int ldb, int N,
                                                          float* C, int ldc, const float* bias0, const float* bias1, const float* bias2, int bias_mod) {
  pdl_prologue();
  extern __shared__ float As[];                    // [R][K]
  for (int i = threadIdx.x; i < R * K; i += blockDim.x) As[i] = A[(long long)(i / K) * lda + (i % K)];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; r++) acc[r] = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; k++) {                      // (8 loads in flight: the single wave of blocks is latency-bound)
    const float b = B[(long long)k * ldb + n];
#pragma unroll
    for (int r = 0; r < 16; r++) if (r < R) acc[r] = fmaf(As[r * K + k], b, acc[r]);
  }
  float bs = 0.f;
  if (bias0) { const int bi = n % bias_mod; bs = bias0[bi]; if (bias1) bs += bias1[bi]; if (bias2) bs += bias2[bi]; }
#pragma unroll
  for (int r = 0; r < 16; r++) if (r < R) C[(long long)r * ldc + n] = acc[r] + bs;
}

// Speaker-branch backward, once per call (model/vae.py:51-70: embedding lookup -> fully_connected_1 -> + biases, hoisted
// per speaker): from the per-speaker sums dP[S, Nm] of the merge gradient (rows z.. of the merge weight gradient)
//   blocks [0, nA):   dW_y[k, n] += sum_s emb[s, k] dP[s, n]   and   dbias[n] += sum_s dP[s, n]       (thread = column n)
//   blocks [nA, ..):  demb[s, d] += sum_n dP[s, n] W_y^T[n, d]  over a chunk of 32 columns             (thread = (s, d) pairs)
// One launch instead of three latency-bound ones (a 10-row weight gradient, a 10-row GEMM with K = 1672, a column sum).
constexpr int SPK_CH = 32;
__global__ void __launch_bounds__(256) speaker_bwd_kernel(const float* dP, const float* emb, int lde, const float* WyT, float* dWy, float* demb, int ldd,
                                                          float* dbias, int S, int Z, int Nm, int nA) {
  pdl_prologue();
  extern __shared__ float ssm[];                   // role A: emb [S][Z]; role B: dP chunk [S][SPK_CH]
  if ((int)blockIdx.x < nA) {
    for (int i = threadIdx.x; i < S * Z; i += blockDim.x) ssm[i] = emb[(long long)(i / Z) * lde + (i % Z)];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nm) return;
    float dp[16]; float cs = 0.f;
#pragma unroll
    for (int r = 0; r < 16; r++) { dp[r] = (r < S) ? dP[(long long)r * Nm + n] : 0.f; cs += dp[r]; }
    dbias[n] += cs;
    for (int k = 0; k < Z; k++) {                    // plain stores: this launch is the only writer of dW_y (zeroed per call, one launch per call);
      float a = 0.f;                                 // a read-modify-write here is 128 dependent L2 round trips per thread
#pragma unroll
      for (int r = 0; r < 16; r++) if (r < S) a = fmaf(ssm[r * Z + k], dp[r], a);
      dWy[(long long)k * Nm + n] = a;
    }
  } else {
    const int n0 = ((int)blockIdx.x - nA) * SPK_CH;
    for (int i = threadIdx.x; i < S * SPK_CH; i += blockDim.x) {
      const int r = i / SPK_CH, n = n0 + i % SPK_CH;
      ssm[i] = (n < Nm) ? dP[(long long)r * Nm + n] : 0.f;
    }
    __syncthreads();
    const int nn = min(SPK_CH, Nm - n0);
    for (int pq = threadIdx.x; pq < S * Z; pq += blockDim.x) {          // consecutive threads = consecutive d: W_y^T rows read coalesced
      const int r = pq / Z, d = pq % Z;
      float a = 0.f;
      for (int j = 0; j < nn; j++) a = fmaf(ssm[r * SPK_CH + j], WyT[(long long)(n0 + j) * Z + d], a);
      atomicAdd(&demb[(long long)r * ldd + d], a);
    }
  }
}

// =============================================================================================
// (W) for tiny K x N (first layer: 7 taps x 16 channels) and millions of rows: every thread keeps
// the whole K x N gradient in registers over its rows, one block reduction + K*N atomics per block.
// =============================================================================================
template <int KMAX, int NMAX>
__global__ void __launch_bounds__(256) wgrad_tiny_kernel(WgradArgs g, long long rows_per_block) {
  pdl_prologue();
  __shared__ float red[8][KMAX * NMAX];
  float acc[KMAX][NMAX];
#pragma unroll
  for (int k = 0; k < KMAX; k++)
#pragma unroll
    for (int n = 0; n < NMAX; n++) acc[k][n] = 0.f;
  const long long rb = (long long)blockIdx.x * rows_per_block;
  long long re = rb + rows_per_block; if (re > g.rows) re = g.rows;
  for (long long r = rb + threadIdx.x; r < re; r += 256) {
    const long long fa = r / g.A.R; const int ja = (int)(r - fa * g.A.R);
    const int inf = ja * g.A.rs + g.A.off;
    const float* ap = g.A.p + fa * g.A.fs + inf;
    float x[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      bool ok = k < g.K;
      if (g.A.pred) { const int q = inf + k; ok = ok && q >= 0 && q < g.A.flen; }
      x[k] = ok ? ap[k] : 0.f;
    }
    const long long fd = r / g.D.R; const int jd = (int)(r - fd * g.D.R);
    const int dinf = jd * g.D.rs + g.D.off;
    float d[NMAX];
#pragma unroll
    for (int n4 = 0; n4 < NMAX / 4; n4++) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n4 * 4 < g.N) v = view_ld4(g.D, fd, dinf + n4 * 4);
      d[n4 * 4] = v.x; d[n4 * 4 + 1] = v.y; d[n4 * 4 + 2] = v.z; d[n4 * 4 + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < KMAX; k++)
#pragma unroll
      for (int n = 0; n < NMAX; n++) acc[k][n] = fmaf(x[k], d[n], acc[k][n]);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KMAX; k++)
#pragma unroll
    for (int n = 0; n < NMAX; n++) {
      float v = acc[k][n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[w][k * NMAX + n] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < KMAX * NMAX; i += 256) {
    const int k = i / NMAX, n = i % NMAX;
    if (k < g.K && n < g.N) {
      float v = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ww++) v += red[ww][i];
      atomicAdd(g.out + (long long)k * g.ld + n, v);
    }
  }
}

// =============================================================================================
// (W) for small K x N (<= 48 x 24: the last stride-3 transposed conv) and millions of rows:
// 6 x 4 register blocks (24 FMAs per 4 shared-memory loads), 4 row-slices per block, rows staged
// through shared memory with coalesced 16-byte loads, one reduction + K*N atomics per block.
// =============================================================================================
template <int KT, int NT>
__global__ void __launch_bounds__(192) wgrad_small_kernel(WgradArgs g, long long rows_per_block) {
  pdl_prologue();
  constexpr int RB = 64, TG = (KT / 6) * (NT / 4), SL = 4;
  static_assert(TG * SL == 192, "thread layout");
  __shared__ __align__(16) float As[RB][KT];
  __shared__ __align__(16) float Ds[RB][NT];
  __shared__ float red[SL][KT * NT];
  const int tid = threadIdx.x, sl = tid / TG, tg = tid % TG;
  const int tk = tg / (NT / 4), tn = tg % (NT / 4);
  float acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  const long long rb = (long long)blockIdx.x * rows_per_block;
  long long re = rb + rows_per_block; if (re > g.rows) re = g.rows;
  for (long long r0 = rb; r0 < re; r0 += RB) {
    for (int i = tid; i < RB * (KT / 4); i += 192) {
      const int rr = i / (KT / 4), q = i % (KT / 4);
      const long long r = r0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < re && q * 4 < g.K) {
        const long long f = r / g.A.R; const int j = (int)(r - f * g.A.R);
        v = view_ld4(g.A, f, j * g.A.rs + g.A.off + q * 4);            // (fp32 or bf16 hi / lo planes)
      }
      *reinterpret_cast<float4*>(&As[rr][q * 4]) = v;
    }
    for (int i = tid; i < RB * (NT / 4); i += 192) {
      const int rr = i / (NT / 4), q = i % (NT / 4);
      const long long r = r0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < re && q * 4 < g.N) {
        const long long f = r / g.D.R; const int j = (int)(r - f * g.D.R);
        v = view_ld4(g.D, f, j * g.D.rs + g.D.off + q * 4);
      }
      *reinterpret_cast<float4*>(&Ds[rr][q * 4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = sl; rr < RB; rr += SL) {
      const float2 a0 = *reinterpret_cast<const float2*>(&As[rr][tk * 6]);
      const float2 a1 = *reinterpret_cast<const float2*>(&As[rr][tk * 6 + 2]);
      const float2 a2 = *reinterpret_cast<const float2*>(&As[rr][tk * 6 + 4]);
      const float4 d = *reinterpret_cast<const float4*>(&Ds[rr][tn * 4]);
      const float a[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
#pragma unroll
      for (int i = 0; i < 6; i++) {
        acc[i][0] = fmaf(a[i], d.x, acc[i][0]); acc[i][1] = fmaf(a[i], d.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], d.z, acc[i][2]); acc[i][3] = fmaf(a[i], d.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) red[sl][(tk * 6 + i) * NT + tn * 4 + j] = acc[i][j];
  __syncthreads();
  for (int i = tid; i < KT * NT; i += 192) {
    const int k = i / NT, n = i % NT;
    if (k < g.K && n < g.N) atomicAdd(g.out + (long long)k * g.ld + n, red[0][i] + red[1][i] + red[2][i] + red[3][i]);
  }
}

// =============================================================================================
// block reductions
// =============================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block (blockDim.x multiple of 32, <= 1024); result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) { float t = (lane < nw) ? red[lane] : 0.f; t = warp_sum(t); if (lane == 0) red[32] = t; }
  __syncthreads();
  return red[32];
}

// =============================================================================================
// Layernorm + lrelu forward  (util/layers.py:10-44,147-149): one block per frame
// =============================================================================================
struct LnFwdArgs {
  const float* in; float* mean; float* aout; float* rstd; const float* gamma; const float* beta;
  int L, Cn, out_flen, out_off; long long frames;
  int out_split;      // aout is a split (bf16 hi / lo) buffer
};

__global__ void __launch_bounds__(256) ln_fwd_kernel(LnFwdArgs g) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];     // L floats
  __shared__ float red[40];
  const long long f = blockIdx.x;
  const float* x = g.in + f * g.L;
  const int L4 = g.L >> 2;
  float s = 0.f;
  for (int i = threadIdx.x; i < L4; i += blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(sm)[i] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = block_sum(s, red) / (float)g.L;
  float q = 0.f;
  for (int i = threadIdx.x; i < L4; i += blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(sm)[i];
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = block_sum(q, red) / (float)g.L;
  const float rs = rsqrtf(var + NPVC_LN_EPS);
  if (threadIdx.x == 0) { g.rstd[f] = rs; g.mean[f] = mean; }
  float* ao = g.aout + f * g.out_flen;
  uint16_t* ah = reinterpret_cast<uint16_t*>(g.aout) + f * 2 * g.out_flen;
  const int off4 = g.out_off >> 2, F4 = g.out_flen >> 2;
  for (int i = threadIdx.x; i < F4; i += blockDim.x) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ii = i - off4;
    if (ii >= 0 && ii < L4) {
      float4 v = reinterpret_cast<const float4*>(sm)[ii];
      const int c = (ii * 4) % g.Cn;
      float4 h = make_float4((v.x - mean) * rs, (v.y - mean) * rs, (v.z - mean) * rs, (v.w - mean) * rs);
      o.x = lrelu_f(fmaf(h.x, g.gamma[c], g.beta[c]));
      o.y = lrelu_f(fmaf(h.y, g.gamma[c + 1], g.beta[c + 1]));
      o.z = lrelu_f(fmaf(h.z, g.gamma[c + 2], g.beta[c + 2]));
      o.w = lrelu_f(fmaf(h.w, g.gamma[c + 3], g.beta[c + 3]));
    }
    if (g.out_split) split_st4(ah + 4 * i, g.out_flen, o);
    else reinterpret_cast<float4*>(ao)[i] = o;
  }
}

// =============================================================================================
// Layernorm + lrelu backward: dy (grad wrt the activation) -> dc (grad wrt the conv output, written
// into a zero-padded frame), dgamma, dbeta, dbias.  Persistent blocks, grid-stride over frames.
// =============================================================================================
struct LnBwdArgs {
  const float* dy; const float* cin; const float* mean; const float* rstd; const float* gamma; const float* beta;   // cin = raw conv output
  float* dc; float* dgamma; float* dbeta; float* dbias;
  int L, Cn, out_flen, out_off; long long frames;
  int out_split;      // dc is a split (bf16 hi / lo) buffer
  int prefetch;       // ln_bwd_reg_kernel: the next frame lands in shared memory while this one is computed (off by default:
                      // engine.cu, PREFETCH_MIN_FRAMES)
};

__global__ void __launch_bounds__(256) ln_bwd_kernel(LnBwdArgs g) {
  pdl_prologue();
  extern __shared__ __align__(16) float lsm[];    // [L] dxhat | [L] xhat | [3*Cn] channel sums
  __shared__ float red[40];
  float* sdx = lsm; float* sxh = lsm + g.L; float* chs = lsm + 2 * g.L;
  for (int i = threadIdx.x; i < 3 * g.Cn; i += blockDim.x) chs[i] = 0.f;
  const int L4 = g.L >> 2;
  const bool fixed = ((blockDim.x * 4) % g.Cn) == 0;   // thread -> channel mapping constant across i
  float adg[4] = {0, 0, 0, 0}, adb[4] = {0, 0, 0, 0}, adc[4] = {0, 0, 0, 0};
  const int off4 = g.out_off >> 2, F4 = g.out_flen >> 2;
  const float invL = 1.0f / (float)g.L;
  __syncthreads();
  for (long long f = blockIdx.x; f < g.frames; f += gridDim.x) {
    const float4* dy4 = reinterpret_cast<const float4*>(g.dy + f * g.L);
    const float4* xh4 = reinterpret_cast<const float4*>(g.cin + f * g.L);
    const float rs = g.rstd[f], mu = g.mean[f];
    float s1 = 0.f, s2 = 0.f;
    // pass 1 (global -> smem): dxhat = dy * lrelu'(u) * gamma, partial sums, dgamma / dbeta
    for (int i = threadIdx.x; i < L4; i += blockDim.x) {
      float4 d = dy4[i], h = xh4[i];
      h.x = (h.x - mu) * rs; h.y = (h.y - mu) * rs; h.z = (h.z - mu) * rs; h.w = (h.w - mu) * rs;   // xhat, as the forward formed it
      const int c = (i * 4) % g.Cn;
      float dv[4] = {d.x, d.y, d.z, d.w}, hv[4] = {h.x, h.y, h.z, h.w}, ox[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        float gm = g.gamma[c + e];
        float u = fmaf(hv[e], gm, g.beta[c + e]);
        float du = dv[e] * (u >= 0.f ? 1.0f : 0.02f);
        ox[e] = du * gm;
        s1 += ox[e]; s2 += ox[e] * hv[e];
        if (fixed) { adg[e] += du * hv[e]; adb[e] += du; }
        else { atomicAdd(&chs[c + e], du * hv[e]); atomicAdd(&chs[g.Cn + c + e], du); }
      }
      reinterpret_cast<float4*>(sdx)[i] = make_float4(ox[0], ox[1], ox[2], ox[3]);
      reinterpret_cast<float4*>(sxh)[i] = h;
    }
    s1 = block_sum(s1, red) * invL;
    s2 = block_sum(s2, red) * invL;
    float4* dc4 = reinterpret_cast<float4*>(g.dc + f * g.out_flen);
    uint16_t* dch = reinterpret_cast<uint16_t*>(g.dc) + f * 2 * g.out_flen;
    // pass 2 (smem -> global): dc = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), zero pads
    for (int i = threadIdx.x; i < F4; i += blockDim.x) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ii = i - off4;
      if (ii >= 0 && ii < L4) {
        float4 d = reinterpret_cast<const float4*>(sdx)[ii], h = reinterpret_cast<const float4*>(sxh)[ii];
        o.x = rs * (d.x - s1 - h.x * s2); o.y = rs * (d.y - s1 - h.y * s2);
        o.z = rs * (d.z - s1 - h.z * s2); o.w = rs * (d.w - s1 - h.w * s2);
        if (fixed) { adc[0] += o.x; adc[1] += o.y; adc[2] += o.z; adc[3] += o.w; }
        else {
          const int c = (ii * 4) % g.Cn;
          atomicAdd(&chs[2 * g.Cn + c], o.x); atomicAdd(&chs[2 * g.Cn + c + 1], o.y);
          atomicAdd(&chs[2 * g.Cn + c + 2], o.z); atomicAdd(&chs[2 * g.Cn + c + 3], o.w);
        }
      }
      if (g.out_split) split_st4(dch + 4 * i, g.out_flen, o);
      else dc4[i] = o;
    }
    __syncthreads();                               // smem frame buffers are reused by the next frame
  }
  if (fixed) {
    // pass 1: thread t handles float4 index t + it*blockDim -> channels (4t) mod Cn + {0..3};
    // pass 2: index ii = t + it*blockDim - off4    -> channels (4t - out_off) mod Cn + {0..3}
    const int c1 = (threadIdx.x * 4) % g.Cn;
    const int c2 = (int)(((long long)threadIdx.x * 4 - g.out_off) % g.Cn + g.Cn) % g.Cn;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      atomicAdd(&chs[c1 + e], adg[e]); atomicAdd(&chs[g.Cn + c1 + e], adb[e]); atomicAdd(&chs[2 * g.Cn + c2 + e], adc[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.Cn; i += blockDim.x) {
    atomicAdd(&g.dgamma[i], chs[i]); atomicAdd(&g.dbeta[i], chs[g.Cn + i]); atomicAdd(&g.dbias[i], chs[2 * g.Cn + i]);
  }
}

// =============================================================================================
// Register-resident Layernorm kernels (the fast path): G threads own one frame, each thread holds up to
// 4 units of 8 consecutive elements in registers -- one global read, group reductions by warp shuffles
// (+ one shared-memory hop when G > 32), 16-byte plane stores.  Requirements (else the kernels above):
// L, out_off, out_flen multiples of 8; L <= 32 G; Cn divides 8 G (a thread's channels are then fixed).
// =============================================================================================
// Per-channel sums at the end of a block: lanes P apart in a warp hold the same channels (P = channel period in lanes, a
// power of two), so their sums are combined with shuffles first and only lanes < P touch the shared accumulators.
// (Without this all 256 threads of a block add into the same few addresses: with 8 channels that is a 256-way
// serialised shared-memory atomic per value -- ~80 us per block, measured as 0.1 ms for a 16-frame Layernorm backward.)
__device__ __forceinline__ float lane_period_sum(float v, int P) {
  for (int off = 16; off >= P; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ int lane_period(int channels, int unit) {     // channels / unit lanes, clamped to [1, 32]
  int P = channels / unit; if (P < 1) P = 1; if (P > 32) P = 32;
  return P;
}

template <int G, int NV>
__device__ __forceinline__ void group_sum(float (&v)[NV], float* red /* [8][NV] */) {
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  if (G > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; i++) red[warp * NV + i] = v[i];
    }
    __syncthreads();
    const int w0 = (warp / (G / 32)) * (G / 32);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < G / 32; w++) t += red[(w0 + w) * NV + i];
      v[i] = t;
    }
  }
}
// The same with ONE barrier: the warp partials are double-buffered (red[2][8][NV], `par` toggles per call).  A buffer
// written in call k is next written in call k + 2, and every thread passes the barrier of call k + 1 -- after its reads of
// call k -- before any thread gets there.
template <int G, int NV>
__device__ __forceinline__ void group_sum_db(float (&v)[NV], float* red /* [2][8][NV] */, int& par) {
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  if (G > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rp = red + par * 8 * NV;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; i++) rp[warp * NV + i] = v[i];
    }
    __syncthreads();
    const int w0 = (warp / (G / 32)) * (G / 32);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < G / 32; w++) t += rp[(w0 + w) * NV + i];
      v[i] = t;
    }
    par ^= 1;
  }
}
__device__ __forceinline__ void ld8(const float* p, float (&x)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// 8 consecutive outputs at element index e of a frame: fp32 (two float4) or planes (one uint4 per plane)
__device__ __forceinline__ void st8(float* base, long long f, int flen, int e, const float (&o)[8], int split) {
  if (split) {
    uint16_t* hp = reinterpret_cast<uint16_t*>(base) + f * 2 * flen + e;
    uint4 h, l;
    h.x = split_pack2(o[0], o[1], l.x); h.y = split_pack2(o[2], o[3], l.y);
    h.z = split_pack2(o[4], o[5], l.z); h.w = split_pack2(o[6], o[7], l.w);
    *reinterpret_cast<uint4*>(hp) = h; *reinterpret_cast<uint4*>(hp + flen) = l;
  } else {
    float4* op = reinterpret_cast<float4*>(base + f * flen + e);
    op[0] = make_float4(o[0], o[1], o[2], o[3]); op[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}
// zero the pad units of an output frame: units [0, off8) and [off8 + L8, F8)
__device__ __forceinline__ void zero_pads(float* base, long long f, int flen, int off8, int L8, int F8, int t, int G, int split) {
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int npad = F8 - L8;
  for (int p = t; p < npad; p += G) st8(base, f, flen, 8 * (p < off8 ? p : p + L8), z, split);
}

// Ampere-style asynchronous copies (LDGSTS): the next frame travels to shared memory while this one is computed.
// A thread copies exactly the units it will read itself, so its own wait_group is all the synchronisation they need.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void* smem, const void* gmem, int src_bytes) {      // src_bytes 0: writes a zero
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G>
__global__ void __launch_bounds__(256, 3) ln_fwd_reg_kernel(LnFwdArgs g) {
  pdl_prologue();
  constexpr int V = 4, FPB = 256 / G;
  __shared__ float red[2 * 8];
  int par = 0;
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  const int L8 = g.L >> 3, off8 = g.out_off >> 3, F8 = g.out_flen >> 3;
  float gm[8], bt[8];
  { const int c0 = (8 * t) % g.Cn;
#pragma unroll
    for (int e = 0; e < 8; e++) { gm[e] = g.gamma[c0 + e]; bt[e] = g.beta[c0 + e]; } }
  const float invL = 1.0f / (float)g.L;
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    float x[V][8];
    float s[1] = {0.f};
#pragma unroll
    for (int k = 0; k < V; k++) {
      const int u = t + k * G;
      if (fok && u < L8) {
        ld8(g.in + f * g.L + 8 * u, x[k]);
#pragma unroll
        for (int e = 0; e < 8; e++) s[0] += x[k][e];
      }
    }
    group_sum_db<G, 1>(s, red, par);
    const float mean = s[0] * invL;
    float q[1] = {0.f};
#pragma unroll
    for (int k = 0; k < V; k++) {
      if (fok && t + k * G < L8) {
#pragma unroll
        for (int e = 0; e < 8; e++) { const float d = x[k][e] - mean; q[0] = fmaf(d, d, q[0]); }
      }
    }
    group_sum_db<G, 1>(q, red, par);
    const float rs = rsqrtf(q[0] * invL + NPVC_LN_EPS);
    if (!fok) continue;                    // (no block-wide barrier after this point in the iteration)
    if (t == 0) { g.rstd[f] = rs; g.mean[f] = mean; }
#pragma unroll
    for (int k = 0; k < V; k++) {
      const int u = t + k * G;
      if (u < L8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = lrelu_f(fmaf((x[k][e] - mean) * rs, gm[e], bt[e]));
        st8(g.aout, f, g.out_flen, 8 * (u + off8), o, g.out_split);
      }
    }
    zero_pads(g.aout, f, g.out_flen, off8, L8, F8, t, G, g.out_split);
  }
}

template <int G>
__global__ void __launch_bounds__(256, 2) ln_bwd_reg_kernel(LnBwdArgs g) {
  pdl_prologue();
  constexpr int V = 4, FPB = 256 / G;
  extern __shared__ __align__(16) float chs[];   // [3 * Cn] channel sums: dgamma | dbeta | dbias, [2 * Cn] gamma | beta, [FPB][2][L] landing zone (dy | c)
  __shared__ float red[2 * 16];
  int par = 0;
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  const int L8 = g.L >> 3, off8 = g.out_off >> 3, F8 = g.out_flen >> 3;
  float* sgm = chs + 3 * g.Cn; float* sbt = sgm + g.Cn;
  for (int i = threadIdx.x; i < 3 * g.Cn; i += blockDim.x) chs[i] = 0.f;
  for (int i = threadIdx.x; i < g.Cn; i += blockDim.x) { sgm[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  __syncthreads();
  const int c0 = (8 * t) % g.Cn;
  float adg[8], adb[8], adc[8];
#pragma unroll
  for (int e = 0; e < 8; e++) adg[e] = adb[e] = adc[e] = 0.f;
  const float invL = 1.0f / (float)g.L;
  float* stg = sbt + g.Cn + (size_t)grp * 2 * g.L;     // this frame slot's [dy | c] landing zone (g.prefetch)
  const bool pf = g.prefetch != 0;
  float rs_n = 0.f, mu_n = 0.f;
  auto fetch = [&](long long fbn) {                    // one commit group per call (empty past the end)
    const long long fn = fbn * FPB + grp;
    if (!pf) return;
    if (fn < g.frames) {
      const float* sd = g.dy + fn * g.L; const float* sc = g.cin + fn * g.L;
#pragma unroll
      for (int k = 0; k < V; k++) {
        const int u = t + k * G;
        if (u < L8) {
          cp_async16(stg + 8 * u, sd + 8 * u); cp_async16(stg + 8 * u + 4, sd + 8 * u + 4);
          cp_async16(stg + g.L + 8 * u, sc + 8 * u); cp_async16(stg + g.L + 8 * u + 4, sc + 8 * u + 4);
        }
      }
      rs_n = g.rstd[fn]; mu_n = g.mean[fn];
    }
    cp_async_commit();
  };
  fetch(blockIdx.x);
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    float dx[V][8], xh[V][8];
    float rs = rs_n, mu = mu_n;
    if (pf) {
      cp_async_wait<0>();
#pragma unroll
      for (int k = 0; k < V; k++) {
        const int u = t + k * G;
        if (fok && u < L8) { ld8(stg + 8 * u, dx[k]); ld8(stg + g.L + 8 * u, xh[k]); }
      }
    } else {
      if (fok) { rs = g.rstd[f]; mu = g.mean[f]; }
#pragma unroll
      for (int k = 0; k < V; k++) {
        const int u = t + k * G;
        if (fok && u < L8) { ld8(g.dy + f * g.L + 8 * u, dx[k]); ld8(g.cin + f * g.L + 8 * u, xh[k]); }
      }
    }
    float s[2] = {0.f, 0.f};
    float gm[8], bt[8];                                           // this thread's 8 channels (fixed: Cn | 8 G)
    { const float4 a = *reinterpret_cast<const float4*>(sgm + c0), b = *reinterpret_cast<const float4*>(sgm + c0 + 4);
      const float4 c = *reinterpret_cast<const float4*>(sbt + c0), d = *reinterpret_cast<const float4*>(sbt + c0 + 4);
      gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w; gm[4] = b.x; gm[5] = b.y; gm[6] = b.z; gm[7] = b.w;
      bt[0] = c.x; bt[1] = c.y; bt[2] = c.z; bt[3] = c.w; bt[4] = d.x; bt[5] = d.y; bt[6] = d.z; bt[7] = d.w; }
#pragma unroll
    for (int k = 0; k < V; k++) {
      if (fok && t + k * G < L8) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float h = (xh[k][e] - mu) * rs;                 // xhat, as the forward formed it
          const float u = fmaf(h, gm[e], bt[e]);
          const float du = dx[k][e] * (u >= 0.f ? 1.0f : 0.02f);
          const float ox = du * gm[e];
          s[0] += ox; s[1] = fmaf(ox, h, s[1]);
          adg[e] = fmaf(du, h, adg[e]); adb[e] += du;
          dx[k][e] = ox; xh[k][e] = h;
        }
      }
    }
    group_sum_db<G, 2>(s, red, par);
    fetch(fb + gridDim.x);                 // (the sums consumed everything this thread read from its landing units)
    if (!fok) continue;
    const float s1 = s[0] * invL, s2 = s[1] * invL;
#pragma unroll
    for (int k = 0; k < V; k++) {
      const int u = t + k * G;
      if (u < L8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { o[e] = rs * (dx[k][e] - s1 - xh[k][e] * s2); adc[e] += o[e]; }
        st8(g.dc, f, g.out_flen, 8 * (u + off8), o, g.out_split);
      }
    }
    zero_pads(g.dc, f, g.out_flen, off8, L8, F8, t, G, g.out_split);
  }
  __syncthreads();
  {
    const int P = lane_period(g.Cn, 8); const bool owner = (int)(threadIdx.x & 31) < P;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const float a = lane_period_sum(adg[e], P), b = lane_period_sum(adb[e], P), c = lane_period_sum(adc[e], P);
      if (owner) { atomicAdd(&chs[c0 + e], a); atomicAdd(&chs[g.Cn + c0 + e], b); atomicAdd(&chs[2 * g.Cn + c0 + e], c); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.Cn; i += blockDim.x) {
    atomicAdd(&g.dgamma[i], chs[i]); atomicAdd(&g.dbeta[i], chs[g.Cn + i]); atomicAdd(&g.dbias[i], chs[2 * g.Cn + i]);
  }
}

// =============================================================================================
// Layernorm + lrelu backward for LARGE frames (more than 2048 floats): persistent blocks, the (dy, c) rows of
// frame i+1 arrive by bulk async copy (cp.async.bulk + mbarrier) while frame i is processed from shared
// memory, so the frame stream never stalls on the per-frame reductions.  Same arithmetic and requirements
// as ln_bwd_reg_kernel (units of 8 elements, a thread's channels fixed: Cn divides 2048).
// dynamic smem: [2 stages][2][L] floats | [3 Cn] channel sums | [2 Cn] gamma, beta
// =============================================================================================
__global__ void __launch_bounds__(256, 3) ln_bwd_bulk_kernel(LnBwdArgs g) {
  pdl_prologue();
  extern __shared__ __align__(16) float bsm[];
  __shared__ float red[40];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int L = g.L, L8 = L >> 3, off8 = g.out_off >> 3, F8 = g.out_flen >> 3;
  float* chs = bsm + 4 * L; float* sgm = chs + 3 * g.Cn; float* sbt = sgm + g.Cn;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]);
  const uint32_t bytes = (uint32_t)L * 4u;
  auto issue = [&](long long f, int st) {              // thread 0: both rows of frame f -> stage st
    const uint32_t bar = bar0 + 8u * (uint32_t)st;
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(bsm + (size_t)st * 2 * L), d1 = d0 + bytes;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic accesses to this stage are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d0), "l"(g.dy + f * L), "r"(bytes), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d1), "l"(g.cin + f * L), "r"(bytes), "r"(bar) : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * g.Cn; i += blockDim.x) chs[i] = 0.f;
  for (int i = threadIdx.x; i < g.Cn; i += blockDim.x) { sgm[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < g.frames) issue(blockIdx.x, 0);
  const int c0 = (8 * threadIdx.x) % g.Cn;            // this thread's 8 channels (blockDim * 8 = 2048 is a multiple of Cn)
  float gm[8], bt[8], adg[8], adb[8], adc[8];
#pragma unroll
  for (int e = 0; e < 8; e++) { gm[e] = sgm[c0 + e]; bt[e] = sbt[c0 + e]; adg[e] = adb[e] = adc[e] = 0.f; }
  const float invL = 1.0f / (float)L;
  int k = 0;
  for (long long f = blockIdx.x; f < g.frames; f += gridDim.x, k++) {
    const int st = k & 1;
    if (threadIdx.x == 0 && f + gridDim.x < g.frames) issue(f + gridDim.x, st ^ 1);   // (stage st^1 was released by the barrier that ended iteration k-1)
    const float rs = g.rstd[f], mu = g.mean[f];
    {                                                  // wait for this frame's rows
      const uint32_t bar = bar0 + 8u * (uint32_t)st, parity = (uint32_t)((k >> 1) & 1);
      uint32_t done = 0;
      while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    float* sdy = bsm + (size_t)st * 2 * L; float* sc = sdy + L;
    float s1 = 0.f, s2 = 0.f;
    // pass 1: dxhat = dy * lrelu'(u) * gamma, xhat; frame sums, dgamma / dbeta.  Nothing is written back: pass 2 recomputes
    // (6 more instructions per element against a quarter of the shared-memory traffic -- the kernel ran at 88 % L1/TEX
    // throughput, 63 % DRAM)
    for (int u = threadIdx.x; u < L8; u += blockDim.x) {
      float d[8], h[8];
      ld8(sdy + 8 * u, d); ld8(sc + 8 * u, h);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const float xh = (h[e] - mu) * rs;             // xhat, as the forward formed it
        const float uu = fmaf(xh, gm[e], bt[e]);
        const float du = d[e] * (uu >= 0.f ? 1.0f : 0.02f);
        const float ox = du * gm[e];
        s1 += ox; s2 = fmaf(ox, xh, s2);
        adg[e] = fmaf(du, xh, adg[e]); adb[e] += du;
      }
    }
    {                                                  // both frame sums behind ONE barrier (partials double-buffered by frame parity)
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      float* rp = red + (k & 1) * 16;
      if ((threadIdx.x & 31) == 0) { rp[2 * (threadIdx.x >> 5)] = s1; rp[2 * (threadIdx.x >> 5) + 1] = s2; }
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) { s1 += rp[2 * w]; s2 += rp[2 * w + 1]; }
      s1 *= invL; s2 *= invL;
    }
    // pass 2: dc = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) -> padded output frame
    for (int u = threadIdx.x; u < L8; u += blockDim.x) {
      float d[8], h[8], o[8];
      ld8(sdy + 8 * u, d); ld8(sc + 8 * u, h);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const float xh = (h[e] - mu) * rs;
        const float uu = fmaf(xh, gm[e], bt[e]);
        const float ox = d[e] * (uu >= 0.f ? 1.0f : 0.02f) * gm[e];
        o[e] = rs * (ox - s1 - xh * s2); adc[e] += o[e];
      }
      st8(g.dc, f, g.out_flen, 8 * (u + off8), o, g.out_split);
    }
    zero_pads(g.dc, f, g.out_flen, off8, L8, F8, threadIdx.x, blockDim.x, g.out_split);
    __syncthreads();                                   // every thread is done with stage st
  }
  {
    const int P = lane_period(g.Cn, 8); const bool owner = (int)(threadIdx.x & 31) < P;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const float a = lane_period_sum(adg[e], P), b = lane_period_sum(adb[e], P), c = lane_period_sum(adc[e], P);
      if (owner) { atomicAdd(&chs[c0 + e], a); atomicAdd(&chs[g.Cn + c0 + e], b); atomicAdd(&chs[2 * g.Cn + c0 + e], c); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.Cn; i += blockDim.x) {
    atomicAdd(&g.dgamma[i], chs[i]); atomicAdd(&g.dbeta[i], chs[g.Cn + i]); atomicAdd(&g.dbias[i], chs[2 * g.Cn + i]);
  }
}

// =============================================================================================
// The N(0,1) draw of GaussianSampleLayer (util/layers.py:154, tf.random_normal) made in-kernel: counter-based
// Philox4x32-10 (Salmon et al., SC'11 -- the generator family TF's RandomStandardNormal uses; TF's own stream is
// not reproducible) keyed by the caller's seed, counter = (dim, frame index, pass counter), Box-Muller on two of
// the four output words.  A draw depends only on (seed, pass, frame, dim): identical for any chunking, launch
// geometry or rank layout, and the backward regenerates it instead of reading 512 B / frame back.
// =============================================================================================
struct StepState { unsigned long long seed; long long draws; long long step; long long reserved; };   // == npvc_step_state

__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long draws, unsigned long long frame, uint32_t dim) {
  uint32_t c0 = dim, c1 = (uint32_t)frame, c2 = (uint32_t)(frame >> 32), c3 = (uint32_t)draws;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(draws >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u1 = (float)((c0 >> 8) + 1u) * 5.9604644775390625e-8f;     // (0, 1], 24 bits
  const float u2 = (float)(c1 >> 8) * 5.9604644775390625e-8f;            // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// =============================================================================================
// sampler + KL  (util/layers.py:152-156,170-183); blockDim = 2z threads, thread = column
// eps != nullptr: the caller's draw; else st != nullptr: in-kernel Philox (frame index = frame0 + f); else no sampling
// =============================================================================================
__global__ void sample_kl_kernel(const float* hz, const float* eps, const StepState* st, long long frame0, float* mu, float* lv, float* zout,
                                 double* acc_kl, int z, long long frames, int frames_per_block) {
  pdl_prologue();
  __shared__ float red[40];
  const int col = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * frames_per_block;
  float kl = 0.f;
  for (int i = 0; i < frames_per_block; i++) {
    long long f = f0 + i; if (f >= frames) break;
    float v = hz[f * 2 * z + col];
    if (col < z) { mu[f * z + col] = v; }
    else {
      const int d = col - z;
      lv[f * z + d] = v;
      float m = hz[f * 2 * z + d];
      float ev = expf(v);
      if (eps) zout[f * z + d] = fmaf(eps[f * z + d], sqrtf(ev), m);
      else if (st) zout[f * z + d] = fmaf(philox_normal(st->seed, (unsigned long long)st->draws, (unsigned long long)(frame0 + f), (uint32_t)d), sqrtf(ev), m);
      kl += 0.5f * (-v + (ev + m * m) / NPVC_ONE_PLUS_EPS - 1.0f);
    }
  }
  if (acc_kl) { float t = block_sum(kl, red); if (threadIdx.x == 0) atomicAdd(acc_kl, (double)t); }
}
// the in-kernel draw alone (tests: statistics / reproducibility of the generator): out[n, z]
__global__ void philox_normal_kernel(const StepState* st, long long frame0, float* out, int z, long long n) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * z) out[i] = philox_normal(st->seed, (unsigned long long)st->draws, (unsigned long long)(frame0 + i / z), (uint32_t)(i % z));
}

// Zero pads of a split (bf16 hi / lo planes) per-frame buffer whose interior [i0, i0 + i1) is written by the producing
// GEMM: only the pad elements [0, i0) and [i0 + i1, flen) of both planes are cleared (all multiples of 8 elements).
__global__ void zero_pads_kernel(uint16_t* buf, int flen, int i0, int i1, long long frames) {
  pdl_prologue();
  const int front8 = i0 >> 3, back8 = (flen - i0 - i1) >> 3, per = 2 * (front8 + back8);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * per) return;
  const long long f = i / per; int q = (int)(i - f * per);
  const int plane = q / (front8 + back8); q -= plane * (front8 + back8);
  const int e = (q < front8) ? 8 * q : i0 + i1 + 8 * (q - front8);
  *reinterpret_cast<uint4*>(buf + f * 2 * flen + (long long)plane * flen + e) = make_uint4(0u, 0u, 0u, 0u);
}

// GaussianSampleLayer alone (util/layers.py:152-156)
__global__ void sample_only_kernel(const float* mu, const float* lv, const float* eps, float* z, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = fmaf(eps[i], sqrtf(expf(lv[i])), mu[i]);
}

// dz, eps, (mu|lv) -> (dmu|dlv); column sums -> head-bias grads.  inv_n = 1 / (frames the means span)
__global__ void sample_bwd_kernel(const float* dz, const float* eps, const StepState* st, long long frame0, const float* hz, float* dhz, float* dbh,
                                  int z, long long frames, int frames_per_block, float inv_n, int out_split) {
  pdl_prologue();
  const int col = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * frames_per_block;
  float cs = 0.f;
  for (int i = 0; i < frames_per_block; i++) {
    long long f = f0 + i; if (f >= frames) break;
    float o;
    if (col < z) {
      float m = hz[f * 2 * z + col];
      o = dz[f * z + col] + m / NPVC_ONE_PLUS_EPS * inv_n;
    } else {
      const int d = col - z;
      float l = hz[f * 2 * z + col];
      float ev = expf(l);
      const float e = eps ? eps[f * z + d] : philox_normal(st->seed, (unsigned long long)st->draws, (unsigned long long)(frame0 + f), (uint32_t)d);
      o = dz[f * z + d] * e * 0.5f * sqrtf(ev) + 0.5f * (ev / NPVC_ONE_PLUS_EPS - 1.0f) * inv_n;
    }
    if (out_split) split_st1(reinterpret_cast<uint16_t*>(dhz) + f * 4 * z + col, 2 * z, o);
    else dhz[f * 2 * z + col] = o;
    cs += o;
  }
  atomicAdd(&dbh[col], cs);
}

// =============================================================================================
// Gaussian log-density + d/dxh  (util/layers.py:159-167): one warp per frame
// =============================================================================================
__global__ void recon_kernel(const float* x, const float* xh, float* dxh, float* dbias, double* acc_logp,
                             int H, int ld, int Co, long long frames, float inv_n, int out_split) {
  pdl_prologue();
  __shared__ float red[40];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long f = (long long)blockIdx.x * nw + w;
  float lp = 0.f, db = 0.f;
  if (f < frames) {
    const float* xr = x + f * H; const float* hr = xh + f * ld;       // (xh rows at the pitch of dxh)
    for (int i0 = lane; i0 < ld; i0 += 128) {          // 4 x 2 loads in flight per lane (same element order per lane as a plain loop)
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = i0 + 32 * u; const bool ok = i < H; a[u] = ok ? hr[i] : 0.f; b[u] = ok ? xr[i] : 0.f; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + 32 * u;
        if (i >= ld) break;
        float g = 0.f;
        if (i < H) {
          float d = a[u] - b[u];
          lp += -0.5f * (NPVC_LOG_2PI + d * d / NPVC_ONE_PLUS_EPS);
          g = d / NPVC_ONE_PLUS_EPS * inv_n;
          db += g;
        }
        if (dxh) {
          if (out_split) split_st1(reinterpret_cast<uint16_t*>(dxh) + f * 2 * ld + i, ld, g);
          else dxh[f * ld + i] = g;
        }
      }
    }
  }
  float t = block_sum(lp, red);
  if (threadIdx.x == 0) atomicAdd(acc_logp, (double)t);
  if (dbias) {   // Co == 1 on this path (single output channel)
    float s = block_sum(db, red);
    if (threadIdx.x == 0) atomicAdd(dbias, s);
  }
}

__global__ void colsum_kernel(const float* in, float* out, int N, int rows) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int r = 0; r < rows; r++) s += in[(long long)r * N + c];
  out[c] += s;
}

// =============================================================================================
// pack / unpack / adam / finalize / tanhize / records
// =============================================================================================
__global__ void pack_kernel(const float* theta, const int* src, float* arena, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int s = src[i]; arena[i] = (s >= 0) ? theta[s] : 0.f; }
}

// bf16 operand packs of the tensor path (plan.h): src = index | flags; bit 29: the source is the fp32 pack
// arena[index] instead of theta[index].  Entry i -> hi pack element i and lo pack element i + n (one gather for both).
__global__ void pack16_kernel(const float* theta, const float* arena, const int* src, uint16_t* arena16, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int s = src[i];
    uint32_t hi = 0u, lo = 0u;
    if (s >= 0) {
      const int idx = s & ((1 << 29) - 1);
      const float v = (s & (1 << 29)) ? arena[idx] : theta[idx];
      hi = split_pack2(v, 0.f, lo);
    }
    arena16[i] = (uint16_t)(hi & 0xffffu); arena16[i + n] = (uint16_t)(lo & 0xffffu);
  }
}
// fp32 operand packs, only the positions some op reads (plan.h, Plan::pack_list)
__global__ void pack_list_kernel(const float* theta, const int* src, const int* list, float* arena, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int q = list[i]; const int s = src[q]; arena[q] = (s >= 0) ? theta[s] : 0.f; }
}

// zs[f] = [z[f] (zd floats) | one-hot(y[f]) (yp floats)], fp32 or split planes (zd, yp multiples of 4):
// the merge GEMM's A operand (model/vae.py:64-70,89-90 -- embedding lookup folded into the GEMM)
__global__ void zcat_kernel(const float* z, const long long* y, float* out, int zd, int yp, long long frames, int out_split) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // float4 index
  const int W4 = (zd + yp) >> 2;
  if (i >= frames * W4) return;
  const long long f = i / W4; const int q = (int)(i - f * W4);
  float4 v;
  if (4 * q < zd) v = reinterpret_cast<const float4*>(z + f * zd)[q];
  else {
    const int s = (int)y[f] - (4 * q - zd);
    v = make_float4(s == 0 ? 1.f : 0.f, s == 1 ? 1.f : 0.f, s == 2 ? 1.f : 0.f, s == 3 ? 1.f : 0.f);
  }
  const int W = zd + yp;
  if (out_split) split_st4(reinterpret_cast<uint16_t*>(out) + f * 2 * W + 4 * q, W, v);
  else reinterpret_cast<float4*>(out + f * W)[q] = v;
}

// grad[t] += sum of the packed-gradient entries of parameter t (CSR).  Parameters gathered from many
// positions (the 1025-tap kernel: one Toeplitz diagonal of up to 513 entries each) are left to
// unpack_heavy_kernel (one warp per parameter).
constexpr int UNPACK_HEAVY = 32;
__global__ void unpack_kernel(const float* adw, const int* ptr, const int* idx, float* grad, long long n) {
  pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int b = ptr[t], e = ptr[t + 1];
  if (b == e || e - b >= UNPACK_HEAVY) return;
  float s = 0.f;
  for (int i = b; i < e; i++) s += adw[idx[i]];
  grad[t] += s;
}
__global__ void unpack_heavy_kernel(const float* adw, const int* ptr, const int* idx, const int* heavy, int n_heavy, float* grad) {
  pdl_prologue();
  const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n_heavy) return;
  const int t = heavy[w];
  const int b = ptr[t], e = ptr[t + 1];
  float s = 0.f;
  for (int i = b + lane; i < e; i += 32) s += adw[idx[i]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) grad[t] += s;
}

// TF-form Adam (trainer/vae.py:16-24): theta -= lr_t * m / (sqrt(v) + eps)
// st != nullptr: the step t lives on the device (StepState::step, advanced by the fwd+bwd pass) and
// lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is formed here -- nothing the host computes changes between steps, so a
// captured CUDA graph of the whole training step replays unchanged.  lr_t then carries the base rate lr.
__global__ void adam_kernel(float* theta, const float* grad, float* m, float* v, long long n,
                            float lr_t, float b1, float b2, float eps, float gscale, const StepState* st) {
  pdl_prologue();
  if (st) {
    __shared__ float s_lr;
    if (threadIdx.x == 0) {
      const double t = (double)st->step;
      s_lr = (float)((double)lr_t * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    }
    __syncthreads();
    lr_t = s_lr;
  }
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = grad[i] * gscale;
  float mm = b1 * m[i] + (1.0f - b1) * g;
  float vv = b2 * v[i] + (1.0f - b2) * g * g;
  m[i] = mm; v[i] = vv;
  theta[i] -= lr_t * mm / (sqrtf(vv) + eps);
}

// losses = {G, D_KL, logP}; st != nullptr: the pass counter of the in-kernel sampler advances, and -- after a pass
// that produced a gradient -- the step counter the device-side Adam reads
__global__ void finalize_losses_kernel(const double* acc, float* losses, double inv_n, StepState* st, int had_grad) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double kl = acc[0] * inv_n, lp = acc[1] * inv_n;
    if (losses) { losses[0] = (float)(-lp + kl); losses[1] = (float)kl; losses[2] = (float)lp; }
    if (st) { st->draws += 1; if (had_grad) st->step += 1; }
  }
}

__global__ void tanhize_fwd_kernel(const float* x, const float* xmin, const float* xmax, float* out, long long n, int dim) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dim) return;
  int d = (int)(i % dim);
  float t = (x[i] - xmin[d]) / (xmax[d] - xmin[d]);
  out[i] = fminf(fmaxf(t, 0.f), 1.f) * 2.f - 1.f;
}
__global__ void tanhize_bwd_kernel(const float* x, const float* xmin, const float* xmax, float* out, long long n, int dim) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dim) return;
  int d = (int)(i % dim);
  out[i] = (x[i] * 0.5f + 0.5f) * (xmax[d] - xmin[d]) + xmin[d];
}
// analyzer.py:111-127: record = [sp(513) | ap(513) | f0 | en | spk]; feature = Tanhize(sp), speaker = int64(last)
__global__ void unpack_records_kernel(const float* rec, long long n, int rec_floats, int sp_dim,
                                      const float* xmin, const float* xmax, float* x, long long* y) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * sp_dim) return;
  long long f = i / sp_dim; int d = (int)(i - f * sp_dim);
  float v = rec[f * rec_floats + d];
  if (xmin) { float t = (v - xmin[d]) / (xmax[d] - xmin[d]); v = fminf(fmaxf(t, 0.f), 1.f) * 2.f - 1.f; }
  x[i] = v;
  if (d == 0) y[f] = (long long)rec[f * rec_floats + rec_floats - 1];
}

}  // namespace npvc
