// First encoder layer, fused (sm_100a, CUDA cores): the layer reads the caller's 513-bin frame (one input channel,
// k <= 8 taps -- K = 7 is no GEMM) and produces the largest activation of the encoder (171 x 16 floats), so it is
// pure HBM streaming.  Two kernels replace four launches of the layer-wise plan:
//
//   e0_fwd_kernel   conv (tf.layers.conv2d SAME, util/layers.py:56-64) + bias + Layernorm + lrelu
//                   (util/layers.py:10-44,147-149) in registers: x is read once (2 KB / frame), the raw conv output is
//                   written only for training (the backward recomputes xhat from it), the activation goes out as the
//                   zero-padded bf16 hi / lo planes the next layer's TMA boxes read.
//   e0_bwd_kernel   Layernorm + lrelu backward AND the layer's weight gradient: dc (the gradient w.r.t. the conv
//                   output) never leaves the registers -- the first layer has no data gradient, so nothing else
//                   reads it -- and dW[k][c] += x[s j + k - pl] * dc[j][c] is accumulated per thread.
//
// Thread mapping as in the register-resident Layernorm kernels (kernels.cuh): G threads own one frame, a thread's
// units are U consecutive channels of one position (U = 8 forward: 16-byte plane stores; U = 4 backward: float4
// loads only), unit u = t + i * G, so a thread's channels are fixed (U * G is a multiple of Co).
#pragma once
#include "kernels.cuh"

namespace npvc {

struct E0FwdArgs {
  const float* x; const float* W; const float* bias; const float* gamma; const float* beta;
  float* c;                     // raw conv output [frames][Ho * Co] (nullptr: inference, not kept)
  float* mean; float* rstd; float* aout;
  int Hi, Ho, Co, k, s, pl, out_flen, out_off, out_split;
  int xp;                       // floats of a staged input row: pl zeros | Hi inputs | zeros up to s (Ho - 1) + KT, rounded up to 4
  long long frames;
};
struct E0BwdArgs {
  const float* x; const float* dy; const float* cin; const float* mean; const float* rstd;
  const float* gamma; const float* beta;
  float* dW;                    // [k][Co] packed weight gradient (accumulated)
  float* dgamma; float* dbeta; float* dbias;
  int Hi, Ho, Co, k, s, pl, xp; long long frames;
  int prefetch;                 // dy / c of the next frame land in shared memory while this one is computed (LnBwdArgs::prefetch)
};

constexpr int E0_KT = 8;        // taps per position (weights beyond k are zero)
// threads per block: a frame's G threads (>= 128) are a block of their own, so the per-frame barriers span nothing else
#define E0_BLOCK(G) ((G) >= 128 ? (G) : 128)
__host__ __device__ inline int e0_row_floats(int Ho, int s, int pl, int Hi) {
  int need = s * (Ho - 1) + E0_KT; if (need < pl + Hi) need = pl + Hi;
  return (need + 3) / 4 * 4;
}

// the frame's input row with its SAME padding (zeros) into shared memory - the taps read it without predicates or
// address math -, asynchronously: one commit group per call (empty when !fok)
__device__ __forceinline__ void e0_fetch_x(float* xs, const float* xp, bool fok, int t, int G, int XP, int pl, int Hi) {
  if (fok)
    for (int i = t; i < XP; i += G) { const int xi = i - pl; const bool ok = xi >= 0 && xi < Hi; cp_async4_zfill(xs + i, ok ? xp + xi : xp, ok ? 4 : 0); }
  cp_async_commit();
}

// G threads per frame, V units of 8 consecutive channels per thread (L <= 8 G V); the next frame's input row
// is in flight while this one is computed (three row buffers: the one written in iteration k was last read in k - 2, and
// every thread has passed the barriers of k - 1 since)
template <int G, int V>
__global__ void __launch_bounds__(E0_BLOCK(G), 768 / E0_BLOCK(G)) e0_fwd_kernel(E0FwdArgs g) {
  pdl_prologue();
  constexpr int FPB = E0_BLOCK(G) / G;
  extern __shared__ __align__(16) float e0sm[];      // [KT][Co] weights | bias | gamma | beta | [3][FPB][xp] staged frames
  __shared__ float red[2 * 8];
  int par = 0;
  const int Co = g.Co, XP = g.xp;
  float* sw = e0sm; float* sb = sw + E0_KT * Co; float* sg = sb + Co; float* sbt = sg + Co;
  for (int i = threadIdx.x; i < E0_KT * Co; i += blockDim.x) sw[i] = (i < g.k * Co) ? g.W[i] : 0.f;
  for (int i = threadIdx.x; i < Co; i += blockDim.x) { sb[i] = g.bias[i]; sg[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  float* xs0 = sbt + Co + grp * XP;
  int xb = 0;
  const int cpb = Co >> 3;                             // 8-channel blocks per position (a power of two: divides G)
  const int cshift = 31 - __clz(cpb);
  const int c0 = (t & (cpb - 1)) << 3;                 // this thread's channels
  const int L = g.Ho * Co, L8 = L >> 3, off8 = g.out_off >> 3, F8 = g.out_flen >> 3;
  const float invL = 1.0f / (float)L;
  const float2* w2 = reinterpret_cast<const float2*>(sw + c0);
  { const long long f0 = (long long)blockIdx.x * FPB + grp; e0_fetch_x(xs0, g.x + f0 * g.Hi, f0 < g.frames, t, G, XP, g.pl, g.Hi); }
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    const float* xs = xs0 + xb * FPB * XP;
    xb = xb == 2 ? 0 : xb + 1;
    { const long long fn = (fb + gridDim.x) * FPB + grp; e0_fetch_x(xs0 + xb * FPB * XP, g.x + fn * g.Hi, fn < g.frames, t, G, XP, g.pl, g.Hi); }
    cp_async_wait<1>();
    __syncthreads();                                   // this frame's row (and, first pass, the weights) published
    float2 v[V][4];
    float s[1] = {0.f};
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (u < L8) {
        const float* xw = xs + g.s * (u >> cshift);   // window of output position u / cpb
#pragma unroll
        for (int q = 0; q < 4; q++) v[i][q] = *reinterpret_cast<const float2*>(sb + c0 + 2 * q);
#pragma unroll
        for (int kk = 0; kk < E0_KT; kk++) {
          const float xv = xw[kk]; const float2 x2 = make_float2(xv, xv);
#pragma unroll
          for (int q = 0; q < 4; q++) v[i][q] = __ffma2_rn(x2, w2[(kk * Co >> 1) + q], v[i][q]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) s[0] += v[i][q].x + v[i][q].y;
      }
    }
    group_sum_db<G, 1>(s, red, par);
    const float mean = s[0] * invL;
    float q2[1] = {0.f};
#pragma unroll
    for (int i = 0; i < V; i++) {
      if (t + i * G < L8) {
#pragma unroll
        for (int q = 0; q < 4; q++) { const float d0 = v[i][q].x - mean, d1 = v[i][q].y - mean; q2[0] = fmaf(d0, d0, q2[0]); q2[0] = fmaf(d1, d1, q2[0]); }
      }
    }
    group_sum_db<G, 1>(q2, red, par);
    const float rs = rsqrtf(q2[0] * invL + NPVC_LN_EPS);
    if (!fok) continue;                    // (every thread still reaches the barriers at the top of the next iteration)
    if (t == 0) { g.rstd[f] = rs; g.mean[f] = mean; }
    float gm[8], bt[8];
    { const float4 a = *reinterpret_cast<const float4*>(sg + c0), b = *reinterpret_cast<const float4*>(sg + c0 + 4);
      const float4 c = *reinterpret_cast<const float4*>(sbt + c0), d = *reinterpret_cast<const float4*>(sbt + c0 + 4);
      gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w; gm[4] = b.x; gm[5] = b.y; gm[6] = b.z; gm[7] = b.w;
      bt[0] = c.x; bt[1] = c.y; bt[2] = c.z; bt[3] = c.w; bt[4] = d.x; bt[5] = d.y; bt[6] = d.z; bt[7] = d.w; }
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (u < L8) {
        const float r[8] = {v[i][0].x, v[i][0].y, v[i][1].x, v[i][1].y, v[i][2].x, v[i][2].y, v[i][3].x, v[i][3].y};
        if (g.c) st8(g.c, f, L, 8 * u, r, 0);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = lrelu_f(fmaf((r[e] - mean) * rs, gm[e], bt[e]));
        st8(g.aout, f, g.out_flen, 8 * (u + off8), o, g.out_split);
      }
    }
    zero_pads(g.aout, f, g.out_flen, off8, L8, F8, t, G, g.out_split);
  }
}

// floats of dynamic shared memory of e0_bwd_kernel (FPB frames per block)
constexpr int E0_PF = 1;        // frames in flight ahead of the one being computed (backward); 2 measured no faster
__host__ __device__ inline size_t e0_bwd_smem_floats(int Co, int FPB, int xp, int L, int prefetch) {
  return (size_t)(E0_KT + 5) * Co + (size_t)(E0_PF + 2) * FPB * xp + (prefetch ? (size_t)(E0_PF + 1) * FPB * 2 * L : (size_t)0);
}

// G threads per frame, V units of 4 consecutive channels per thread (L <= 4 G V).  A block's next E0_PF frames (dy, the
// raw conv output, the input row with its SAME padding, mean / rstd) are in flight while the current one is computed
// : dy / c land in a per-thread staging ring (a thread copies
// exactly the units it will read: no barrier), the input row - read by the whole group - in a ring one buffer longer.
template <int G, int V>
__global__ void __launch_bounds__(E0_BLOCK(G), 512 / E0_BLOCK(G)) e0_bwd_kernel(E0BwdArgs g) {
  pdl_prologue();
  constexpr int FPB = E0_BLOCK(G) / G;
  extern __shared__ __align__(16) float e0sm[];      // [3 Co] dgamma | dbeta | dbias sums, [KT Co] dW sums, gamma, beta, [PF + 2][FPB][xp] rows, [PF + 1][FPB][2][L] dy | c
  __shared__ float red[2][E0_BLOCK(G) / 32][2];
  const int Co = g.Co, XP = g.xp;
  float* chs = e0sm; float* sdw = chs + 3 * Co; float* sgm = sdw + E0_KT * Co; float* sbt = sgm + Co;
  for (int i = threadIdx.x; i < (3 + E0_KT) * Co; i += blockDim.x) e0sm[i] = 0.f;
  for (int i = threadIdx.x; i < Co; i += blockDim.x) { sgm[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  __syncthreads();
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  float* xs0 = sbt + Co + grp * XP;
  const int qpp = Co >> 2;                             // channel quads per position (a power of two: divides G)
  const int qshift = 31 - __clz(qpp);
  const int c0 = (t & (qpp - 1)) << 2;
  const int L = g.Ho * Co, L4 = L >> 2;
  const float invL = 1.0f / (float)L;
  float* stg0 = sbt + Co + (E0_PF + 2) * FPB * XP + (size_t)grp * 2 * L;        // this frame slot's [dy | c], + buf * FPB * 2 * L
  float gm[4], bt[4], adg[4], adb[4], adc[4];
  float2 dw[E0_KT][2];
#pragma unroll
  for (int e = 0; e < 4; e++) { gm[e] = sgm[c0 + e]; bt[e] = sbt[c0 + e]; adg[e] = adb[e] = adc[e] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < E0_KT; kk++) dw[kk][0] = dw[kk][1] = make_float2(0.f, 0.f);
  const bool pf = g.prefetch != 0;                     // (else only the input row travels ahead; dy / c are read in place)
  // one commit group per call, empty past the end: group k holds frame-block k of this block
  auto fetch = [&](long long fbn, int buf, int xb, float& rs_n, float& mu_n) {
    const long long fn = fbn * FPB + grp;
    if (fn < g.frames) {
      if (pf) {
        float4* sd = reinterpret_cast<float4*>(stg0 + (size_t)buf * FPB * 2 * L) + t;
        const float4* dyp = reinterpret_cast<const float4*>(g.dy + fn * L) + t;
        const float4* cp = reinterpret_cast<const float4*>(g.cin + fn * L) + t;
#pragma unroll
        for (int i = 0; i < V; i++)
          if (t + i * G < L4) { cp_async16(sd + i * G, dyp + i * G); cp_async16(sd + L4 + i * G, cp + i * G); }
      }
      float* xs = xs0 + xb * FPB * XP; const float* xp = g.x + fn * g.Hi;
      for (int i = t; i < XP; i += G) { const int xi = i - g.pl; const bool ok = xi >= 0 && xi < g.Hi; cp_async4_zfill(xs + i, ok ? xp + xi : xp, ok ? 4 : 0); }
      rs_n = g.rstd[fn]; mu_n = g.mean[fn];
    }
    cp_async_commit();
  };
  int par = 0, sb = 0, xb = 0;                          // parity of the warp partials, staging / row ring positions of this frame
  float rs_n[E0_PF], mu_n[E0_PF];
#pragma unroll
  for (int d = 0; d < E0_PF; d++) { rs_n[d] = mu_n[d] = 0.f; fetch(blockIdx.x + (long long)d * gridDim.x, d, d, rs_n[d], mu_n[d]); }
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    float dx[V][4], xh[V][4];
    const float rs = rs_n[0], mu = mu_n[0];
#pragma unroll
    for (int d = 0; d + 1 < E0_PF; d++) { rs_n[d] = rs_n[d + 1]; mu_n[d] = mu_n[d + 1]; }
    // the staging buffer written here is the one this thread read in the previous iteration; the row buffer was last read
    // two iterations ago, and every thread has since passed a barrier behind those reads (the previous iteration's)
    fetch(fb + (long long)E0_PF * gridDim.x, (sb + E0_PF) % (E0_PF + 1), (xb + E0_PF) % (E0_PF + 2), rs_n[E0_PF - 1], mu_n[E0_PF - 1]);
    cp_async_wait<E0_PF>();                            // this frame's group has landed (own copies: visible to this thread)
    {
      const float4* sd = pf ? reinterpret_cast<const float4*>(stg0 + (size_t)sb * FPB * 2 * L) + t : reinterpret_cast<const float4*>(g.dy + f * L) + t;
      const float4* sc = pf ? sd + L4 : reinterpret_cast<const float4*>(g.cin + f * L) + t;
#pragma unroll
      for (int i = 0; i < V; i++) {
        if (fok && t + i * G < L4) {
          const float4 a = sd[i * G], b = sc[i * G];
          dx[i][0] = a.x; dx[i][1] = a.y; dx[i][2] = a.z; dx[i][3] = a.w; xh[i][0] = b.x; xh[i][1] = b.y; xh[i][2] = b.z; xh[i][3] = b.w;
        }
      }
    }
    float* xs = xs0 + xb * FPB * XP;                   // (published to the group by the frame's barrier below)
    sb = (sb + 1) % (E0_PF + 1); xb = (xb + 1) % (E0_PF + 2);
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < V; i++) {
      if (fok && t + i * G < L4) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float h = (xh[i][e] - mu) * rs;                  // xhat, as the forward formed it
          const float u = fmaf(h, gm[e], bt[e]);
          const float du = dx[i][e] * (u >= 0.f ? 1.0f : 0.02f);
          const float ox = du * gm[e];
          s[0] += ox; s[1] = fmaf(ox, h, s[1]);
          adg[e] = fmaf(du, h, adg[e]); adb[e] += du;
          dx[i][e] = ox; xh[i][e] = h;
        }
      }
    }
    {
      // frame sums over the group's warps (partials double-buffered by parity: written in k, next written in k + 2, and
      // every thread passes the barrier of k + 1 after its reads of k); the barrier also publishes the input row.
      s[0] = warp_sum(s[0]); s[1] = warp_sum(s[1]);
      const int warp = threadIdx.x >> 5;
      if ((threadIdx.x & 31) == 0) { red[par][warp][0] = s[0]; red[par][warp][1] = s[1]; }
      __syncthreads();
      const int w0 = (warp / (G / 32)) * (G / 32);
      s[0] = 0.f; s[1] = 0.f;
#pragma unroll
      for (int w = 0; w < G / 32; w++) { s[0] += red[par][w0 + w][0]; s[1] += red[par][w0 + w][1]; }
    }
    par ^= 1;
    if (!fok) continue;
    const float s1 = s[0] * invL, s2 = s[1] * invL;
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (u < L4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { o[e] = rs * (dx[i][e] - s1 - xh[i][e] * s2); adc[e] += o[e]; }
        const float2 o01 = make_float2(o[0], o[1]), o23 = make_float2(o[2], o[3]);
        const float* xw = xs + g.s * (u >> qshift);
#pragma unroll
        for (int kk = 0; kk < E0_KT; kk++) {
          const float xv = xw[kk]; const float2 x2 = make_float2(xv, xv);
          dw[kk][0] = __ffma2_rn(x2, o01, dw[kk][0]); dw[kk][1] = __ffma2_rn(x2, o23, dw[kk][1]);
        }
      }
    }
  }
  __syncthreads();
  {
    const int P = lane_period(Co, 4); const bool owner = (int)(threadIdx.x & 31) < P;    // lanes P apart hold the same channel quad
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float a = lane_period_sum(adg[e], P), b = lane_period_sum(adb[e], P), c = lane_period_sum(adc[e], P);
      if (owner) { atomicAdd(&chs[c0 + e], a); atomicAdd(&chs[Co + c0 + e], b); atomicAdd(&chs[2 * Co + c0 + e], c); }
    }
#pragma unroll
    for (int kk = 0; kk < E0_KT; kk++) {
      const float w0 = lane_period_sum(dw[kk][0].x, P), w1 = lane_period_sum(dw[kk][0].y, P);
      const float w2 = lane_period_sum(dw[kk][1].x, P), w3 = lane_period_sum(dw[kk][1].y, P);
      if (owner) { atomicAdd(&sdw[kk * Co + c0 + 0], w0); atomicAdd(&sdw[kk * Co + c0 + 1], w1); atomicAdd(&sdw[kk * Co + c0 + 2], w2); atomicAdd(&sdw[kk * Co + c0 + 3], w3); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Co; i += blockDim.x) {
    atomicAdd(&g.dgamma[i], chs[i]); atomicAdd(&g.dbeta[i], chs[Co + i]); atomicAdd(&g.dbias[i], chs[2 * Co + i]);
  }
  for (int i = threadIdx.x; i < g.k * Co; i += blockDim.x) atomicAdd(&g.dW[i], sdw[i]);
}

}  // namespace npvc
