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
  int Hi, Ho, Co, k, s, pl, out_flen, out_off, out_split; long long frames;
};
struct E0BwdArgs {
  const float* x; const float* dy; const float* cin; const float* mean; const float* rstd;
  const float* gamma; const float* beta;
  float* dW;                    // [k][Co] packed weight gradient (accumulated)
  float* dgamma; float* dbeta; float* dbias;
  int Hi, Ho, Co, k, s, pl; long long frames;
};

constexpr int E0_KT = 8;        // taps held per thread (weights beyond k are zero)

template <int G>
__global__ void __launch_bounds__(256) e0_fwd_kernel(E0FwdArgs g) {
  constexpr int V = 4, FPB = 256 / G;
  extern __shared__ __align__(16) float e0sm[];      // [KT][Co] weights | bias | gamma | beta
  __shared__ float red[8];
  const int Co = g.Co;
  float* sw = e0sm; float* sb = sw + E0_KT * Co; float* sg = sb + Co; float* sbt = sg + Co;
  for (int i = threadIdx.x; i < E0_KT * Co; i += blockDim.x) sw[i] = (i < g.k * Co) ? g.W[i] : 0.f;
  for (int i = threadIdx.x; i < Co; i += blockDim.x) { sb[i] = g.bias[i]; sg[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  __syncthreads();
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  const int cpb = Co >> 3;                             // 8-channel blocks per position (a power of two: divides G)
  const int cshift = 31 - __clz(cpb);
  const int c0 = (t & (cpb - 1)) << 3;                 // this thread's channels
  const int L = g.Ho * Co, L8 = L >> 3, off8 = g.out_off >> 3, F8 = g.out_flen >> 3;
  const float invL = 1.0f / (float)L;
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    const float* xp = g.x + f * g.Hi;
    float v[V][8];
    float s[1] = {0.f};
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (fok && u < L8) {
        const int i0 = g.s * (u >> cshift) - g.pl;     // first input position of the window (SAME padding: zeros outside)
        { const float4 a = *reinterpret_cast<const float4*>(sb + c0), b = *reinterpret_cast<const float4*>(sb + c0 + 4);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w; }
#pragma unroll
        for (int kk = 0; kk < E0_KT; kk++) {
          const int xi = i0 + kk;
          const float xv = (kk < g.k && xi >= 0 && xi < g.Hi) ? __ldg(xp + xi) : 0.f;
          const float4 a = *reinterpret_cast<const float4*>(sw + kk * Co + c0), b = *reinterpret_cast<const float4*>(sw + kk * Co + c0 + 4);
          v[i][0] = fmaf(xv, a.x, v[i][0]); v[i][1] = fmaf(xv, a.y, v[i][1]); v[i][2] = fmaf(xv, a.z, v[i][2]); v[i][3] = fmaf(xv, a.w, v[i][3]);
          v[i][4] = fmaf(xv, b.x, v[i][4]); v[i][5] = fmaf(xv, b.y, v[i][5]); v[i][6] = fmaf(xv, b.z, v[i][6]); v[i][7] = fmaf(xv, b.w, v[i][7]);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) s[0] += v[i][e];
      }
    }
    group_sum<G, 1>(s, red);
    const float mean = s[0] * invL;
    float q[1] = {0.f};
#pragma unroll
    for (int i = 0; i < V; i++) {
      if (fok && t + i * G < L8) {
#pragma unroll
        for (int e = 0; e < 8; e++) { const float d = v[i][e] - mean; q[0] = fmaf(d, d, q[0]); }
      }
    }
    group_sum<G, 1>(q, red);
    const float rs = rsqrtf(q[0] * invL + NPVC_LN_EPS);
    if (!fok) continue;                    // (no block-wide barrier after this point in the iteration)
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
        if (g.c) st8(g.c, f, L, 8 * u, v[i], 0);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = lrelu_f(fmaf((v[i][e] - mean) * rs, gm[e], bt[e]));
        st8(g.aout, f, g.out_flen, 8 * (u + off8), o, g.out_split);
      }
    }
    zero_pads(g.aout, f, g.out_flen, off8, L8, F8, t, G, g.out_split);
  }
}

template <int G>
__global__ void __launch_bounds__(256, 2) e0_bwd_kernel(E0BwdArgs g) {
  constexpr int V = 4, FPB = 256 / G;
  extern __shared__ __align__(16) float e0sm[];      // [3 Co] dgamma | dbeta | dbias sums, [KT Co] dW sums, gamma, beta
  __shared__ float red[16];
  const int Co = g.Co;
  float* chs = e0sm; float* sdw = chs + 3 * Co; float* sgm = sdw + E0_KT * Co; float* sbt = sgm + Co;
  for (int i = threadIdx.x; i < (3 + E0_KT) * Co; i += blockDim.x) e0sm[i] = 0.f;
  for (int i = threadIdx.x; i < Co; i += blockDim.x) { sgm[i] = g.gamma[i]; sbt[i] = g.beta[i]; }
  __syncthreads();
  const int t = threadIdx.x % G, grp = threadIdx.x / G;
  const int qpp = Co >> 2;                             // channel quads per position (a power of two: divides G)
  const int qshift = 31 - __clz(qpp);
  const int c0 = (t & (qpp - 1)) << 2;
  const int L = g.Ho * Co, L4 = L >> 2;
  const float invL = 1.0f / (float)L;
  float gm[4], bt[4], adg[4], adb[4], adc[4], dw[E0_KT][4];
#pragma unroll
  for (int e = 0; e < 4; e++) { gm[e] = sgm[c0 + e]; bt[e] = sbt[c0 + e]; adg[e] = adb[e] = adc[e] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < E0_KT; kk++)
#pragma unroll
    for (int e = 0; e < 4; e++) dw[kk][e] = 0.f;
  for (long long fb = blockIdx.x; fb * FPB < g.frames; fb += gridDim.x) {
    const long long f = fb * FPB + grp; const bool fok = f < g.frames;
    float dx[V][4], xh[V][4];
    float rs = 0.f, mu = 0.f;
    if (fok) { rs = g.rstd[f]; mu = g.mean[f]; }
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (fok && u < L4) {
        const float4 a = *reinterpret_cast<const float4*>(g.dy + f * L + 4 * u), b = *reinterpret_cast<const float4*>(g.cin + f * L + 4 * u);
        dx[i][0] = a.x; dx[i][1] = a.y; dx[i][2] = a.z; dx[i][3] = a.w; xh[i][0] = b.x; xh[i][1] = b.y; xh[i][2] = b.z; xh[i][3] = b.w;
      }
    }
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
    group_sum<G, 2>(s, red);
    if (!fok) continue;
    const float s1 = s[0] * invL, s2 = s[1] * invL;
    const float* xp = g.x + f * g.Hi;
#pragma unroll
    for (int i = 0; i < V; i++) {
      const int u = t + i * G;
      if (u < L4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { o[e] = rs * (dx[i][e] - s1 - xh[i][e] * s2); adc[e] += o[e]; }
        const int i0 = g.s * (u >> qshift) - g.pl;
#pragma unroll
        for (int kk = 0; kk < E0_KT; kk++) {
          const int xi = i0 + kk;
          const float xv = (kk < g.k && xi >= 0 && xi < g.Hi) ? __ldg(xp + xi) : 0.f;
#pragma unroll
          for (int e = 0; e < 4; e++) dw[kk][e] = fmaf(xv, o[e], dw[kk][e]);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; e++) {
    atomicAdd(&chs[c0 + e], adg[e]); atomicAdd(&chs[Co + c0 + e], adb[e]); atomicAdd(&chs[2 * Co + c0 + e], adc[e]);
  }
#pragma unroll
  for (int kk = 0; kk < E0_KT; kk++)
#pragma unroll
    for (int e = 0; e < 4; e++) atomicAdd(&sdw[kk * Co + c0 + e], dw[kk][e]);
  __syncthreads();
  for (int i = threadIdx.x; i < Co; i += blockDim.x) {
    atomicAdd(&g.dgamma[i], chs[i]); atomicAdd(&g.dbeta[i], chs[Co + i]); atomicAdd(&g.dbias[i], chs[2 * Co + i]);
  }
  for (int i = threadIdx.x; i < g.k * Co; i += blockDim.x) atomicAdd(&g.dW[i], sdw[i]);
}

}  // namespace npvc
