// Whole SwinTransformerBlock for the very narrow layers of the UpscalingHead (C = 24 and C = 12, 3 heads,
// 120 000 / 480 000 tokens per diffraction; SwinWNet.py:236-280 via :656-678) as ONE fp32 CUDA-core kernel:
//   x -> LN1 -> qkv -> 5x5 window attention (+rel-pos bias, optional shift mask) -> proj -> +x
//     -> LN2 -> fc1 -> GELU -> fc2 -> +  -> out
// With K = 12/24 a tcgen05 tile is >75 % padding and the four-kernel path is bound by per-tile latency, not by
// math: these blocks took 59 ms of a 203 ms step.  Here a thread owns a token; all weights (12 C^2 floats) sit
// in shared memory and are read as broadcast float4; a CTA of 128 threads handles 5 windows (125 tokens) per
// iteration and exchanges only k|v through shared memory; CTAs are persistent so the weights are loaded once.
// HBM traffic per token is one fp32 read and one fp32 write of the row (8 C bytes instead of 36 C).
#include "common.cuh"
#include "kernels.h"

namespace swn {

constexpr int SB_THREADS = 128;
constexpr int SB_WIN = 5;        // windows per CTA iteration
constexpr int SB_TOK = 25;

template <int C, int NH>
__global__ void __launch_bounds__(SB_THREADS) swin_block_small_kernel(const SmallBlockParams p, int nWy, int nWx,
                                                                       long long n_windows) {
  constexpr int HD = C / NH, C3 = 3 * C, C4 = 4 * C;
  extern __shared__ __align__(16) float sb_smem[];
  float* Wqkv_s = sb_smem;               // [3C][C]
  float* Wp_s = Wqkv_s + C3 * C;         // [C][C]
  float* W1_s = Wp_s + C * C;            // [4C][C]
  float* W2t_s = W1_s + C4 * C;          // [4C][C]  (fc2 weight transposed: hidden-major)
  float* bqkv_s = W2t_s + C4 * C;        // [3C]
  float* bp_s = bqkv_s + C3;             // [C]
  float* b1_s = bp_s + C;                // [4C]
  float* b2_s = b1_s + C4;               // [C]
  float* ln_s = b2_s + C;                // n1w, n1b, n2w, n2b  [4][C]
  float* tab_s = ln_s + 4 * C;           // [81*NH] (+pad)
  float* kv_s = tab_s + ((81 * NH + 3) & ~3);   // [125][2C]
  int* rid_s = reinterpret_cast<int*>(kv_s + SB_WIN * SB_TOK * 2 * C);  // [128]

  for (int i = threadIdx.x; i < C3 * C; i += SB_THREADS) Wqkv_s[i] = p.Wqkv[i];
  for (int i = threadIdx.x; i < C * C; i += SB_THREADS) Wp_s[i] = p.Wp[i];
  for (int i = threadIdx.x; i < C4 * C; i += SB_THREADS) {
    W1_s[i] = p.W1[i];
    W2t_s[i] = p.W2[(i % C) * C4 + i / C];   // W2 is [C][4C] -> [4C][C]
  }
  for (int i = threadIdx.x; i < C3; i += SB_THREADS) bqkv_s[i] = p.bqkv[i];
  for (int i = threadIdx.x; i < C4; i += SB_THREADS) b1_s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += SB_THREADS) {
    bp_s[i] = p.bp[i];
    b2_s[i] = p.b2[i];
    ln_s[i] = p.n1w[i];
    ln_s[C + i] = p.n1b[i];
    ln_s[2 * C + i] = p.n2w[i];
    ln_s[3 * C + i] = p.n2b[i];
  }
  for (int i = threadIdx.x; i < 81 * NH; i += SB_THREADS) tab_s[i] = p.table[i];
  __syncthreads();

  const int t = threadIdx.x;
  const int wl = t / SB_TOK, ti = t - wl * SB_TOK;        // window slot, token inside the window
  const bool slot_ok = t < SB_WIN * SB_TOK;
  const int yi = ti / 5, xi = ti - yi * 5;
  const float qscale = rsqrtf((float)HD);
  const long long n_groups = (n_windows + SB_WIN - 1) / SB_WIN;

  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long w = grp * SB_WIN + wl;
    long long tok = -1;       // -1: padded / inactive slot
    bool win_ok = slot_ok && w < n_windows;
    int rid = 0;
    if (win_ok) {
      const int b = (int)(w / (nWy * nWx));
      const int wr = (int)(w - (long long)b * nWy * nWx);
      const int Y = (wr / nWx) * 5 + yi, X = (wr % nWx) * 5 + xi;
      if (Y < p.H && X < p.W) {
        int y = Y, x = X;
        if (p.shift > 0) {
          y = (Y + p.shift) % p.H;
          x = (X + p.shift) % p.W;
        }
        tok = ((long long)b * p.H + y) * p.W + x;
      }
      if (p.shift > 0) {
        const int Hp = nWy * 5, Wp = nWx * 5;
        rid = (Y < Hp - 5 ? 0 : (Y < Hp - p.shift ? 1 : 2)) * 3 + (X < Wp - 5 ? 0 : (X < Wp - p.shift ? 1 : 2));
      }
    }
    // ---- load row, LN1 (padded slots: zero vector AFTER the norm, SwinWNet.py:242,254) ----
    float x[C], xn[C];
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tok >= 0) v = __ldg(reinterpret_cast<const float4*>(p.x + tok * C + c));
      x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
    }
    {
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) mean += x[c];
      mean *= (1.0f / C);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) var = fmaf(x[c] - mean, x[c] - mean, var);
      const float rstd = rsqrtf(var * (1.0f / C) + p.eps);
#pragma unroll
      for (int c = 0; c < C; ++c) xn[c] = tok >= 0 ? fmaf((x[c] - mean) * rstd, ln_s[c], ln_s[C + c]) : 0.f;
    }
    // ---- qkv: q stays in registers (pre-scaled), k|v go to shared memory ----
    float q[C];
#pragma unroll
    for (int o = 0; o < C3; ++o) {
      float acc = bqkv_s[o];
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(Wqkv_s + o * C + c);
        acc = fmaf(wv.x, xn[c], acc); acc = fmaf(wv.y, xn[c + 1], acc);
        acc = fmaf(wv.z, xn[c + 2], acc); acc = fmaf(wv.w, xn[c + 3], acc);
      }
      if (o < C) q[o] = acc * qscale;
      else if (slot_ok) kv_s[t * 2 * C + (o - C)] = acc;
    }
    if (slot_ok) rid_s[t] = rid;
    __syncthreads();
    // ---- attention over the 25 tokens of the own window ----
    float o_att[C];
    const float* kvw = kv_s + (slot_ok ? wl : 0) * SB_TOK * 2 * C;   // (the 3 spare threads compute on slot 0, unused)
    const int ri = yi * 9 + xi + 40;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float s[SB_TOK];
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < SB_TOK; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(kvw + j * 2 * C + h * HD + d);
          acc = fmaf(q[h * HD + d], kk.x, acc); acc = fmaf(q[h * HD + d + 1], kk.y, acc);
          acc = fmaf(q[h * HD + d + 2], kk.z, acc); acc = fmaf(q[h * HD + d + 3], kk.w, acc);
        }
        acc += tab_s[(ri - ((j / 5) * 9 + j % 5)) * NH + h];
        if (p.shift > 0 && slot_ok && rid_s[wl * SB_TOK + j] != rid) acc -= 100.0f;
        s[j] = acc;
        mx = fmaxf(mx, acc);
      }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < SB_TOK; ++j) {
        s[j] = ex2_approx((s[j] - mx) * 1.4426950408889634f);
        den += s[j];
      }
      const float inv = 1.0f / den;
      float acc_o[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) acc_o[d] = 0.f;
#pragma unroll
      for (int j = 0; j < SB_TOK; ++j) {
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(kvw + j * 2 * C + C + h * HD + d);
          acc_o[d] = fmaf(s[j], vv.x, acc_o[d]); acc_o[d + 1] = fmaf(s[j], vv.y, acc_o[d + 1]);
          acc_o[d + 2] = fmaf(s[j], vv.z, acc_o[d + 2]); acc_o[d + 3] = fmaf(s[j], vv.w, acc_o[d + 3]);
        }
      }
#pragma unroll
      for (int d = 0; d < HD; ++d) o_att[h * HD + d] = acc_o[d] * inv;
    }
    __syncthreads();   // kv_s is rewritten by the next iteration
    if (tok < 0) continue;
    // ---- proj + shortcut ----
#pragma unroll
    for (int o = 0; o < C; ++o) {
      float acc = bp_s[o];
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(Wp_s + o * C + c);
        acc = fmaf(wv.x, o_att[c], acc); acc = fmaf(wv.y, o_att[c + 1], acc);
        acc = fmaf(wv.z, o_att[c + 2], acc); acc = fmaf(wv.w, o_att[c + 3], acc);
      }
      x[o] += acc;
    }
    // ---- LN2 + MLP + shortcut ----
    {
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) mean += x[c];
      mean *= (1.0f / C);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) var = fmaf(x[c] - mean, x[c] - mean, var);
      const float rstd = rsqrtf(var * (1.0f / C) + p.eps);
#pragma unroll
      for (int c = 0; c < C; ++c) xn[c] = fmaf((x[c] - mean) * rstd, ln_s[2 * C + c], ln_s[3 * C + c]);
    }
    float y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) y[c] = b2_s[c];
#pragma unroll 2
    for (int u = 0; u < C4; ++u) {
      float acc = b1_s[u];
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(W1_s + u * C + c);
        acc = fmaf(wv.x, xn[c], acc); acc = fmaf(wv.y, xn[c + 1], acc);
        acc = fmaf(wv.z, xn[c + 2], acc); acc = fmaf(wv.w, xn[c + 3], acc);
      }
      const float hval = gelu_erf(acc);
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(W2t_s + u * C + c);
        y[c] = fmaf(wv.x, hval, y[c]); y[c + 1] = fmaf(wv.y, hval, y[c + 1]);
        y[c + 2] = fmaf(wv.z, hval, y[c + 2]); y[c + 3] = fmaf(wv.w, hval, y[c + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c += 4)
      *reinterpret_cast<float4*>(p.out + tok * C + c) = make_float4(x[c] + y[c], x[c + 1] + y[c + 1], x[c + 2] + y[c + 2], x[c + 3] + y[c + 3]);
  }
}

int launch_swin_block_small(SmallBlockParams p, int num_sms, cudaStream_t stream) {
  SWN_CHECK(p.nH == 3 && (p.C == 12 || p.C == 24), "swin_block_small: only C in {12,24} with 3 heads (got C=%d nH=%d)", p.C, p.nH);
  if (p.shift > 0) SWN_CHECK(p.H % 5 == 0 && p.W % 5 == 0 && p.shift < 5, "swin_block_small: shift needs H,W multiples of 5");
  const int nWy = (p.H + 4) / 5, nWx = (p.W + 4) / 5;
  const long long n_windows = (long long)p.B * nWy * nWx;
  const int C = p.C;
  const size_t floats = (size_t)12 * C * C + 3 * C + C + 4 * C + C + 4 * C + ((81 * 3 + 3) & ~3) + (size_t)SB_WIN * SB_TOK * 2 * C;
  const size_t smem = floats * 4 + SB_THREADS * 4 + 16;
  const long long n_groups = (n_windows + SB_WIN - 1) / SB_WIN;
  long long grid = (long long)num_sms * 4;
  if (grid > n_groups) grid = n_groups;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, SB_THREADS, smem, stream>>>(p, nWy, nWx, n_windows);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  return C == 12 ? go(swin_block_small_kernel<12, 3>) : go(swin_block_small_kernel<24, 3>);
}

}  // namespace swn
