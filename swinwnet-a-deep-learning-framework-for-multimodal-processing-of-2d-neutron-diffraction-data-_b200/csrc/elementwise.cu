// Bandwidth-bound kernels of the SwinWNet forward: ScaleAwarePatchEmbed (SwinWNet.py:53-82), the
// SegmentationHead / UpscalingHead conv tails (SwinWNet.py:507-531, 682-688), skip-concat column copy
// (SwinWNet.py:483) and the ST inference pipeline glue (ST_Inference_Pipline.py:32-67, 90-97, 127-134).
// All are vectorised, coalesced CUDA-core kernels judged on achieved HBM GB/s.
#include "common.cuh"
#include "kernels.h"

namespace swn {

// ---------------------------------------------------------------------------------------------
// Patch embed: 2x2 conv with stride 2*s and dilation s (s=1: LR, s=2: HR), bias, LayerNorm(48).
// One thread per token; results are staged through shared memory so the [tokens x 48] fp32 output is
// written fully coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int PE_E = 48, PE_THREADS = 128;

__global__ void __launch_bounds__(PE_THREADS) patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ ln_w,
                                                                  const float* __restrict__ ln_b, float* __restrict__ out,
                                                                  int B, int Cin, int H, int W, int Ho, int Wo, int s) {
  __shared__ float w_s[4 * 4 * PE_E];  // [Cin*4 taps][48], Cin <= 4
  __shared__ float p_s[3 * PE_E];      // bias, ln_w, ln_b
  __shared__ float o_s[PE_THREADS * (PE_E + 1)];
  const int taps = Cin * 4;
  for (int i = threadIdx.x; i < taps * PE_E; i += PE_THREADS) {
    const int tap = i / PE_E, e = i % PE_E;  // w is [E][Cin][2][2] -> tap = c*4 + a*2 + b
    w_s[i] = w[e * taps + tap];
  }
  for (int i = threadIdx.x; i < PE_E; i += PE_THREADS) {
    p_s[i] = bias[i];
    p_s[PE_E + i] = ln_w[i];
    p_s[2 * PE_E + i] = ln_b[i];
  }
  __syncthreads();
  const long long total = (long long)B * Ho * Wo;
  const long long t0 = (long long)blockIdx.x * PE_THREADS;
  const long long t = t0 + threadIdx.x;
  if (t < total) {
    const int b = (int)(t / ((long long)Ho * Wo));
    const int rem = (int)(t - (long long)b * Ho * Wo);
    const int i = rem / Wo, j = rem - i * Wo;
    float acc[PE_E];
#pragma unroll
    for (int e = 0; e < PE_E; ++e) acc[e] = p_s[e];
    for (int c = 0; c < Cin; ++c) {
#pragma unroll
      for (int ab = 0; ab < 4; ++ab) {
        const int y = i * 2 * s + (ab >> 1) * s, xx = j * 2 * s + (ab & 1) * s;
        const float v = (y < H && xx < W) ? __ldg(x + (((long long)b * Cin + c) * H + y) * W + xx) : 0.f;
        const float* wr = w_s + (c * 4 + ab) * PE_E;
#pragma unroll
        for (int e = 0; e < PE_E; ++e) acc[e] = fmaf(v, wr[e], acc[e]);
      }
    }
    float mean = 0.f;
#pragma unroll
    for (int e = 0; e < PE_E; ++e) mean += acc[e];
    mean *= (1.0f / PE_E);
    float var = 0.f;
#pragma unroll
    for (int e = 0; e < PE_E; ++e) {
      const float d = acc[e] - mean;
      var += d * d;
    }
    const float rstd = rsqrtf(var * (1.0f / PE_E) + 1e-5f);
#pragma unroll
    for (int e = 0; e < PE_E; ++e)
      o_s[threadIdx.x * (PE_E + 1) + e] = (acc[e] - mean) * rstd * p_s[PE_E + e] + p_s[2 * PE_E + e];
  }
  __syncthreads();
  const long long n_here = min((long long)PE_THREADS, total - t0);
  float* dst = out + t0 * PE_E;
  for (int i = threadIdx.x; i < n_here * PE_E; i += PE_THREADS) dst[i] = o_s[(i / PE_E) * (PE_E + 1) + (i % PE_E)];
}

int launch_patch_embed(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b,
                       float* out, int B, int Cin, int H, int W, int Ho, int Wo, int scale, cudaStream_t st) {
  SWN_CHECK(Cin >= 1 && Cin <= 4, "patch_embed: in_chans %d unsupported (1..4)", Cin);
  const long long total = (long long)B * Ho * Wo;
  patch_embed_kernel<<<(unsigned)((total + PE_THREADS - 1) / PE_THREADS), PE_THREADS, 0, st>>>(x, w, b, ln_w, ln_b, out, B,
                                                                                            Cin, H, W, Ho, Wo, scale);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Head tail: tokens (NHWC, CI channels) -> conv3x3(CI->CM, pad 1) -> GELU -> conv1x1(CM->Cout) ->
// NCHW output cropped to [Hout, Wout].  fp32 CUDA-core kernel (the heads feed the returned logits / image directly,
// so they stay in full precision).  A thread owns PX consecutive pixels of a row and all CM mid channels: every
// broadcast float4 of the [tap][ci][cm] weights in shared memory feeds 4*PX FMAs, which moves the kernel from the
// shared-memory pipe (one LDS.128 per 4 FMAs in the one-pixel-per-thread form, 3-4x off) to the FMA pipe.
// ---------------------------------------------------------------------------------------------
template <int CI, int CM, int PX>
__global__ void __launch_bounds__(128) conv_head_kernel(const float* __restrict__ tok, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ out, int B,
                                                        int Hh, int Wh, int Cout, int Hout, int Wout) {
  extern __shared__ __align__(16) float cw_s[];  // [9][CI][CM] + b1[CM] + w2[Cout][CM] + b2[Cout]
  float* b1_s = cw_s + 9 * CI * CM;
  float* w2_s = b1_s + CM;
  float* b2_s = w2_s + 2 * CM;
  for (int i = threadIdx.x; i < 9 * CI * CM; i += blockDim.x) {
    const int tap = i / (CI * CM), ci = (i / CM) % CI, cm = i % CM;  // w1 is [CM][CI][3][3]
    cw_s[i] = w1[(cm * CI + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < CM; i += blockDim.x) b1_s[i] = b1[i];
  for (int i = threadIdx.x; i < Cout * CM; i += blockDim.x) w2_s[i] = w2[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) b2_s[i] = b2[i];
  __syncthreads();
  const int groups = (Wout + PX - 1) / PX;                 // pixel groups per output row
  const long long total = (long long)B * Hout * groups;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int b = (int)(t / ((long long)Hout * groups));
  const int rem = (int)(t - (long long)b * Hout * groups);
  const int y = rem / groups, x0 = (rem - y * groups) * PX;
  float acc[PX][CM];
#pragma unroll
  for (int px = 0; px < PX; ++px)
#pragma unroll
    for (int m = 0; m < CM; ++m) acc[px][m] = b1_s[m];
#pragma unroll 1
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    if (yy < 0 || yy >= Hh) continue;
    const float* rowp = tok + ((long long)b * Hh + yy) * Wh * CI;
#pragma unroll 1
    for (int c4 = 0; c4 < CI / 4; ++c4) {
      float4 v[PX + 2];
#pragma unroll
      for (int i = 0; i < PX + 2; ++i) {
        const int xx = x0 - 1 + i;
        v[i] = (xx >= 0 && xx < Wh) ? __ldg(reinterpret_cast<const float4*>(rowp + (long long)xx * CI + c4 * 4))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wr = reinterpret_cast<const float4*>(cw_s + ((dy * 3 + dx) * CI + c4 * 4 + u) * CM);
#pragma unroll
          for (int m4 = 0; m4 < CM / 4; ++m4) {
            const float4 ww = wr[m4];
#pragma unroll
            for (int px = 0; px < PX; ++px) {
              const float4 vv = v[px + dx];
              const float a = u == 0 ? vv.x : (u == 1 ? vv.y : (u == 2 ? vv.z : vv.w));
              acc[px][m4 * 4 + 0] = fmaf(a, ww.x, acc[px][m4 * 4 + 0]);
              acc[px][m4 * 4 + 1] = fmaf(a, ww.y, acc[px][m4 * 4 + 1]);
              acc[px][m4 * 4 + 2] = fmaf(a, ww.z, acc[px][m4 * 4 + 2]);
              acc[px][m4 * 4 + 3] = fmaf(a, ww.w, acc[px][m4 * 4 + 3]);
            }
          }
        }
    }
  }
#pragma unroll
  for (int px = 0; px < PX; ++px)
#pragma unroll
    for (int m = 0; m < CM; ++m) {
      const float a = acc[px][m];
      acc[px][m] = 0.5f * a * (1.0f + erff(a * 0.70710678118654752f));
    }
  for (int co = 0; co < Cout; ++co) {
    float r[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      r[px] = b2_s[co];
#pragma unroll
      for (int m = 0; m < CM; ++m) r[px] = fmaf(acc[px][m], w2_s[co * CM + m], r[px]);
    }
    float* dst = out + (((long long)b * Cout + co) * Hout + y) * Wout + x0;
    if (PX == 4 && (Wout & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int px = 0; px < PX; ++px)
        if (x0 + px < Wout) dst[px] = r[px];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core form of the same head tail (implicit GEMM on mma.sync m16n8k8 TF32, fp32 accumulate): M = pixels,
// K = 9 taps x CI (taps padded to whole k8 steps), N = CM.  Persistent CTAs stage the tap-major weights once (rounded
// to TF32), then loop over 8 x 32 pixel tiles: the (8+2) x (32+2) x CI input halo is staged in shared memory (rounded to
// TF32 once, pixel stride padded so that the A-fragment loads are bank-conflict free), each warp owns one tile row
// (two m16 tiles); epilogue = +bias, exact-erf GELU, the 1x1 conv as per-lane partial dot products reduced over the
// four lanes that share a pixel, cropped NCHW store.  TF32 keeps 10 mantissa bits of the operands (error ~1e-4 of the
// output scale, 200x inside the 2e-2 gate); the fp32 CUDA-core kernel above stays as the exact variant.
// ---------------------------------------------------------------------------------------------
// fp32 -> TF32, round-to-nearest (ties away from zero on the magnitude): adding half a TF32 ulp to the bit pattern lets
// the mma's truncation of the low 13 mantissa bits do the rounding.  One integer add; `cvt.rna.tf32.f32` runs on a
// slow conversion pipe and was 12 % of the recon head's stall samples (profiles/r2_ncu_heads.txt).
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CI, int CM>
__global__ void __launch_bounds__(256) conv_head_mma_kernel(const float* __restrict__ tok, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, float* __restrict__ out, int B,
                                                            int Hh, int Wh, int Cout, int Hout, int Wout) {
  constexpr int TH = 8, TW = 32;
  constexpr int CIP = (CI % 32 == 16) ? CI + 4 : CI;      // pixel stride (floats): conflict-free fragment loads
  constexpr int KST = (CI + 7) / 8;                       // k8 steps per tap
  constexpr int NT = (CM + 7) / 8;                        // n8 tiles
  constexpr int CMP = 24;                                 // weight row stride (floats): conflict-free B loads
  static_assert(CI % 4 == 0 && CM <= CMP && CM % 2 == 0, "unsupported head shape");
  extern __shared__ __align__(16) float cm_s[];
  float* w_s = cm_s;                                      // [9][KST][8][CMP]
  float* in_s = w_s + 9 * KST * 8 * CMP;                  // [(TH+2)*(TW+2)][CIP] (+8 floats slack for the k padding)
  float* b1_s = in_s + (TH + 2) * (TW + 2) * CIP + 8;     // [NT*8]
  float* w2_s = b1_s + NT * 8;                            // [2][NT*8]
  float* b2_s = w2_s + 2 * NT * 8;                        // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < 9 * KST * 8 * CMP; i += 256) {
    const int n = i % CMP, k = (i / CMP) % 8, ks = (i / (CMP * 8)) % KST, tap = i / (CMP * 8 * KST);
    const int ci = ks * 8 + k;
    w_s[i] = (n < CM && ci < CI) ? to_tf32(w1[(n * CI + ci) * 9 + tap]) : 0.f;     // w1 is [CM][CI][3][3]
  }
  for (int i = tid; i < NT * 8; i += 256) {
    b1_s[i] = i < CM ? b1[i] : 0.f;
    w2_s[i] = i < CM ? w2[i] : 0.f;
    w2_s[NT * 8 + i] = (i < CM && Cout > 1) ? w2[CM + i] : 0.f;
  }
  if (tid < 2) b2_s[tid] = tid < Cout ? b2[tid] : 0.f;
  if (tid < 8) in_s[(TH + 2) * (TW + 2) * CIP + tid] = 0.f;
  const int tiles_x = (Wout + TW - 1) / TW, tiles_y = (Hout + TH - 1) / TH;
  const long long n_tiles = (long long)B * tiles_y * tiles_x;
  // halo tile of the input: NLD float4 per thread.  Narrow inputs (CI <= 16: the recon head) are register-prefetched one
  // tile ahead — the loads of tile i+1 are in flight while tile i runs its mma phase (they were 31 % of the stall samples
  // when they sat in front of the staging stores); wide inputs load in one batch at the top of the tile.
  constexpr int NLD = ((TH + 2) * (TW + 2) * (CI / 4) + 255) / 256;
  constexpr bool PREFETCH = NLD <= 4;
  float4 pre[PREFETCH ? NLD : 1];
  auto load_tile = [&](long long tile, float4* dst) {
    const int b = (int)(tile / (tiles_y * tiles_x));
    const int tr = (int)(tile - (long long)b * tiles_y * tiles_x);
    const int y0 = (tr / tiles_x) * TH, x0 = (tr % tiles_x) * TW;
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int i = tid + q * 256;
      const int c4 = i % (CI / 4), pix = i / (CI / 4);
      const int yy = y0 - 1 + pix / (TW + 2), xx = x0 - 1 + pix % (TW + 2);
      dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < (TH + 2) * (TW + 2) * (CI / 4) && yy >= 0 && yy < Hh && xx >= 0 && xx < Wh)
        dst[q] = __ldg(reinterpret_cast<const float4*>(tok + (((long long)b * Hh + yy) * Wh + xx) * CI + c4 * 4));
    }
  };
  auto store_tile = [&](const float4* src) {
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int i = tid + q * 256;
      if (i < (TH + 2) * (TW + 2) * (CI / 4)) {
        const int c4 = i % (CI / 4), pix = i / (CI / 4);
        *reinterpret_cast<float4*>(in_s + pix * CIP + c4 * 4) = make_float4(to_tf32(src[q].x), to_tf32(src[q].y), to_tf32(src[q].z), to_tf32(src[q].w));
      }
    }
  };
  if constexpr (PREFETCH) if ((long long)blockIdx.x < n_tiles) load_tile(blockIdx.x, pre);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_y * tiles_x));
    const int tr = (int)(tile - (long long)b * tiles_y * tiles_x);
    const int y0 = (tr / tiles_x) * TH, x0 = (tr % tiles_x) * TW;
    __syncthreads();   // previous tile fully consumed (and the weights staged, first time round)
    if constexpr (PREFETCH) {
      store_tile(pre);
      if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, pre);
    } else {   // (a 16-float4 register batch was measured slower here: 124 registers)
      for (int i = tid; i < (TH + 2) * (TW + 2) * (CI / 4); i += 256) {
        const int c4 = i % (CI / 4), pix = i / (CI / 4);
        const int yy = y0 - 1 + pix / (TW + 2), xx = x0 - 1 + pix % (TW + 2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < Hh && xx >= 0 && xx < Wh) v = __ldg(reinterpret_cast<const float4*>(tok + (((long long)b * Hh + yy) * Wh + xx) * CI + c4 * 4));
        *reinterpret_cast<float4*>(in_s + pix * CIP + c4 * 4) = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
    }
    __syncthreads();
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* arow = in_s + ((warp + tap / 3) * (TW + 2) + tap % 3 + g) * CIP + t4;
      const float* wt = w_s + tap * KST * 8 * CMP + t4 * CMP + g;
#pragma unroll
      for (int ks = 0; ks < KST; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const float* ap = arow + m * 16 * CIP + ks * 8;
          a[m][0] = __float_as_uint(ap[0]);
          a[m][1] = __float_as_uint(ap[8 * CIP]);
          a[m][2] = __float_as_uint(ap[4]);
          a[m][3] = __float_as_uint(ap[8 * CIP + 4]);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint32_t b0 = __float_as_uint(wt[ks * 8 * CMP + n * 8]), b1v = __float_as_uint(wt[(ks * 8 + 4) * CMP + n * 8]);
          mma_tf32(acc[0][n], a[0], b0, b1v);
          mma_tf32(acc[1][n], a[1], b0, b1v);
        }
      }
    }
    // epilogue: accumulator (row g | g+8 = pixel, col 2*t4 | 2*t4+1 = channel)
    const int y = y0 + warp;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [pixel half][cout]
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ch = n * 8 + 2 * t4 + (e & 1);
          const float av = acc[m][n][e] + b1_s[ch];
          const float h = gelu_erf(av);     // A&S 7.1.26 erf, |err| <= 1.5e-7: exact at fp32 output precision
          o[e >> 1][0] = fmaf(h, w2_s[ch], o[e >> 1][0]);
          o[e >> 1][1] = fmaf(h, w2_s[NT * 8 + ch], o[e >> 1][1]);
        }
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf)
#pragma unroll
        for (int co = 0; co < 2; ++co) {
          float r = o[hlf][co];
          r += __shfl_xor_sync(0xffffffffu, r, 1);
          r += __shfl_xor_sync(0xffffffffu, r, 2);
          const int x = x0 + m * 16 + hlf * 8 + g;
          if (t4 == 0 && co < Cout && y < Hout && x < Wout) out[(((long long)b * Cout + co) * Hout + y) * Wout + x] = r + b2_s[co];
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fp16-operand variant of the head kernel (default): same implicit GEMM on mma.sync.m16n8k16 with fp32 accumulation.  fp16 has
// the mantissa of TF32 (10 bits), so the rounding error of the operands is the same as in the TF32 kernel above, but one mma
// covers 16 input channels instead of 8 (half the mma and fragment-load instructions: the heads are issue-bound) and the
// staged halo tile is half the size.  Inputs are saturated to the fp16 range by the conversion (pack_half2_sat).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// GELU of the conv heads' hidden layer: erf form (Abramowitz-Stegun, |d| <= 7e-7).  The tanh form of the MLP kernels
// (-DSWN_HEAD_GELU=gelu_fast, |d| <= 3e-5) saves 0.08 ms per step, but these are the OUTPUT layers: on a random-init model, whose
// logits are tiny, it doubled the max-norm error of the HR logits in bench.py's parity check (4.1e-3 -> 8.9e-3).
#ifndef SWN_HEAD_GELU
#define SWN_HEAD_GELU gelu_erf
#endif
template <int CI, int CM>
__global__ void __launch_bounds__(256) conv_head_h_kernel(const float* __restrict__ tok, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, float* __restrict__ out, int B,
                                                          int Hh, int Wh, int Cout, int Hout, int Wout) {
  constexpr int TH = 8, TW = 32;
  constexpr int KST = (CI + 15) / 16;                     // k16 steps per tap
  constexpr int K16 = KST * 16;
  constexpr int PSB = ((K16 * 2 / 16) % 2 == 1) ? K16 * 2 : K16 * 2 + 16;   // pixel stride in bytes: odd number of 16-byte chunks
  constexpr int NT = (CM + 7) / 8;                        // n8 tiles
  constexpr int WSB = 48;                                 // weight row (one n, 16 k) stride in bytes: conflict-free B loads
  static_assert(CI % 4 == 0 && CM % 2 == 0, "unsupported head shape");
  extern __shared__ __align__(16) uint8_t ch_s[];
  uint8_t* w_s = ch_s;                                    // [9][KST][NT*8][WSB]
  uint8_t* in_s = w_s + 9 * KST * NT * 8 * WSB;           // [(TH+2)*(TW+2)][PSB]
  float* b1_s = reinterpret_cast<float*>(in_s + (TH + 2) * (TW + 2) * PSB);   // [NT*8]
  float* w2_s = b1_s + NT * 8;                            // [2][NT*8]
  float* b2_s = w2_s + 2 * NT * 8;                        // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < 9 * KST * NT * 8 * 8; i += 256) {          // one half2 (k pair) per step
    const int kp = i % 8, n = (i / 8) % (NT * 8), ks = (i / (8 * NT * 8)) % KST, tap = i / (8 * NT * 8 * KST);
    const int ci = ks * 16 + kp * 2;
    const float lo = (n < CM && ci < CI) ? w1[(n * CI + ci) * 9 + tap] : 0.f;
    const float hi = (n < CM && ci + 1 < CI) ? w1[(n * CI + ci + 1) * 9 + tap] : 0.f;
    *reinterpret_cast<uint32_t*>(w_s + ((tap * KST + ks) * NT * 8 + n) * WSB + kp * 4) = pack_half2_sat(lo, hi);
  }
  for (int i = tid; i < NT * 8; i += 256) {
    b1_s[i] = i < CM ? b1[i] : 0.f;
    w2_s[i] = i < CM ? w2[i] : 0.f;
    w2_s[NT * 8 + i] = (i < CM && Cout > 1) ? w2[CM + i] : 0.f;
  }
  if (tid < 2) b2_s[tid] = tid < Cout ? b2[tid] : 0.f;
  // zero the k padding of every pixel once (channels CI..K16 are never written by the staging loop)
  if (K16 > CI)
    for (int i = tid; i < (TH + 2) * (TW + 2); i += 256)
      for (int c = CI; c < K16; c += 2) *reinterpret_cast<uint32_t*>(in_s + i * PSB + c * 2) = 0u;
  const int tiles_x = (Wout + TW - 1) / TW, tiles_y = (Hout + TH - 1) / TH;
  const int n_tiles = B * tiles_y * tiles_x;      // < 2^31 (checked by the launcher): 32-bit tile arithmetic
  constexpr int NLD = ((TH + 2) * (TW + 2) * (CI / 4) + 255) / 256;
  constexpr bool PREFETCH = NLD <= 4;
  float4 pre[PREFETCH ? NLD : 1];
  auto load_tile = [&](int tile, float4* dst) {
    const int b = tile / (tiles_y * tiles_x);
    const int tr = tile - b * tiles_y * tiles_x;
    const int y0 = (tr / tiles_x) * TH, x0 = (tr % tiles_x) * TW;
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int i = tid + q * 256;
      const int c4 = i % (CI / 4), pix = i / (CI / 4);
      const int yy = y0 - 1 + pix / (TW + 2), xx = x0 - 1 + pix % (TW + 2);
      dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < (TH + 2) * (TW + 2) * (CI / 4) && yy >= 0 && yy < Hh && xx >= 0 && xx < Wh)
        dst[q] = __ldg(reinterpret_cast<const float4*>(tok + (((long long)b * Hh + yy) * Wh + xx) * CI + c4 * 4));
    }
  };
  if constexpr (PREFETCH) if ((int)blockIdx.x < n_tiles) load_tile((int)blockIdx.x, pre);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / (tiles_y * tiles_x);
    const int tr = tile - b * tiles_y * tiles_x;
    const int y0 = (tr / tiles_x) * TH, x0 = (tr % tiles_x) * TW;
    __syncthreads();   // previous tile fully consumed (and the weights staged, first time round)
    if constexpr (PREFETCH) {
#pragma unroll
      for (int q = 0; q < NLD; ++q) {
        const int i = tid + q * 256;
        if (i < (TH + 2) * (TW + 2) * (CI / 4)) {
          const int c4 = i % (CI / 4), pix = i / (CI / 4);
          *reinterpret_cast<uint2*>(in_s + pix * PSB + c4 * 8) = make_uint2(pack_half2_sat(pre[q].x, pre[q].y), pack_half2_sat(pre[q].z, pre[q].w));
        }
      }
      if (tile + (int)gridDim.x < n_tiles) load_tile(tile + (int)gridDim.x, pre);
    } else {
      for (int i = tid; i < (TH + 2) * (TW + 2) * (CI / 4); i += 256) {
        const int c4 = i % (CI / 4), pix = i / (CI / 4);
        const int yy = y0 - 1 + pix / (TW + 2), xx = x0 - 1 + pix % (TW + 2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < Hh && xx >= 0 && xx < Wh) v = __ldg(reinterpret_cast<const float4*>(tok + (((long long)b * Hh + yy) * Wh + xx) * CI + c4 * 4));
        *reinterpret_cast<uint2*>(in_s + pix * PSB + c4 * 8) = make_uint2(pack_half2_sat(v.x, v.y), pack_half2_sat(v.z, v.w));
      }
    }
    __syncthreads();
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      // A rows = 16 consecutive pixels of image row (warp + tap/3) starting at x offset tap%3 (+16 for the second m tile)
      const uint8_t* arow = in_s + ((warp + tap / 3) * (TW + 2) + tap % 3 + g) * PSB + t4 * 4;
      const uint8_t* wt = w_s + (tap * KST * NT * 8 + g) * WSB + t4 * 4;
#pragma unroll
      for (int ks = 0; ks < KST; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const uint8_t* ap = arow + m * 16 * PSB + ks * 32;
          a[m][0] = *reinterpret_cast<const uint32_t*>(ap);
          a[m][1] = *reinterpret_cast<const uint32_t*>(ap + 8 * PSB);
          a[m][2] = *reinterpret_cast<const uint32_t*>(ap + 16);
          a[m][3] = *reinterpret_cast<const uint32_t*>(ap + 8 * PSB + 16);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint8_t* wp = wt + (ks * NT * 8 + n * 8) * WSB;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp), b1v = *reinterpret_cast<const uint32_t*>(wp + 16);
          mma_f16(acc[0][n], a[0], b0, b1v);
          mma_f16(acc[1][n], a[1], b0, b1v);
        }
      }
    }
    // epilogue: accumulator (row g | g+8 = pixel, col 2*t4 | 2*t4+1 = channel)
    const int y = y0 + warp;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [pixel half][cout]
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ch = n * 8 + 2 * t4 + (e & 1);
          const float h = SWN_HEAD_GELU(acc[m][n][e] + b1_s[ch]);
          o[e >> 1][0] = fmaf(h, w2_s[ch], o[e >> 1][0]);
          o[e >> 1][1] = fmaf(h, w2_s[NT * 8 + ch], o[e >> 1][1]);
        }
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf)
#pragma unroll
        for (int co = 0; co < 2; ++co) {
          float r = o[hlf][co];
          r += __shfl_xor_sync(0xffffffffu, r, 1);
          r += __shfl_xor_sync(0xffffffffu, r, 2);
          const int x = x0 + m * 16 + hlf * 8 + g;
          if (t4 == 0 && co < Cout && y < Hout && x < Wout) out[(((long long)b * Cout + co) * Hout + y) * Wout + x] = r + b2_s[co];
        }
    }
  }
}

#ifndef SWN_HEAD_F16
#define SWN_HEAD_F16 1
#endif

template <int CI, int CM>
static int launch_conv_head_mma(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                                int B, int Hh, int Wh, int Cout, int Hout, int Wout, int ctas_per_sm, cudaStream_t st) {
  const long long n_tiles_h = (long long)B * ((Hout + 7) / 8) * ((Wout + 31) / 32);
  if (SWN_HEAD_F16) {
    constexpr int KSTH = (CI + 15) / 16, K16 = KSTH * 16, NTH = (CM + 7) / 8;
    constexpr int PSB = ((K16 * 2 / 16) % 2 == 1) ? K16 * 2 : K16 * 2 + 16;
    const size_t smem_h = (size_t)9 * KSTH * NTH * 8 * 48 + (size_t)10 * 34 * PSB + (size_t)(NTH * 8 * 3 + 2) * sizeof(float) + 16;
    SWN_CUDA(cudaFuncSetAttribute(conv_head_h_kernel<CI, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
    int dev_h = 0, sms_h = 148;
    if (cudaGetDevice(&dev_h) == cudaSuccess) cudaDeviceGetAttribute(&sms_h, cudaDevAttrMultiProcessorCount, dev_h);
    long long grid_h = (long long)sms_h * ctas_per_sm;
    if (grid_h > n_tiles_h) grid_h = n_tiles_h;
    conv_head_h_kernel<CI, CM><<<(unsigned)grid_h, 256, smem_h, st>>>(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout);
    SWN_CUDA(cudaGetLastError());
    return 0;
  }
  constexpr int CIP = (CI % 32 == 16) ? CI + 4 : CI, KST = (CI + 7) / 8, NT = (CM + 7) / 8;
  const size_t smem = (size_t)(9 * KST * 8 * 24 + 10 * 34 * CIP + 8 + NT * 8 * 3 + 2) * sizeof(float);
  SWN_CUDA(cudaFuncSetAttribute(conv_head_mma_kernel<CI, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (long long)B * ((Hout + 7) / 8) * ((Wout + 31) / 32);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  conv_head_mma_kernel<CI, CM><<<(unsigned)grid, 256, smem, st>>>(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

#ifndef SWN_HEAD_TF32
#define SWN_HEAD_TF32 1
#endif
constexpr bool HEAD_TF32 = SWN_HEAD_TF32 != 0;   // 0: exact fp32 CUDA-core heads

// bilinear upsample (align_corners=False, integer scale) of [B,Hq,Wq] to [B,Hout,Wout] (cropped)
__global__ void bilinear_up_kernel(const float* __restrict__ lo, float* __restrict__ out, int B, int Hq, int Wq, int up,
                                   int Hout, int Wout) {
  const long long total = (long long)B * Hout * Wout;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int b = (int)(t / ((long long)Hout * Wout));
  const int rem = (int)(t - (long long)b * Hout * Wout);
  const int y = rem / Wout, x = rem - y * Wout;
  const float inv = 1.0f / (float)up;
  const float sy = fmaxf((y + 0.5f) * inv - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * inv - 0.5f, 0.f);
  const int y0 = min((int)sy, Hq - 1), x0 = min((int)sx, Wq - 1);
  const int y1 = min(y0 + 1, Hq - 1), x1 = min(x0 + 1, Wq - 1);
  const float fy = sy - (float)y0, fx = sx - (float)x0;
  const float* p = lo + (long long)b * Hq * Wq;
  const float top = p[y0 * Wq + x0] * (1.f - fx) + p[y0 * Wq + x1] * fx;
  const float bot = p[y1 * Wq + x0] * (1.f - fx) + p[y1 * Wq + x1] * fx;
  out[t] = top * (1.f - fy) + bot * fy;
}

int launch_seg_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2, float* lowres,
                    float* out, int B, int Hq, int Wq, int up, int Hout, int Wout, cudaStream_t st) {
  constexpr int CI = 48, CM = 24;
  const size_t smem = (size_t)(9 * CI * CM + CM + 2 * CM + 2) * sizeof(float);
  if (HEAD_TF32) {
    const int rc = launch_conv_head_mma<CI, CM>(tok, w1, b1, w2, b2, lowres, B, Hq, Wq, 1, Hq, Wq, 2, st);
    if (rc) return rc;
  } else {
    constexpr int PX = 4;
    SWN_CUDA(cudaFuncSetAttribute(conv_head_kernel<CI, CM, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_lo = (long long)B * Hq * ((Wq + PX - 1) / PX);
    conv_head_kernel<CI, CM, PX><<<(unsigned)((n_lo + 127) / 128), 128, smem, st>>>(tok, w1, b1, w2, b2, lowres, B, Hq, Wq, 1, Hq, Wq);
    SWN_CUDA(cudaGetLastError());
  }
  const long long n_hi = (long long)B * Hout * Wout;
  bilinear_up_kernel<<<(unsigned)((n_hi + 255) / 256), 256, 0, st>>>(lowres, out, B, Hq, Wq, up, Hout, Wout);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

int launch_recon_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                      int B, int Hh, int Wh, int Cout, int Hout, int Wout, cudaStream_t st) {
  constexpr int CI = 12, CM = 12;
  SWN_CHECK(Cout >= 1 && Cout <= 2, "recon_head: Cout must be 1 or 2");
  SWN_CHECK(Hout <= Hh && Wout <= Wh, "recon_head: crop larger than source");
  const size_t smem = (size_t)(9 * CI * CM + CM + 2 * CM + 2) * sizeof(float);
  if (HEAD_TF32) return launch_conv_head_mma<CI, CM>(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout, 4, st);
  constexpr int PX = 4;
  const long long n = (long long)B * Hout * ((Wout + PX - 1) / PX);
  conv_head_kernel<CI, CM, PX><<<(unsigned)((n + 127) / 128), 128, smem, st>>>(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Column-block copy: dst[r, 0:cols] = src[r, 0:cols] with independent row strides (skip -> right half
// of the decoder concat buffer).
// ---------------------------------------------------------------------------------------------
__global__ void copy_cols_kernel(const float4* __restrict__ src, int lds4, float4* __restrict__ dst, int ldd4,
                                 long long rows, int cols4) {
  const long long total = rows * cols4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4);
    dst[r * ldd4 + c] = __ldg(src + r * lds4 + c);
  }
}
int launch_copy_cols(const float* src, int lds, float* dst, int ldd, long long rows, int cols, cudaStream_t st) {
  SWN_CHECK(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "copy_cols: need multiples of 4");
  const long long total = rows * (cols / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  copy_cols_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), lds / 4,
                                                    reinterpret_cast<float4*>(dst), ldd / 4, rows, cols / 4);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// ST pipeline glue
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__global__ void minmax_init_kernel(float* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = (i & 1) ? -INFINITY : INFINITY;
}
// seg_map = sigmoid(seg); images2 = ensure_2ch(img) (optional); masked = images * seg_map; per-(b,c) min/max
__global__ void __launch_bounds__(256) sigmoid_mask_kernel(const float* __restrict__ img, int Cimg,
                                                           const float* __restrict__ seg, float* __restrict__ images2,
                                                           float* __restrict__ seg_map, float* __restrict__ masked,
                                                           float* __restrict__ minmax, int Cout, int HW) {
  const int b = blockIdx.z, c = blockIdx.y;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float sg = 1.0f / (1.0f + __expf(-seg[(long long)b * HW + i]));
    float v;
    if (c < Cimg) v = img[((long long)b * Cimg + c) * HW + i];
    else v = sqrtf(fabsf(img[((long long)b * Cimg) * HW + i]));
    if (images2) images2[((long long)b * Cout + c) * HW + i] = v;
    if (c == 0 && seg_map) seg_map[(long long)b * HW + i] = sg;
    const float mv = v * sg;
    masked[((long long)b * Cout + c) * HW + i] = mv;
    lo = fminf(lo, mv);
    hi = fmaxf(hi, mv);
  }
  if (minmax) {
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    __shared__ float slo[8], shi[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
      atomic_min_f(minmax + ((long long)b * Cout + c) * 2, lo);
      atomic_max_f(minmax + ((long long)b * Cout + c) * 2 + 1, hi);
    }
  }
}
int launch_sigmoid_mask(const float* img, int Cimg, const float* seg, float* images2, float* seg_map, float* masked,
                        float* minmax, int B, int Cout, int H, int W, cudaStream_t st) {
  SWN_CHECK(Cout == Cimg || (Cimg == 1 && Cout == 2 && images2), "sigmoid_mask: bad channel configuration");
  if (minmax) {
    minmax_init_kernel<<<(B * Cout * 2 + 255) / 256, 256, 0, st>>>(minmax, B * Cout * 2);
    SWN_CUDA(cudaGetLastError());
  }
  const int HW = H * W;
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, Cout, B);
  sigmoid_mask_kernel<<<grid, 256, 0, st>>>(img, Cimg, seg, images2, seg_map, masked, minmax, Cout, HW);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// ensure_2ch: [B,1,HW] -> [B,2,HW], channel 1 = sqrt(|channel 0|)
__global__ void ensure_2ch_kernel(const float* __restrict__ x, float* __restrict__ out, int HW) {
  const int b = blockIdx.y;
  const float* src = x + (long long)b * HW;
  float* d0 = out + (long long)b * 2 * HW;
  float* d1 = d0 + HW;
  if ((HW & 3) == 0) {
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < HW; i += gridDim.x * blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      *reinterpret_cast<float4*>(d0 + i) = v;
      *reinterpret_cast<float4*>(d1 + i) = make_float4(sqrtf(fabsf(v.x)), sqrtf(fabsf(v.y)), sqrtf(fabsf(v.z)), sqrtf(fabsf(v.w)));
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      const float v = src[i];
      d0[i] = v;
      d1[i] = sqrtf(fabsf(v));
    }
  }
}
int launch_ensure_2ch(const float* x, float* out, int B, int HW, cudaStream_t st) {
  int bx = (HW / 4 + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  ensure_2ch_kernel<<<dim3(bx, B), 256, 0, st>>>(x, out, HW);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// normalize_piecewise / denormalize_piecewise; minmax is [B*C][2]
__global__ void normalize_kernel(const float* __restrict__ x, const float* __restrict__ minmax, float* __restrict__ out,
                                 int HW, float thr, float eps, int inverse) {
  const int bc = blockIdx.y;
  const float lo = minmax[bc * 2], hi = minmax[bc * 2 + 1];
  const float range = hi - lo + eps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float v = x[(long long)bc * HW + i];
    float r;
    if (!inverse) {
      const float x01 = (v - lo) / range;
      r = x01 > thr ? log1pf(x01) : x01;
    } else {
      const float x01 = v > thr ? expm1f(v) : v;
      r = x01 * range + lo;
    }
    out[(long long)bc * HW + i] = r;
  }
}
int launch_normalize(const float* x, const float* minmax, float* out, int BC, int H, int W, float thr, float eps,
                     int inverse, cudaStream_t st) {
  const int HW = H * W;
  int bx = (HW + 255) / 256;
  if (bx > 128) bx = 128;
  dim3 grid(bx, BC);
  normalize_kernel<<<grid, 256, 0, st>>>(x, minmax, out, HW, thr, eps, inverse);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// d-space front end of the physics metrics (Qwrapper.tensor_to_d, Diffraction_metrics.py:35-70): I(d)[b, bin] = sum of
// channel 0 of image b over the pixels whose d = L / (2 sin(|theta|/2)) falls into the bin.  The pixel -> bin map
// depends only on (H, W, centers); it is computed once on the host with the reference's own fp32 bucketize (so that
// pixels on bin edges fall exactly where the reference puts them; -1 = d > 7.5, dropped) and the whole batch is
// reduced by ONE launch: a CTA accumulates a slice of one image in a shared-memory histogram and flushes it with
// global atomics.  (The reference loops over the batch in Python with a device->host copy per sample.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dspace_hist_kernel(const float* __restrict__ img, long long img_stride,
                                                          const int* __restrict__ bin_of_pixel, int n_pix, int n_bins,
                                                          float* __restrict__ out) {
  extern __shared__ float hist_s[];
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) hist_s[i] = 0.f;
  __syncthreads();
  const float* src = img + (long long)blockIdx.y * img_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += gridDim.x * blockDim.x) {
    const int bin = bin_of_pixel[i];
    if (bin >= 0) atomicAdd(&hist_s[bin], src[i]);
  }
  __syncthreads();
  float* dst = out + (long long)blockIdx.y * n_bins;
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
    const float v = hist_s[i];
    if (v != 0.f) atomicAdd(&dst[i], v);
  }
}

int launch_dspace_hist(const float* img, long long img_stride, const int* bin_of_pixel, int B, int n_pix, int n_bins,
                       float* out, cudaStream_t st) {
  SWN_CHECK(B > 0 && n_pix > 0 && n_bins > 0 && n_bins <= 12000, "dspace_hist: bad sizes (B=%d pixels=%d bins=%d)", B, n_pix, n_bins);
  SWN_CUDA(cudaMemsetAsync(out, 0, (size_t)B * n_bins * sizeof(float), st));
  int slices = (n_pix + 256 * 64 - 1) / (256 * 64);     // ~64 pixels per thread
  if (slices < 1) slices = 1;
  dspace_hist_kernel<<<dim3((unsigned)slices, (unsigned)B), 256, (size_t)n_bins * sizeof(float), st>>>(img, img_stride, bin_of_pixel, n_pix,
                                                                                                        n_bins, out);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace swn
