// Window attention core (W-MSA / SW-MSA) for 5x5 windows:  softmax(q k^T * hd^-0.5 + rel_pos_bias [+ shift
// mask]) v, reading the token-ordered qkv tensor produced by the fused norm1+qkv GEMM and writing the
// token-ordered attention output consumed by the proj GEMM.  window_partition / window_reverse, the
// zero padding to a multiple of 5, torch.roll and compute_mask of the reference (SwinWNet.py:86-149,
// 183-206, 246-272) are pure index math here: no window tensor is ever materialised.
//
// Padding semantics: the reference pads AFTER norm1, so a padded token is an exact zero vector whose
// q/k/v equal the qkv bias; such tokens take part as keys and their outputs are dropped.
// Shift semantics (shift>0, beyond what the reference can execute): standard Swin —
// windows are taken on the grid rolled by -shift, tokens attend only inside their region id.
#include "common.cuh"
#include "kernels.h"

namespace swn {

constexpr int WA_THREADS = 256;
constexpr int WS = 5, WN = 25;

template <int HD>
__global__ void __launch_bounds__(WA_THREADS) window_attn_kernel(const WinAttnParams p, int nWy, int nWx, int wpb,
                                                                  long long n_windows) {
  extern __shared__ uint8_t smem_raw[];
  const int C = p.C, nH = p.nH;
  // smem: k,v as bf16 [wpb][25][2C] ; rel-pos table fp32 [81*nH] ; token index [wpb][25]
  __nv_bfloat16* kv_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  float* tab_s = reinterpret_cast<float*>(smem_raw + (size_t)wpb * WN * 2 * C * sizeof(__nv_bfloat16));
  long long* tok_s = reinterpret_cast<long long*>(tab_s + ((81 * nH + 1) & ~1));
  int* rid_s = reinterpret_cast<int*>(tok_s + wpb * WN);

  const long long w0 = (long long)blockIdx.x * wpb;
  for (int i = threadIdx.x; i < 81 * nH; i += WA_THREADS) tab_s[i] = p.rpb_table[i];
  // token map: window-local slot -> global token row (or -1 for a padded slot), region id for the mask
  for (int i = threadIdx.x; i < wpb * WN; i += WA_THREADS) {
    const long long w = w0 + i / WN;
    long long tok = -1;
    int rid = 0;
    if (w < n_windows) {
      const int t = i % WN;
      const int b = (int)(w / (nWy * nWx));
      const int wr = (int)(w - (long long)b * nWy * nWx);
      const int Y = (wr / nWx) * WS + t / WS, X = (wr % nWx) * WS + t % WS;
      if (Y < p.H && X < p.W) {
        int y = Y, x = X;
        if (p.shift > 0) {
          y = (Y + p.shift) % p.H;
          x = (X + p.shift) % p.W;
        }
        tok = ((long long)b * p.H + y) * p.W + x;
      }
      if (p.shift > 0) {
        const int Hp = nWy * WS, Wp = nWx * WS;
        const int ry = Y < Hp - WS ? 0 : (Y < Hp - p.shift ? 1 : 2);
        const int rx = X < Wp - WS ? 0 : (X < Wp - p.shift ? 1 : 2);
        rid = ry * 3 + rx;
      }
    }
    tok_s[i] = tok;
    rid_s[i] = rid;
  }
  __syncthreads();
  // stage k|v (2C bf16 per token) ; padded slots take the bias
  const int vec_per_tok = (2 * C) / 4;  // 8-byte vectors
  for (int i = threadIdx.x; i < wpb * WN * vec_per_tok; i += WA_THREADS) {
    const int slot = i / vec_per_tok, c4 = (i - slot * vec_per_tok) * 4;
    const long long tok = tok_s[slot];
    uint2 val;
    if (tok >= 0) {
      val = *reinterpret_cast<const uint2*>(p.qkv + tok * (3 * C) + C + c4);
    } else {
      const float4 bv = *reinterpret_cast<const float4*>(p.qkv_bias + C + c4);
      val = make_uint2(pack_bf16(bv.x, bv.y), pack_bf16(bv.z, bv.w));
    }
    *reinterpret_cast<uint2*>(kv_s + (size_t)slot * 2 * C + c4) = val;
  }
  __syncthreads();

  const float scale = rsqrtf((float)HD);
  const int pairs = wpb * nH * WN;
  for (int pr = threadIdx.x; pr < pairs; pr += WA_THREADS) {
    const int wl = pr / (nH * WN);
    const int rem = pr - wl * nH * WN;
    const int h = rem / WN, i = rem - h * WN;
    const long long tok = tok_s[wl * WN + i];
    if (tok < 0) continue;  // padded query: output dropped
    float q[HD];
    {
      const __nv_bfloat16* qp = p.qkv + tok * (3 * C) + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const uint2 raw = *reinterpret_cast<const uint2*>(qp + d);
        q[d] = bf16_lo(raw.x) * scale;
        q[d + 1] = bf16_hi(raw.x) * scale;
        q[d + 2] = bf16_lo(raw.y) * scale;
        q[d + 3] = bf16_hi(raw.y) * scale;
      }
    }
    const int yi = i / WS, xi = i % WS;
    const int my_rid = rid_s[wl * WN + i];
    const __nv_bfloat16* kbase = kv_s + (size_t)wl * WN * 2 * C + h * HD;
    float s[WN];
    float mx = -1e30f;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const __nv_bfloat16* kp = kbase + (size_t)j * 2 * C;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const uint2 raw = *reinterpret_cast<const uint2*>(kp + d);
        acc = fmaf(q[d], bf16_lo(raw.x), acc);
        acc = fmaf(q[d + 1], bf16_hi(raw.x), acc);
        acc = fmaf(q[d + 2], bf16_lo(raw.y), acc);
        acc = fmaf(q[d + 3], bf16_hi(raw.y), acc);
      }
      const int yj = j / WS, xj = j % WS;
      acc += tab_s[((yi - yj + WS - 1) * (2 * WS - 1) + (xi - xj + WS - 1)) * nH + h];
      if (p.shift > 0 && rid_s[wl * WN + j] != my_rid) acc -= 100.0f;
      s[j] = acc;
      mx = fmaxf(mx, acc);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      s[j] = __expf(s[j] - mx);
      den += s[j];
    }
    const float inv = 1.0f / den;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const __nv_bfloat16* vp = kbase + (size_t)j * 2 * C + C;
      const float pj = s[j] * inv;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const uint2 raw = *reinterpret_cast<const uint2*>(vp + d);
        o[d] = fmaf(pj, bf16_lo(raw.x), o[d]);
        o[d + 1] = fmaf(pj, bf16_hi(raw.x), o[d + 1]);
        o[d + 2] = fmaf(pj, bf16_lo(raw.y), o[d + 2]);
        o[d + 3] = fmaf(pj, bf16_hi(raw.y), o[d + 3]);
      }
    }
    __nv_bfloat16* op = p.out + tok * C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<uint2*>(op + d) = make_uint2(pack_bf16(o[d], o[d + 1]), pack_bf16(o[d + 2], o[d + 3]));
  }
}

int launch_window_attn(WinAttnParams p, cudaStream_t stream) {
  SWN_CHECK(p.C % p.nH == 0, "window_attn: C %% nH != 0");
  const int hd = p.C / p.nH;
  SWN_CHECK(hd == 4 || hd == 8 || hd == 16 || hd == 32, "window_attn: unsupported head_dim %d", hd);
  if (p.shift > 0)
    SWN_CHECK(p.H % WS == 0 && p.W % WS == 0 && p.shift < WS,
              "window_attn: shift>0 needs H,W multiples of the window size (reference semantics undefined otherwise)");
  const int nWy = (p.H + WS - 1) / WS, nWx = (p.W + WS - 1) / WS;
  const long long n_windows = (long long)p.B * nWy * nWx;
  int wpb = WA_THREADS / (p.nH * WN);
  if (wpb < 1) wpb = 1;
  if (wpb > 8) wpb = 8;
  const size_t smem = (size_t)wpb * WN * 2 * p.C * 2 + (size_t)((81 * p.nH + 1) & ~1) * 4 + (size_t)wpb * WN * 12 + 16;
  const long long grid = (n_windows + wpb - 1) / wpb;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, WA_THREADS, smem, stream>>>(p, nWy, nWx, wpb, n_windows);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  switch (hd) {
    case 4: return go(window_attn_kernel<4>);
    case 8: return go(window_attn_kernel<8>);
    case 16: return go(window_attn_kernel<16>);
    default: return go(window_attn_kernel<32>);
  }
}

}  // namespace swn
