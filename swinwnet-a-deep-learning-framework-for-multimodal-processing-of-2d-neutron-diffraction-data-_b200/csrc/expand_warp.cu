// PatchExpanding (+ crop) for the narrow decoder layers (K = 24 / 48 / 96 input channels; SwinWNet.py:397-424): Linear(K -> 2K,
// no bias), pixel shuffle to the 2x grid, LayerNorm(K/2) — with the GEMM on warp-level mma.sync instead of a tcgen05 tile.
// At K = 24 a 128 x 64-k tcgen05 tile is 62 % padding and the per-tile chain (rows -> smem A tile -> MMA -> TMEM -> two
// round trips of the LayerNorm epilogue) ran at 2.3 TB/s.  Here a warp owns 32 consecutive token rows: they are loaded straight
// into A fragments (fp32 -> 16 bit), multiplied with the four channel-group chunks of the weight — the B fragments are read
// directly from the UMMA SWIZZLE_128B image pack_rowgemm produces (row r, 16-byte chunk c at r*128 + ((c ^ r) & 7)*16: the
// 8 rows x 4 lanes of a fragment load hit 32 different banks) — and the LayerNorm of each 12- / 24-channel output row runs on the
// accumulator fragments (two-pass, quad shuffles).  No shared-memory staging of rows, no block barrier after the weight load.
#include "common.cuh"
#include "kernels.h"

namespace swn {

namespace {
constexpr int XW_THREADS = 256;
#ifndef SWN_XW_MINB
#define SWN_XW_MINB 2     // CTAs per SM (3: 80 registers, a few spills)
#endif

__device__ __forceinline__ void xw_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float xw_quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
}  // namespace

template <int KT, int NT8>   // k-steps of 16 input channels, 8-column tiles per channel-group chunk
__global__ void __launch_bounds__(XW_THREADS, SWN_XW_MINB) expand_warp_kernel(const RowGemmParams p) {
  constexpr int KJ = 2 * KT;      // 8-column tiles of the (zero-padded) input row
  extern __shared__ __align__(16) uint8_t xw_smem[];
  constexpr int KB = (KT + 3) / 4;                                      // 64-column k-blocks of the weight image
  uint8_t* w_s = xw_smem;                                               // [4 chunks][KB][NT rows][128 B] swizzled
  float* lnw = reinterpret_cast<float*>(xw_smem + 4 * KB * p.NT * 128); // [NT8 * 8]
  float* lnb = lnw + NT8 * 8;
  for (int i = threadIdx.x; i < 4 * KB * p.NT * 8; i += XW_THREADS)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wp) + i);
  for (int i = threadIdx.x; i < NT8 * 8; i += XW_THREADS) {
    lnw[i] = i < p.n_valid ? p.ln2_w[i] : 0.f;
    lnb[i] = i < p.n_valid ? p.ln2_b[i] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const float* A = reinterpret_cast<const float*>(p.A);
  float* out = reinterpret_cast<float*>(p.out);
  const int hw = p.xH * p.xW;
  const float inv_n = 1.0f / (float)p.n_valid;
  const long long ngroups = ((long long)p.M + 31) / 32;
  for (long long grp = (long long)blockIdx.x * (XW_THREADS / 32) + warp; grp < ngroups; grp += (long long)gridDim.x * (XW_THREADS / 32)) {
    // ---- 32 rows -> A fragments; output coordinates of the four row slots (row 8 s + g of the group) ----
    uint32_t a[2][KT][4];
    long long obase[4];   // output token (2h, 2w) of the row, or -1
    int oy[4], ox[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const long long m = grp * 32 + 8 * s + g;
      const bool ok = m < p.M;
      uint32_t pk[KJ];
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        float2 v = make_float2(0.f, 0.f);
        if (ok && 8 * j + 2 * t < p.K) v = __ldg(reinterpret_cast<const float2*>(A + m * p.lda + 8 * j + 2 * t));
        pk[j] = pack_op(v.x, v.y);
      }
#pragma unroll
      for (int j = 0; j < KJ; ++j) a[s >> 1][j >> 1][(j & 1) * 2 + (s & 1)] = pk[j];
      obase[s] = -1;
      oy[s] = ox[s] = 0;
      if (ok) {
        const int mi = (int)m;          // M < 2^31 (checked by the launcher): 32-bit divisions
        const int eb = mi / hw, rem = mi - eb * hw;
        const int eh = rem / p.xW, ew = rem - eh * p.xW;
        oy[s] = 2 * eh;
        ox[s] = 2 * ew;
        obase[s] = (long long)eb * p.xHs * p.xWs;
      }
    }
    // ---- the four channel groups: chunk n = (i, j) = (n >> 1, n & 1) -> output pixel (2h + i, 2w + j) ----
#pragma unroll 1
    for (int n = 0; n < 4; ++n) {
      const uint8_t* wc = w_s + n * KB * p.NT * 128;
      float acc[2][NT8][4];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        acc[0][nt][0] = acc[0][nt][1] = acc[0][nt][2] = acc[0][nt][3] = 0.f;
        acc[1][nt][0] = acc[1][nt][1] = acc[1][nt][2] = acc[1][nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const uint32_t r = 8 * nt + g;
          const uint8_t* wb = wc + (kt >> 2) * p.NT * 128 + r * 128 + 4 * t;       // k-block kt / 4, 16-byte chunks 2 (kt % 4), + 1
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wb + ((((uint32_t)(2 * (kt & 3))) ^ r) & 7u) * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wb + ((((uint32_t)(2 * (kt & 3) + 1)) ^ r) & 7u) * 16);
          xw_mma(acc[0][nt], a[0][kt], b0, b1);
          xw_mma(acc[1][nt], a[1][kt], b0, b1);
        }
      }
      // LayerNorm over the n_valid channels of every row, on the fragments; store to the shuffled position
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int mt = s >> 1, e0 = (s & 1) * 2;
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) sum += acc[mt][nt][e0] + acc[mt][nt][e0 + 1];     // columns >= n_valid are exact zeros
        const float mean = xw_quad_sum(sum) * inv_n;
        float q = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          const bool cv = 8 * nt + 2 * t < p.n_valid;
          const float d0 = cv ? acc[mt][nt][e0] - mean : 0.f, d1 = cv ? acc[mt][nt][e0 + 1] - mean : 0.f;
          acc[mt][nt][e0] = d0;
          acc[mt][nt][e0 + 1] = d1;
          q = fmaf(d0, d0, q);
          q = fmaf(d1, d1, q);
        }
        float rstd;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rstd) : "f"(xw_quad_sum(q) * inv_n + p.ln_eps));
        const int yy = oy[s] + (n >> 1), xx = ox[s] + (n & 1);
        if (obase[s] >= 0 && yy < p.xHs && xx < p.xWs) {
          float* orow = out + (obase[s] + (long long)yy * p.xWs + xx) * p.ldo;
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) {
            const int c = 8 * nt + 2 * t;
            if (c < p.n_valid) {
              const float2 w2 = *reinterpret_cast<const float2*>(lnw + c), b2 = *reinterpret_cast<const float2*>(lnb + c);
              *reinterpret_cast<float2*>(orow + c) = make_float2(fmaf(acc[mt][nt][e0] * rstd, w2.x, b2.x), fmaf(acc[mt][nt][e0 + 1] * rstd, w2.y, b2.y));
            }
          }
        }
      }
    }
  }
}

// returns 0 and launches if the shape qualifies; -1 (no error set) if the caller should use the tcgen05 kernel
int launch_expand_warp(RowGemmParams p, int num_sms, cudaStream_t stream) {
#ifndef SWN_EXPAND_WARP
#define SWN_EXPAND_WARP 1
#endif
  if (!SWN_EXPAND_WARP || p.e_mode != E_EXPAND || p.a_mode != A_F32 || p.K > 96 || p.K % 4 != 0 || p.nchunks != 4 || p.n_valid % 4 != 0 ||
      (p.NT != 16 && p.NT != 32 && p.NT != 48) || p.lda % 2 != 0 || p.ldo % 2 != 0)
    return -1;
  const int KT = (p.K + 15) / 16;
  if ((p.NT == 16 && KT > 2) || (p.NT == 32 && KT > 3) || (p.NT == 48 && KT != 6)) return -1;
  const size_t smem = (size_t)4 * ((KT + 3) / 4) * p.NT * 128 + 2 * p.NT * 4;
  auto go = [&](auto kern) -> int {
    const long long ngroups = ((long long)p.M + 31) / 32;
    long long grid = (long long)num_sms * SWN_XW_MINB;
    const long long need = (ngroups + XW_THREADS / 32 - 1) / (XW_THREADS / 32);
    if (grid > need) grid = need;
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, XW_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  if (p.NT == 48) return go(expand_warp_kernel<6, 6>);
  if (p.NT == 16) return KT == 1 ? go(expand_warp_kernel<1, 2>) : go(expand_warp_kernel<2, 2>);
  return KT == 1 ? go(expand_warp_kernel<1, 4>) : (KT == 2 ? go(expand_warp_kernel<2, 4>) : go(expand_warp_kernel<3, 4>));
}

}  // namespace swn
