// Row-tile GEMM on tcgen05 tensor cores:   out = epilogue( prologue(A)[M,K] * W[N,K]^T )
//
// One CTA owns 128 token rows.  Its A operand (the whole K extent, K <= 768) is produced in shared
// memory by a fused prologue (LayerNorm / PatchMerging 2x2 gather + LayerNorm / plain convert) in the
// UMMA K-major SWIZZLE_128B layout and stays resident; the weight matrix is streamed chunk by chunk
// as pre-swizzled [NT x 64] bf16 tiles through an mbarrier ring filled by the TMA engine
// (cp.async.bulk); accumulators live in TMEM (double buffered per N-chunk) and are drained by four
// epilogue warps with a fused epilogue (bias, bf16 store | bias, gamma scale, residual, fp32 store |
// PatchExpanding pixel-shuffle scatter + LayerNorm + crop).
//
// Replaces, per reference call site (SwinWNet.py): norm1+qkv (:242,:185), proj+residual (:207,:277),
// PatchMerging (:295-313), PatchExpanding (:402-410), decoder linears (:489), MultiheadAttention
// in/out projections incl. norm_q/norm_kv and the gamma residual (:779-783).
#include "common.cuh"
#include "kernels.h"

namespace swn {

constexpr int RG_THREADS = 192;  // warp0 = weight producer, warp1 = MMA issuer, warps 2..5 = epilogue
constexpr int RG_PRO_THREADS = 160;  // warps 1..5 build the A tile

struct RgSmem {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a_ready;
  uint32_t tmem_base;
};

template <int KV>
__device__ __forceinline__ void rg_prologue_row(const RowGemmParams& p, uint8_t* a_smem, int r, long long m,
                                                int lane, int K16) {
  float4 v[KV];
  const bool row_ok = m < p.M;
  int bb = 0, ho = 0, wo = 0;
  if (p.a_mode == A_MERGE_LN && row_ok) {
    int hw = p.gHo * p.gWo;
    bb = (int)(m / hw);
    int rem = (int)(m - (long long)bb * hw);
    ho = rem / p.gWo;
    wo = rem - ho * p.gWo;
  }
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    int k = (i * 32 + lane) * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && k < p.K) {
      if (p.a_mode == A_BF16) {
        const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.A) + m * p.lda + k);
        v[i] = make_float4(bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y));
      } else if (p.a_mode == A_MERGE_LN) {
        int q = k / p.gC;
        int ch = k - q * p.gC;
        int y = 2 * ho + (q & 1), x = 2 * wo + (q >> 1);
        if (y < p.gH && x < p.gW)
          v[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.A) +
                                                  (((long long)bb * p.gH + y) * p.gW + x) * p.gC + ch);
      } else {
        v[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.A) + m * p.lda + k);
      }
    }
  }
  if (p.a_mode == A_F32_LN || p.a_mode == A_MERGE_LN) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < KV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);  // out-of-range lanes hold zeros
    const float mean = warp_sum(s) / (float)p.K;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      int k = (i * 32 + lane) * 4;
      if (k < p.K) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)p.K + p.ln_eps);
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      int k = (i * 32 + lane) * 4;
      if (k < p.K) {
        const float4 g = *reinterpret_cast<const float4*>(p.ln_w + k);
        const float4 be = *reinterpret_cast<const float4*>(p.ln_b + k);
        v[i].x = (v[i].x - mean) * rstd * g.x + be.x;
        v[i].y = (v[i].y - mean) * rstd * g.y + be.y;
        v[i].z = (v[i].z - mean) * rstd * g.z + be.z;
        v[i].w = (v[i].w - mean) * rstd * g.w + be.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    int k = (i * 32 + lane) * 4;
    if (k < K16) {
      uint2 o = make_uint2(0u, 0u);
      if (row_ok && k < p.K) o = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
      *reinterpret_cast<uint2*>(a_smem + (k >> 6) * A_KBLOCK_BYTES + sw128_offset(r, k & 63)) = o;
    }
  }
}

template <int KV>
__global__ void __launch_bounds__(RG_THREADS, 1) rowgemm_kernel(const RowGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int K16 = (p.K + 15) & ~15;
  const int KB = (K16 + 63) >> 6;
  const int ksteps_total = K16 >> 4;
  const int stage_bytes = p.NT * 128;
  uint8_t* a_smem = smem;
  uint8_t* ring = smem + KB * A_KBLOCK_BYTES;
  RgSmem* sh = reinterpret_cast<RgSmem*>(ring + p.stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * TILE_M;
  const uint32_t tmem_cols = p.tmem_cols;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->tmem_full[b], 1);
      mbar_init(&sh->tmem_empty[b], 128);
    }
    mbar_init(&sh->a_ready, RG_PRO_THREADS);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&sh->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const int total_tiles = p.nchunks * KB;
  const int nt32 = (p.NT + 31) & ~31;  // TMEM column stride between the two accumulator buffers

  if (warp == 0) {
    // ===== weight producer: stream pre-swizzled [NT x 64] tiles, consumption order =====
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wp);
      for (int t = 0; t < total_tiles; ++t) {
        const int s = t % p.stages;
        const uint32_t ph = (uint32_t)(t / p.stages) & 1u;
        mbar_wait(&sh->empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], (uint32_t)stage_bytes);
        bulk_g2s(ring + s * stage_bytes, src + (size_t)t * stage_bytes, (uint32_t)stage_bytes, &sh->full[s]);
      }
    }
  } else {
    // ===== prologue: warps 1..5 build the resident A tile (bf16, swizzled) =====
    for (int r = warp - 1; r < TILE_M; r += 5) rg_prologue_row<KV>(p, a_smem, r, m0 + r, lane, K16);
    fence_proxy_async();
    mbar_arrive(&sh->a_ready);

    if (warp == 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
        mbar_wait(&sh->a_ready, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(TILE_M, (uint32_t)p.NT);
        const uint32_t a_addr = smem_u32(a_smem);
        int t = 0;
        for (int n = 0; n < p.nchunks; ++n) {
          const int buf = n & 1;
          mbar_wait(&sh->tmem_empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * nt32);
          for (int kb = 0; kb < KB; ++kb, ++t) {
            const int s = t % p.stages;
            mbar_wait(&sh->full[s], (uint32_t)(t / p.stages) & 1u);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(ring + s * stage_bytes);
            const int steps = min(4, ksteps_total - kb * 4);
            for (int k = 0; k < steps; ++k) {
              umma_bf16(d_tmem, umma_desc_sw128(a_addr + kb * A_KBLOCK_BYTES + k * 32),
                        umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&sh->empty[s]);
          }
          umma_commit(&sh->tmem_full[buf]);
        }
      }
    } else {
      // ===== epilogue: warps 2..5, thread <-> accumulator row (TMEM lane) =====
      const int lg = warp & 3;  // TMEM lane group this warp may access
      const int r = lg * 32 + lane;
      const long long m = m0 + r;
      const bool row_ok = m < p.M;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
      const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
      const int nblk = p.NT >> 4;

      // expand-scatter target of this row (E_EXPAND)
      int eb = 0, eh = 0, ew = 0;
      if (p.e_mode == E_EXPAND && row_ok) {
        int hw = p.xH * p.xW;
        eb = (int)(m / hw);
        int rem = (int)(m - (long long)eb * hw);
        eh = rem / p.xW;
        ew = rem - eh * p.xW;
      }

      for (int n = 0; n < p.nchunks; ++n) {
        const int buf = n & 1;
        mbar_wait(&sh->tmem_full[buf], ((uint32_t)n >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(buf * nt32);
        const float* bias = p.bias ? p.bias + n * p.NT : nullptr;
        float v[16];

        if (p.e_mode == E_EXPAND) {
          const int oy = 2 * eh + (n >> 1), ox = 2 * ew + (n & 1);
          const bool act = row_ok && oy < p.xHs && ox < p.xWs;
          float s = 0.f;
          for (int jb = 0; jb < nblk; ++jb) {
            tmem_ld16(t_row + jb * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) s += (jb * 16 + j < p.n_valid) ? v[j] : 0.f;
          }
          const float mean = s / (float)p.n_valid;
          float q = 0.f;
          for (int jb = 0; jb < nblk; ++jb) {
            tmem_ld16(t_row + jb * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float d = v[j] - mean;
              q += (jb * 16 + j < p.n_valid) ? d * d : 0.f;
            }
          }
          const float rstd = rsqrtf(q / (float)p.n_valid + p.ln_eps);
          float* orow = reinterpret_cast<float*>(p.out) +
                        (((long long)eb * p.xHs + oy) * p.xWs + ox) * (long long)p.ldo;
          for (int jb = 0; jb < nblk; ++jb) {
            tmem_ld16(t_row + jb * 16, v);
            tmem_ld_wait();
            if (act) {
#pragma unroll
              for (int j4 = 0; j4 < 16; j4 += 4) {
                const int c = jb * 16 + j4;
                if (c < p.n_valid) {  // n_valid % 4 == 0 (checked on the host)
                  const float4 g = *reinterpret_cast<const float4*>(p.ln2_w + c);
                  const float4 be = *reinterpret_cast<const float4*>(p.ln2_b + c);
                  float4 o;
                  o.x = (v[j4 + 0] - mean) * rstd * g.x + be.x;
                  o.y = (v[j4 + 1] - mean) * rstd * g.y + be.y;
                  o.z = (v[j4 + 2] - mean) * rstd * g.z + be.z;
                  o.w = (v[j4 + 3] - mean) * rstd * g.w + be.w;
                  *reinterpret_cast<float4*>(orow + c) = o;
                }
              }
            }
          }
        } else {
          const long long col0 = (long long)n * p.n_valid;
          for (int jb = 0; jb < nblk; ++jb) {
            tmem_ld16(t_row + jb * 16, v);
            tmem_ld_wait();
            if (!row_ok) continue;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const int c = jb * 16 + j4;
              if (c >= p.n_valid) continue;  // n_valid % 4 == 0 (checked on the host)
              float4 a = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
              if (bias) {
                const float4 bv = *reinterpret_cast<const float4*>(bias + c);
                a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
              }
              if (p.e_mode == E_BF16) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + col0 + c;
                *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
              } else {
                a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha;
                if (p.res) {
                  const float4 rv = *reinterpret_cast<const float4*>(p.res + m * p.ldres + col0 + c);
                  a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
                }
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldo + col0 + c) = a;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&sh->tmem_empty[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static uint32_t pow2_cols(int c) {
  uint32_t v = 32;
  while ((int)v < c) v <<= 1;
  return v;
}

int launch_rowgemm(RowGemmParams p, cudaStream_t stream) {
  SWN_CHECK(p.M > 0 && p.K > 0 && p.K <= 768 && p.K % 4 == 0, "rowgemm: unsupported K=%d (need K%%4==0, K<=768)", p.K);
  SWN_CHECK(p.NT >= 16 && p.NT <= 256 && p.NT % 16 == 0, "rowgemm: bad NT=%d", p.NT);
  SWN_CHECK(p.n_valid > 0 && p.n_valid <= p.NT && p.n_valid % 4 == 0, "rowgemm: bad n_valid=%d", p.n_valid);
  SWN_CHECK(p.nchunks >= 1, "rowgemm: bad nchunks");
  SWN_CHECK(p.ldo % 4 == 0, "rowgemm: ldo must be a multiple of 4");
  if (p.a_mode == A_BF16) SWN_CHECK(p.lda % 4 == 0, "rowgemm: bf16 lda must be a multiple of 4");
  if (p.a_mode == A_MERGE_LN) SWN_CHECK(p.K == 4 * p.gC && p.gC % 4 == 0, "rowgemm: merge geometry mismatch");
  const int K16 = (p.K + 15) & ~15;
  const int KB = (K16 + 63) >> 6;
  const int stage_bytes = p.NT * 128;
  const int a_bytes = KB * A_KBLOCK_BYTES;
  const int budget = 232448 - 1024 - a_bytes - (int)sizeof(RgSmem) - 64;
  int stages = budget / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages > p.nchunks * KB) stages = p.nchunks * KB;
  SWN_CHECK(stages >= 1, "rowgemm: K=%d NT=%d does not fit in shared memory", p.K, p.NT);
  p.stages = stages;
  p.tmem_cols = (int)pow2_cols(p.nchunks > 1 ? ((p.NT + 31) & ~31) + p.NT : p.NT);
  SWN_CHECK(p.tmem_cols <= 512, "rowgemm: TMEM overflow");
  const size_t smem = 1024 + a_bytes + (size_t)stages * stage_bytes + sizeof(RgSmem) + 64;
  const int KV = (p.K + 127) / 128;
  const long long grid = ((long long)p.M + TILE_M - 1) / TILE_M;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, RG_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  if (KV <= 1) return go(rowgemm_kernel<1>);
  if (KV <= 3) return go(rowgemm_kernel<3>);
  return go(rowgemm_kernel<6>);
}

}  // namespace swn
