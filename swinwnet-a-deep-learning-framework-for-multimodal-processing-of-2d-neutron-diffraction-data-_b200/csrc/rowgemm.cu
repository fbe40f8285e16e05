// Row-tile GEMM on tcgen05 tensor cores:   out = epilogue( prologue(A)[M,K] * W[N,K]^T )
//
// One CTA (10 warps) owns 128 token rows.  Its A operand (the whole K extent, K <= 768) is produced in
// shared memory by a fused prologue run by ALL warps (LayerNorm / PatchMerging 2x2 gather + LayerNorm /
// plain convert, see build_a_tile) in the UMMA K-major SWIZZLE_128B layout and stays resident; the weight
// matrix is streamed chunk by chunk as pre-swizzled [NT x 64] bf16 tiles through an mbarrier ring filled
// by the TMA engine (cp.async.bulk, warp 0); warp 1 issues the UMMAs; accumulators live in TMEM (double
// buffered per N-chunk) and are drained by eight epilogue warps (two per TMEM lane group, interleaved
// 16-column blocks) with a fused epilogue (bias, bf16 store | bias, gamma scale, residual, fp32 store |
// PatchExpanding pixel-shuffle scatter + LayerNorm + crop).
//
// Replaces, per reference call site (SwinWNet.py): norm1+qkv (:242,:185), proj+residual (:207,:277),
// PatchMerging (:295-313), PatchExpanding (:402-410), decoder linears (:489), MultiheadAttention
// in/out projections incl. norm_q/norm_kv and the gamma residual (:779-783).
#include "common.cuh"
#include "kernels.h"

namespace swn {

#ifndef SWN_RG_UNR2
#define SWN_RG_UNR2 3
#endif
constexpr int RG_WARPS = 10;
constexpr int RG_THREADS = RG_WARPS * 32;

struct RgSmem {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a_ready;
  uint32_t tmem_base;
  // A_MERGE_LN: per-row source of the 2x2 gather, computed once per row (the index divisions were 47 % of the executed
  // instructions of the merge GEMM when every float4 load redid them, profiles/r2_ncu_lines_n_mg48.txt)
  long long row_off[TILE_M];   // element offset of pixel (2ho, 2wo) of the row's image
  int row_flag[TILE_M];        // bit 0: row < M, bit 1: 2ho+1 < gH, bit 2: 2wo+1 < gW
};

// EP: epilogue variant, compile-time so that each one gets its own register allocation (one kernel with all three paths
// needed 150 registers and lost the second co-resident CTA): 0 = direct (E_BF16 / E_F32), 1 = E_EXPAND, 2 = transposed E_F32
template <int LPR, int KV, int EP>
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
  const int total_tiles = p.nchunks * KB;
  const int nt32 = (p.NT + 31) & ~31;  // TMEM column stride between the two accumulator buffers
  const int epi_count = p.e_mode == E_EXPAND ? 128 : 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->tmem_full[b], 1);
      mbar_init(&sh->tmem_empty[b], epi_count / 32);
    }
    fence_barrier_init();
  }
  // Per-CTA setup runs in warp 0 WHILE warps 1.. build the A tile (this kernel is not persistent: with the setup in
  // front of the prologue, 26 % of all stall samples of the qkv GEMM sat at the barrier behind tcgen05.alloc,
  // profiles/r2_ncu_rowgemm_qkv192.txt): first ring fill (same thread as the barrier init: program order), TMEM
  // allocation, bias -> shared memory (the epilogue's per-block __ldg of the bias was another 14 %).
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.Wp);
  const int prefetch = min(p.stages, total_tiles);
  float* bias_s = reinterpret_cast<float*>(sh + 1);
  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < prefetch; ++t) {
        mbar_arrive_expect_tx(&sh->full[t], (uint32_t)stage_bytes);
        bulk_g2s(ring + t * stage_bytes, wsrc + (size_t)t * stage_bytes, (uint32_t)stage_bytes, &sh->full[t]);
      }
    }
    __syncwarp();
    tmem_alloc(&sh->tmem_base, tmem_cols);
    if (p.bias)
      for (int i = lane; i < p.nchunks * p.NT; i += 32) bias_s[i] = p.bias[i];
  }

  // ===== prologue: warps 1.. build the resident A tile (bf16, swizzled) =====
  if (warp > 0) {
    const int pw = warp - 1;
    constexpr int PWARPS = RG_WARPS - 1;
    constexpr int UNR = KV == 1 ? 4 : (KV == 2 ? SWN_RG_UNR2 : (KV == 3 ? 2 : 1));
    const int M = p.M;
    if (p.a_mode == A_BF16) {
      const op_t* A = reinterpret_cast<const op_t*>(p.A);
      const int lda = p.lda;
      build_a_tile<LPR, KV, UNR, false>(a_smem, p.K, K16, nullptr, nullptr, 0.f, pw, PWARPS, lane, [&](int r, int k) {
        const long long m = m0 + r;
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        const uint2 raw = *reinterpret_cast<const uint2*>(A + m * lda + k);
        return make_float4(op_lo(raw.x), op_hi(raw.x), op_lo(raw.y), op_hi(raw.y));
      });
    } else if (p.a_mode == A_MERGE_LN) {
      const float* A = reinterpret_cast<const float*>(p.A);
      const int gH = p.gH, gW = p.gW, gC = p.gC, gWo = p.gWo, hw = p.gHo * p.gWo;
      {
        const int rr = (int)threadIdx.x - 32;       // prologue threads 32..159 <-> rows 0..127
        if (rr < TILE_M) {
          const long long m = m0 + rr;
          int flag = 0;
          long long off = 0;
          if (m < M) {
            const int bb = (int)(m / hw);
            const int rem = (int)(m - (long long)bb * hw);
            const int ho = rem / gWo, wo = rem - ho * gWo;
            off = (((long long)bb * gH + 2 * ho) * gW + 2 * wo) * gC;
            flag = 1 | (2 * ho + 1 < gH ? 2 : 0) | (2 * wo + 1 < gW ? 4 : 0);
          }
          sh->row_off[rr] = off;
          sh->row_flag[rr] = flag;
        }
        asm volatile("bar.sync 1, %0;" ::"n"((RG_WARPS - 1) * 32) : "memory");   // prologue warps only (warp 0 runs the setup)
      }
      const long long dyoff = (long long)gW * gC;
      build_a_tile<LPR, KV, UNR, true>(a_smem, p.K, K16, p.ln_w, p.ln_b, p.ln_eps, pw, PWARPS, lane, [&](int r, int k) {
        const int q = (k >= gC) + (k >= 2 * gC) + (k >= 3 * gC), ch = k - q * gC;     // [x00, x10, x01, x11]: dy = q & 1, dx = q >> 1
        const int flag = sh->row_flag[r];
        if (!(flag & 1) || ((q & 1) && !(flag & 2)) || ((q >> 1) && !(flag & 4))) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(A + sh->row_off[r] + (q & 1) * dyoff + (q >> 1) * gC + ch);
      });
    } else {
      const float* A = reinterpret_cast<const float*>(p.A);
      const int lda = p.lda;
      auto ld = [&](int r, int k) {
        const long long m = m0 + r;
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(A + m * lda + k);
      };
      if (p.a_mode == A_F32_LN)
        build_a_tile<LPR, KV, UNR, true>(a_smem, p.K, K16, p.ln_w, p.ln_b, p.ln_eps, pw, PWARPS, lane, ld);
      else
        build_a_tile<LPR, KV, UNR, false>(a_smem, p.K, K16, nullptr, nullptr, 0.f, pw, PWARPS, lane, ld);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();     // barrier inits, TMEM base address, bias and the A tile are visible to every role
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===== weight producer: remaining tiles, consumption order =====
    if (lane == 0) {
      for (int t = prefetch; t < total_tiles; ++t) {
        const int s = t % p.stages;
        mbar_wait(&sh->empty[s], ((uint32_t)(t / p.stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], (uint32_t)stage_bytes);
        bulk_g2s(ring + s * stage_bytes, wsrc + (size_t)t * stage_bytes, (uint32_t)stage_bytes, &sh->full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit =====
    const uint32_t idesc = umma_idesc_bf16(TILE_M, (uint32_t)p.NT);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(a_smem));
    const uint64_t ring_desc0 = umma_desc_sw128(smem_u32(ring));
    const uint32_t stage_d16 = (uint32_t)(stage_bytes >> 4), kblk_d16 = A_KBLOCK_BYTES >> 4;
    RingPos rp{0, 0u};
    for (int n = 0; n < p.nchunks; ++n) {
      const int buf = n & 1;
      mbar_wait(&sh->tmem_empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * nt32);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&sh->full[rp.s], rp.ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = a_desc0 + (uint64_t)(kb * kblk_d16), bd = ring_desc0 + (uint64_t)(rp.s * stage_d16);
          const int steps = min(4, ksteps_total - kb * 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&sh->empty[rp.s]);
          if (kb == KB - 1) umma_commit(&sh->tmem_full[buf]);
        }
        __syncwarp();
        rp.next(p.stages);
      }
    }
  } else {
    // ===== epilogue: warps 2..9; two warps per TMEM lane group =====
    const int lg = warp & 3;             // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;    // which of the two warps of the lane group
    const int r = lg * 32 + lane;
    const long long m = m0 + r;
    const bool row_ok = m < p.M;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const int nblk = p.NT >> 4;
    float v[16];

    if constexpr (EP == 1) {
      // chunk n = channel group (i = n>>1, j = n&1) -> output pixel (2h+i, 2w+j); warps of `half` own the
      // chunks with n & 1 == half, i.e. always TMEM buffer `half`, and a thread sees its whole row (LayerNorm).
      int eb = 0, eh = 0, ew = 0;
      if (row_ok) {
        const int hw = p.xH * p.xW;
        eb = (int)(m / hw);
        const int rem = (int)(m - (long long)eb * hw);
        eh = rem / p.xW;
        ew = rem - eh * p.xW;
      }
      const float inv_n = 1.0f / (float)p.n_valid;
      for (int n = half; n < p.nchunks; n += 2) {
        mbar_wait(&sh->tmem_full[half], ((uint32_t)n >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(half * nt32);
        const int oy = 2 * eh + (n >> 1), ox = 2 * ew + (n & 1);
        const bool act = row_ok && oy < p.xHs && ox < p.xWs;
        float* orow = reinterpret_cast<float*>(p.out) + (((long long)eb * p.xHs + oy) * p.xWs + ox) * (long long)p.ldo;
        // (keeping the chunk row in registers for ONE TMEM round trip instead of three was measured slower: 127 registers
        // cost the co-resident CTAs these latency-bound shapes live on: expand C=48 0.475 -> 0.757 ms)
        // LayerNorm statistics in ONE pass over the accumulator (shifted moments, shift = first channel: no cancellation
        // for |mean| >> std), second pass normalises and stores: two TMEM round trips per chunk row instead of three
        float s1 = 0.f, s2 = 0.f, v0 = 0.f;
        for (int jb = 0; jb < nblk; ++jb) {
          tmem_ld16(t_row + jb * 16, v);
          tmem_ld_wait();
          if (jb == 0) v0 = v[0];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float d = (jb * 16 + j < p.n_valid) ? v[j] - v0 : 0.f;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
        }
        const float m1 = s1 * inv_n;
        const float mean = v0 + m1;
        const float q = fmaxf(s2 - s1 * m1, 0.f);
        const float rstd = rsqrtf(q * inv_n + p.ln_eps);
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
        tc_fence_before();
        mbar_arrive_warp(&sh->tmem_empty[half]);
      }
    } else if constexpr (EP == 2) {   // launcher chose the transposed epilogue (p.res_stride != 0)
      // out = alpha * (acc + bias) + residual in TRANSPOSED ownership (common.cuh): after tcgen05.ld a lane holds 16 columns
      // of its own row, so direct residual loads / stores touch 32 rows x 16 B per instruction and their latency sat in
      // front of every add (53 % of the stall samples of the C = 384 proj GEMM, profiles/r2_ncu_rowgemm.txt).  The block
      // goes through 2 KB of per-warp scratch; a lane then owns 16-byte chunk (lane & 3) of rows ps*8 + lane/4, the
      // residual of the NEXT block (next chunk included) is requested before the current one is consumed — the first one
      // before the accumulator is even waited for.
      uint8_t* scr = reinterpret_cast<uint8_t*>(bias_s + ((p.nchunks * p.NT + 3) & ~3)) + (warp - 2) * EPI_SCRATCH_BYTES;
      const int cq = (lane & 3) * 4, rq = lg * 32 + (lane >> 2);
      auto load_res = [&](int n, int jb, float4* xr) {
        const int c = jb * 16 + cq;
        const bool on = p.res != nullptr && n < p.nchunks && c < p.n_valid;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const long long mm = m0 + rq + ps * 8;
          xr[ps] = (on && mm < p.M) ? __ldg(reinterpret_cast<const float4*>(p.res + mm * p.ldres + (long long)n * p.n_valid + c))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 xr_cur[4], xr_nxt[4];
      load_res(0, half, xr_cur);
      for (int n = 0; n < p.nchunks; ++n) {
        const int buf = n & 1;
        mbar_wait(&sh->tmem_full[buf], ((uint32_t)n >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(buf * nt32);
        const float* bias = p.bias ? bias_s + n * p.NT : nullptr;
        const long long col0 = (long long)n * p.n_valid;
        for (int jb = half; jb < nblk; jb += 2) {
          if (jb + 2 < nblk) load_res(n, jb + 2, xr_nxt); else load_res(n + 1, half, xr_nxt);
          tmem_ld16(t_row + jb * 16, v);
          tmem_ld_wait();
          const int c0 = jb * 16;
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) bv = *reinterpret_cast<const float4*>(bias + c0 + j4);
            v[j4] = (v[j4] + bv.x) * alpha; v[j4 + 1] = (v[j4 + 1] + bv.y) * alpha;
            v[j4 + 2] = (v[j4 + 2] + bv.z) * alpha; v[j4 + 3] = (v[j4 + 3] + bv.w) * alpha;
          }
          epi_scatter16(scr, v, lane);
          const int c = c0 + cq;
          if (c < p.n_valid) {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const long long mm = m0 + rq + ps * 8;
              if (mm >= p.M) continue;
              const float4 y = epi_gather4(scr, ps, lane);
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + mm * p.ldo + col0 + c) =
                  make_float4(y.x + xr_cur[ps].x, y.y + xr_cur[ps].y, y.z + xr_cur[ps].z, y.w + xr_cur[ps].w);
            }
          }
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) xr_cur[ps] = xr_nxt[ps];
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive_warp(&sh->tmem_empty[buf]);
      }
    } else if constexpr (EP == 3) {   // 16-bit output in TRANSPOSED ownership: 8 rows x 32 contiguous bytes per store instruction
      uint8_t* scr = reinterpret_cast<uint8_t*>(bias_s + ((p.nchunks * p.NT + 3) & ~3)) + (warp - 2) * EPI_SCRATCH_BYTES;
      const int cq = (lane & 3) * 4, rq = lg * 32 + (lane >> 2);
      for (int n = 0; n < p.nchunks; ++n) {
        const int buf = n & 1;
        mbar_wait(&sh->tmem_full[buf], ((uint32_t)n >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(buf * nt32);
        const float* bias = p.bias ? bias_s + n * p.NT : nullptr;
        const long long col0 = (long long)n * p.n_valid;
        for (int jb = half; jb < nblk; jb += 2) {
          tmem_ld16(t_row + jb * 16, v);
          tmem_ld_wait();
          const int c0 = jb * 16;
          if (bias) {
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias + c0 + j4);
              v[j4] += bv.x; v[j4 + 1] += bv.y; v[j4 + 2] += bv.z; v[j4 + 3] += bv.w;
            }
          }
          epi_scatter16(scr, v, lane);
          const int c = c0 + cq;
          if (c < p.n_valid) {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const long long mm = m0 + rq + ps * 8;
              if (mm >= p.M) continue;
              const float4 y = epi_gather4(scr, ps, lane);
              *reinterpret_cast<uint2*>(reinterpret_cast<op_t*>(p.out) + mm * p.ldo + col0 + c) = make_uint2(pack_op(y.x, y.y), pack_op(y.z, y.w));
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive_warp(&sh->tmem_empty[buf]);
      }
    } else {
      const bool vec8 = (p.ldo % 8 == 0) && (p.n_valid % 8 == 0);
      for (int n = 0; n < p.nchunks; ++n) {
        const int buf = n & 1;
        mbar_wait(&sh->tmem_full[buf], ((uint32_t)n >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(buf * nt32);
        const float* bias = p.bias ? bias_s + n * p.NT : nullptr;
        const long long col0 = (long long)n * p.n_valid;
        for (int jb = half; jb < nblk; jb += 2) {
          tmem_ld16(t_row + jb * 16, v);
          tmem_ld_wait();
          if (!row_ok) continue;
          const int c0 = jb * 16;
          if (bias) {
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias + c0 + j4);
              v[j4] += bv.x; v[j4 + 1] += bv.y; v[j4 + 2] += bv.z; v[j4 + 3] += bv.w;
            }
          }
          if (p.e_mode == E_BF16) {
            op_t* o = reinterpret_cast<op_t*>(p.out) + m * p.ldo + col0 + c0;
            if (vec8 && c0 + 16 <= p.n_valid) {
              *reinterpret_cast<uint4*>(o) = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]),
                                                        pack_op(v[6], v[7]));
              *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]),
                                                            pack_op(v[12], v[13]), pack_op(v[14], v[15]));
            } else {
#pragma unroll
              for (int j4 = 0; j4 < 16; j4 += 4)
                if (c0 + j4 < p.n_valid)
                  *reinterpret_cast<uint2*>(o + j4) = make_uint2(pack_op(v[j4], v[j4 + 1]), pack_op(v[j4 + 2], v[j4 + 3]));
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + col0 + c0;
            const float* rs = p.res ? p.res + m * p.ldres + col0 + c0 : nullptr;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              if (c0 + j4 >= p.n_valid) continue;  // n_valid % 4 == 0 (checked on the host)
              float4 a = make_float4(v[j4] * alpha, v[j4 + 1] * alpha, v[j4 + 2] * alpha, v[j4 + 3] * alpha);
              if (rs) {
                const float4 rv = *reinterpret_cast<const float4*>(rs + j4);
                a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
              }
              *reinterpret_cast<float4*>(o + j4) = a;
            }
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&sh->tmem_empty[buf]);
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
  // The transposed fp32 epilogue (coalesced, prefetched residual) pays for wide outputs with a residual (C = 384 proj:
  // 0.262 -> 0.176 ms) and loses on narrow ones (48 columns: 0.54 -> 0.82 ms: its 16 KB of scratch costs co-resident CTAs).
  p.res_stride = (p.e_mode == E_F32 && p.res != nullptr && p.nchunks * p.n_valid >= 128 && p.ldres % 4 == 0) ? 1 : 0;
  // + bias staged in shared memory, + 2 KB transposition scratch per epilogue warp for the transposed fp32 epilogue
#ifndef SWN_RG_T16
#define SWN_RG_T16 1
#endif
  // the same transposition for 16-bit outputs: pays at K = 384 (qkv: 0.293 -> 0.275 ms, 0.098 -> 0.094), loses at K = 192 (0.444 ->
  // 0.506 ms: the 16 KB of scratch take two of the five weight-ring stages of the two co-resident CTAs)
  const bool t16 = SWN_RG_T16 && p.e_mode == E_BF16 && p.K > 256 && p.nchunks * p.n_valid >= 128 && p.ldo % 4 == 0;
  const int fixed = 1024 + a_bytes + (int)sizeof(RgSmem) + 64 + p.nchunks * p.NT * 4 + 16 +
                    ((p.res_stride || t16) ? (RG_WARPS - 2) * EPI_SCRATCH_BYTES : 0);
  // aim for >= 2 co-resident CTAs per SM (smem <= ~113 KB) when that still leaves >= 2 ring stages
  int stages = (113 * 1024 - fixed) / stage_bytes;
  if (stages < 2) stages = (232448 - fixed) / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages > p.nchunks * KB) stages = p.nchunks * KB;
  SWN_CHECK(stages >= 1, "rowgemm: K=%d NT=%d does not fit in shared memory", p.K, p.NT);
  p.stages = stages;
  p.tmem_cols = (int)pow2_cols(p.nchunks > 1 ? ((p.NT + 31) & ~31) + p.NT : p.NT);
  SWN_CHECK(p.tmem_cols <= 512, "rowgemm: TMEM overflow");
  const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
  const long long grid = ((long long)p.M + TILE_M - 1) / TILE_M;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, RG_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  const int ep = p.e_mode == E_EXPAND ? 1 : (p.res_stride ? 2 : (t16 ? 3 : 0));
#define SWN_RG_DISPATCH(LPR_, KV_) \
  return ep == 1 ? go(rowgemm_kernel<LPR_, KV_, 1>) : (ep == 2 ? go(rowgemm_kernel<LPR_, KV_, 2>) : (ep == 3 ? go(rowgemm_kernel<LPR_, KV_, 3>) : go(rowgemm_kernel<LPR_, KV_, 0>)))
  if (p.K <= 16) SWN_RG_DISPATCH(4, 1);
  if (p.K <= 32) SWN_RG_DISPATCH(8, 1);
  if (p.K <= 64) SWN_RG_DISPATCH(16, 1);
  if (p.K <= 128) SWN_RG_DISPATCH(32, 1);
  if (p.K <= 256) SWN_RG_DISPATCH(32, 2);
  if (p.K <= 384) SWN_RG_DISPATCH(32, 3);
  SWN_RG_DISPATCH(32, 6);
#undef SWN_RG_DISPATCH
}

}  // namespace swn
