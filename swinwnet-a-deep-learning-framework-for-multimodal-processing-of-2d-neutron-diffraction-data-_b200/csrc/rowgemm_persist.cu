// Persistent, TMA-staged row-tile GEMM for the narrow (HBM-bound) layers:  out = epilogue(prologue(A) * W^T)
// with K*sizeof(A) <= 400 bytes per row (fp32 K <= 96, bf16 K <= 192) and a weight matrix that fits in
// shared memory (<= 64 KB packed).  Same math and operand formats as rowgemm.cu (A_F32_LN / A_F32 / A_BF16
// prologues, E_BF16 / E_F32 epilogues) — this variant exists because those shapes are bound by memory latency
// and instruction issue, not by the tensor pipe:
//   * one CTA per SM loops over 128-row tiles; the packed weights are loaded ONCE per CTA and stay resident;
//   * warp 2 streams the fp32/bf16 input rows (and the fp32 residual rows) of the NEXT tile into padded
//     shared-memory staging with cp.async.bulk — no thread ever waits on a global load;
//   * warps 4-7: one thread per row, LayerNorm / convert from staging -> bf16 SWIZZLE_128B A tile;
//   * warp 1: tcgen05.mma per N-chunk into a ring of TMEM accumulators (runs ahead of the epilogue);
//   * warps 8-15: TMEM -> registers -> bias / gamma / residual (from staging) -> vectorised global stores.
#include "common.cuh"
#include "kernels.h"

namespace swn {

constexpr int RP_WARPS = 16;
constexpr int RP_THREADS = RP_WARPS * 32;
constexpr int RP_CVT_WARPS = 4;
constexpr int RP_EPI_THREADS = 256;
constexpr int RP_MAX_ACC = 4;

struct RpSmem {
  uint64_t w_full;
  uint64_t in_full[2], in_empty[2];
  uint64_t a_full, a_empty;
  uint64_t acc_full[RP_MAX_ACC], acc_empty[RP_MAX_ACC];
  uint32_t tmem_base;
};

template <int EM>
__global__ void __launch_bounds__(RP_THREADS, 1) rowgemm_persist_kernel(const RowGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int K = p.K, K16 = (K + 15) & ~15;
  const int KB = (K16 + 63) >> 6, ksteps_total = K16 >> 4;
  const int NT = p.NT, nchunks = p.nchunks, n_valid = p.n_valid;
  const int wtile_bytes = NT * 128;
  const int w_bytes = nchunks * KB * wtile_bytes;
  const int rs = p.stg_stride, rrs = p.res_stride;
  const int nt32 = (NT + 31) & ~31, nacc = p.stages;  // `stages` = number of TMEM accumulator buffers here
  const bool is_bf16 = p.a_mode == A_BF16;
  const int esz = is_bf16 ? 2 : 4;

  uint8_t* a_smem = smem;                                   // KB k-blocks
  uint8_t* w_smem = a_smem + KB * A_KBLOCK_BYTES;           // resident packed weights
  uint8_t* stg = w_smem + w_bytes;                          // 2 x [128 x rs] input rows
  uint8_t* rstg = stg + 2 * TILE_M * rs;                    // 2 x [128 x rrs] residual rows (optional)
  float* bias_s = reinterpret_cast<float*>(rstg + 2 * TILE_M * rrs);   // [nchunks * NT]
  float* lnw = bias_s + nchunks * NT;                       // [K16]
  float* lnb = lnw + K16;                                   // [K16]
  RpSmem* sh = reinterpret_cast<RpSmem*>(lnb + K16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.M + TILE_M - 1) / TILE_M;
  const bool has_ln = p.a_mode == A_F32_LN;
  // dense staging (rows contiguous in global memory, copied 32 at a time): the row stride is a multiple of 128 bytes for the
  // widths served here, so thread-per-row readers start at a row-dependent 8-column group to spread the shared-memory banks
  const bool dense = rs == K * esz;

  if (threadIdx.x == 0) {
    mbar_init(&sh->w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->in_full[b], 1);
      mbar_init(&sh->in_empty[b], RP_EPI_THREADS / 32);
    }
    mbar_init(&sh->a_full, RP_CVT_WARPS);
    mbar_init(&sh->a_empty, 1);
    for (int b = 0; b < RP_MAX_ACC; ++b) {
      mbar_init(&sh->acc_full[b], 1);
      mbar_init(&sh->acc_empty[b], RP_EPI_THREADS / 32);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < nchunks * NT; i += RP_THREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  for (int i = threadIdx.x; i < K16; i += RP_THREADS) {
    lnw[i] = (has_ln && i < K) ? p.ln_w[i] : 0.f;
    lnb[i] = (has_ln && i < K) ? p.ln_b[i] : 0.f;
  }
  if (warp == 0) tmem_alloc(&sh->tmem_base, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===== weights: one bulk copy, resident for the CTA's lifetime =====
    if (lane == 0) {
      mbar_arrive_expect_tx(&sh->w_full, (uint32_t)w_bytes);
      for (int off = 0; off < w_bytes; off += 32768) {
        const int n = min(32768, w_bytes - off);
        bulk_g2s(w_smem + off, reinterpret_cast<const uint8_t*>(p.Wp) + off, (uint32_t)n, &sh->w_full);
      }
    }
  } else if (warp == 2) {
    // ===== input producer: per-row bulk copies of the next tile (A rows + residual rows) =====
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m0 = (long long)tile * TILE_M;
      const int rows = (int)min((long long)TILE_M, (long long)p.M - m0);
      const int n_total = nchunks * n_valid;
      if (lane == 0) {
        mbar_wait(&sh->in_empty[s], (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->in_full[s], (uint32_t)(rows * (K * esz + (rrs ? n_total * 4 : 0))));
      }
      __syncwarp();
      uint8_t* dst = stg + s * TILE_M * rs;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.A);
      if (dense) {
        // contiguous rows: one copy per 32 rows (a bulk copy costs ~31 ns of the SM's copy engine whatever its size up to
        // ~1 KB — tools/micro/tma_rate.cu — so 384-byte rows copied one by one cap the kernel at 1.9 TB/s of input)
        if (lane * 32 < rows)
          bulk_g2s(dst + lane * 32 * rs, src + (m0 + lane * 32) * (long long)rs, (uint32_t)(min(32, rows - lane * 32) * rs), &sh->in_full[s]);
      } else {
        for (int r = lane; r < rows; r += 32)
          bulk_g2s(dst + r * rs, src + (m0 + r) * (long long)p.lda * esz, (uint32_t)(K * esz), &sh->in_full[s]);
      }
      if (rrs) {
        uint8_t* rdst = rstg + s * TILE_M * rrs;
        for (int r = lane; r < rows; r += 32)
          bulk_g2s(rdst + r * rrs, p.res + (m0 + r) * (long long)p.ldres, (uint32_t)(n_total * 4), &sh->in_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    mbar_wait(&sh->w_full, 0);
    const uint32_t idesc = umma_idesc_bf16(TILE_M, (uint32_t)NT);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(a_smem));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(w_smem));
    const uint32_t wtile_d16 = (uint32_t)(wtile_bytes >> 4), kblk_d16 = A_KBLOCK_BYTES >> 4;
    int it = 0, buf = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      mbar_wait(&sh->a_full, (uint32_t)it & 1u);
      for (int n = 0; n < nchunks; ++n) {
        mbar_wait(&sh->acc_empty[buf], acc_ph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(buf * nt32);
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ad = a_desc0 + (uint64_t)(kb * kblk_d16);
            const uint64_t bd = w_desc0 + (uint64_t)((n * KB + kb) * wtile_d16);
            const int steps = min(4, ksteps_total - kb * 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < steps) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&sh->acc_full[buf]);
          if (n == nchunks - 1) umma_commit(&sh->a_empty);
        }
        __syncwarp();
        if (++buf == nacc) {
          buf = 0;
          acc_ph ^= 1u;
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + RP_CVT_WARPS) {
    // ===== convert / LayerNorm: one thread per row, staging -> bf16 swizzled A tile =====
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m0 = (long long)tile * TILE_M;
      const int rows = (int)min((long long)TILE_M, (long long)p.M - m0);
      mbar_wait(&sh->in_full[s], ((uint32_t)it >> 1) & 1u);
      mbar_wait(&sh->a_empty, ((uint32_t)it & 1u) ^ 1u);
      const int row = (warp - 4) * 32 + lane;
      const uint8_t* src = stg + s * TILE_M * rs + row * rs;
      const bool row_ok = row < rows;
      float mean = 0.f, rstd = 1.f;
      if (has_ln && row_ok) {
        const float x0 = *reinterpret_cast<const float*>(src);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        const int n4 = K / 4;
        int cr = dense ? row % n4 : 0;
        for (int c4 = 0; c4 < n4; ++c4) {
          const float4 v = *reinterpret_cast<const float4*>(src + cr * 16);
          if (++cr == n4) cr = 0;
          const float d0 = v.x - x0, d1 = v.y - x0, d2 = v.z - x0, d3 = v.w - x0;
          s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
          s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]); s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
        }
        const float inv_k = 1.0f / (float)K;
        const float m1 = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * inv_k;
        const float m2 = ((s2[0] + s2[1]) + (s2[2] + s2[3])) * inv_k;
        mean = x0 + m1;
        rstd = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + p.ln_eps);
      }
      const int ng = K16 >> 3;
      int gr = dense ? row % ng : 0;
      for (int q = 0; q < ng; ++q) {
        const int k = gr * 8;
        if (++gr == ng) gr = 0;
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
        if (row_ok && k < K) {
          if (is_bf16) {
            const uint2 lo = *reinterpret_cast<const uint2*>(src + k * 2);
            pk[0] = lo.x; pk[1] = lo.y;
            if (k + 4 < K) {
              const uint2 hi = *reinterpret_cast<const uint2*>(src + k * 2 + 8);
              pk[2] = hi.x; pk[3] = hi.y;
            }
          } else {
            float y[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int kk = k + hh * 4;
              if (kk < K) {
                const float4 v = *reinterpret_cast<const float4*>(src + kk * 4);
                if (has_ln) {
                  const float4 gw = *reinterpret_cast<const float4*>(lnw + kk);
                  const float4 gb = *reinterpret_cast<const float4*>(lnb + kk);
                  y[hh * 4 + 0] = fmaf((v.x - mean) * rstd, gw.x, gb.x);
                  y[hh * 4 + 1] = fmaf((v.y - mean) * rstd, gw.y, gb.y);
                  y[hh * 4 + 2] = fmaf((v.z - mean) * rstd, gw.z, gb.z);
                  y[hh * 4 + 3] = fmaf((v.w - mean) * rstd, gw.w, gb.w);
                } else {
                  y[hh * 4 + 0] = v.x; y[hh * 4 + 1] = v.y; y[hh * 4 + 2] = v.z; y[hh * 4 + 3] = v.w;
                }
              } else {
                y[hh * 4 + 0] = y[hh * 4 + 1] = y[hh * 4 + 2] = y[hh * 4 + 3] = 0.f;
              }
            }
            pk[0] = pack_op(y[0], y[1]); pk[1] = pack_op(y[2], y[3]);
            pk[2] = pack_op(y[4], y[5]); pk[3] = pack_op(y[6], y[7]);
          }
        }
        *reinterpret_cast<uint4*>(a_smem + (k >> 6) * A_KBLOCK_BYTES + sw128_offset(row, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      mbar_arrive_warp(&sh->a_full);
    }
  } else if (warp >= 8) {
    // ===== epilogue warps 8..15: two warps per TMEM lane group, interleaved 16-column blocks =====
    const int lg = warp & 3;
    const int half = (warp - 8) >> 2;
    const int r = lg * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const int nblk = NT >> 4, ldo = p.ldo;
    const bool vec8 = (ldo % 8 == 0) && (n_valid % 8 == 0);
    const bool has_res = p.res != nullptr;
    float v[16];
    int it = 0, buf = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m = (long long)tile * TILE_M + r;
      const bool row_ok = m < p.M;
      if (rrs) mbar_wait(&sh->in_full[s], ((uint32_t)it >> 1) & 1u);   // residual rows have landed
      const float* res_row = rrs ? reinterpret_cast<const float*>(rstg + s * TILE_M * rrs + r * rrs)
                                 : (has_res ? p.res + m * p.ldres : nullptr);
      // residual read from GLOBAL memory (rows too wide to stage): requested one 16-column block ahead, the first block of
      // a tile before its accumulator is waited for — these loads sat in front of every add (37 % of the stall samples of the
      // C = 192 proj GEMM, profiles/r2_ncu_lines_n_p192.txt)
      const bool gres = EM == E_F32 && !rrs && has_res && row_ok;
      float4 rcur[4], rnxt[4];
      auto ldres = [&](int n, int jb, float4* dst) {
        const int c0 = jb * 16, nvb = n < nchunks ? min(16, n_valid - c0) : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = (gres && q * 4 < nvb) ? __ldg(reinterpret_cast<const float4*>(res_row + n * n_valid + c0 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      ldres(0, half, rcur);
      for (int n = 0; n < nchunks; ++n) {
        mbar_wait(&sh->acc_full[buf], acc_ph);
        tc_fence_after();
        const uint32_t t_row = lane_addr + (uint32_t)(buf * nt32);
        const float* bias = bias_s + n * NT;
        const int col0 = n * n_valid;
        for (int jb = half; jb < nblk; jb += 2) {
          if (jb + 2 < nblk) ldres(n, jb + 2, rnxt); else ldres(n + 1, half, rnxt);
          tmem_ld16(t_row + jb * 16, v);
          tmem_ld_wait();
          const int c0 = jb * 16;
          const int nvb = min(16, n_valid - c0);
          float4 rr[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rr[q] = rcur[q];
            rcur[q] = rnxt[q];
          }
          if (!row_ok || nvb <= 0) continue;
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias + c0 + j4);
            v[j4] += bv.x; v[j4 + 1] += bv.y; v[j4 + 2] += bv.z; v[j4 + 3] += bv.w;
          }
          if (EM == E_BF16) {
            op_t* o = reinterpret_cast<op_t*>(p.out) + m * ldo + col0 + c0;
            if (vec8 && nvb == 16) {
              *reinterpret_cast<uint4*>(o) = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]),
                                                        pack_op(v[6], v[7]));
              *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]),
                                                            pack_op(v[12], v[13]), pack_op(v[14], v[15]));
            } else {
#pragma unroll
              for (int j4 = 0; j4 < 16; j4 += 4)
                if (j4 < nvb) *reinterpret_cast<uint2*>(o + j4) = make_uint2(pack_op(v[j4], v[j4 + 1]), pack_op(v[j4 + 2], v[j4 + 3]));
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + m * ldo + col0 + c0;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              if (j4 < nvb) {
                float4 a = make_float4(v[j4] * alpha, v[j4 + 1] * alpha, v[j4 + 2] * alpha, v[j4 + 3] * alpha);
                if (gres) {
                  const float4 rv = rr[j4 >> 2];
                  a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
                } else if (res_row) {
                  const float4 rv = *reinterpret_cast<const float4*>(res_row + col0 + c0 + j4);
                  a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
                }
                *reinterpret_cast<float4*>(o + j4) = a;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&sh->acc_empty[buf]);
        if (++buf == nacc) {
          buf = 0;
          acc_ph ^= 1u;
        }
      }
      mbar_arrive_warp(&sh->in_empty[s]);   // staging (residual rows) of this tile no longer needed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// returns 0 and launches if the shape qualifies; returns -1 (no error set) if the caller should use rowgemm.cu
int launch_rowgemm_persist(RowGemmParams p, int num_sms, cudaStream_t stream) {
  if (p.a_mode == A_MERGE_LN || p.e_mode == E_EXPAND) return -1;
#ifdef SWN_RP_NO_WIDE_RES
  if (p.e_mode == E_F32 && p.res && p.nchunks * p.n_valid >= 128) return -1;   // A/B: rowgemm.cu's transposed residual epilogue
#endif
  const int esz = p.a_mode == A_BF16 ? 2 : 4;
  // rows shorter than 192 B make the per-row bulk copies (~60 cycles each) the bottleneck: measured slower than rowgemm.cu
  if (p.K * esz > 400 || p.K * esz < 192 || p.K % 4 != 0 || (p.lda * esz) % 16 != 0 || (p.K * esz) % 16 != 0) return -1;
  const int K16 = (p.K + 15) & ~15, KB = (K16 + 63) >> 6;
  const int w_bytes = p.nchunks * KB * p.NT * 128;
  if (w_bytes > 80 * 1024) return -1;
  auto padded = [](int bytes) { int ch = (bytes + 15) / 16; return (ch + ((ch & 1) ? 0 : 1)) * 16; };
  const int n_total = p.nchunks * p.n_valid;
  p.stg_stride = (p.lda == p.K && (p.K * esz) % 128 == 0) ? p.K * esz : padded(p.K * esz);   // dense rows: 32 rows per bulk copy
  p.res_stride = (p.e_mode == E_F32 && p.res && n_total * 4 <= 400 && p.ldres % 4 == 0 && (n_total * 4) % 16 == 0) ? padded(n_total * 4) : 0;
  const int nt32 = (p.NT + 31) & ~31;
  int nacc = 512 / nt32;
  if (nacc > RP_MAX_ACC) nacc = RP_MAX_ACC;
  if (nacc < 1) return -1;
  int tc = 32;
  while (tc < nacc * nt32) tc <<= 1;
  if (tc > 512) {
    --nacc;
    tc = 512;
  }
  p.stages = nacc;
  p.tmem_cols = tc;
  const size_t smem = 1024 + (size_t)KB * A_KBLOCK_BYTES + w_bytes + 2 * TILE_M * (p.stg_stride + p.res_stride) +
                      (size_t)(p.nchunks * p.NT + 2 * K16) * 4 + sizeof(RpSmem) + 64;
  if (smem > 232448) return -1;
  const int ntiles = (p.M + TILE_M - 1) / TILE_M;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, RP_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  return p.e_mode == E_BF16 ? go(rowgemm_persist_kernel<E_BF16>) : go(rowgemm_persist_kernel<E_F32>);
}

}  // namespace swn
