// Persistent, fully pipelined fused Swin MLP for channel widths C <= 96 (the token-heavy, HBM/epilogue-bound
// layers):   out = x + fc2( GELU( fc1( LayerNorm(x) ) ) )      (SwinWNet.py:226-234,278)
//
// One CTA per SM loops over 128-row token tiles.  16 warps, specialised:
//   warp 0  weight producer   pre-swizzled fc1/fc2 tiles -> smem ring            (cp.async.bulk + mbarrier)
//   warp 2  input producer    fp32 rows of the NEXT tile -> padded smem staging  (cp.async.bulk per row; the
//                             staged tile is both the LayerNorm input and the residual, read from HBM once)
//   warps 4-7  LayerNorm      staging -> bf16 K-major SWIZZLE_128B A tile
//   warp 1  MMA issuer        tcgen05.mma: Hacc = A W1_j^T (TMEM, double buffered), Y += Hs W2_j^T (TMEM)
//   warps 8-15 epilogue       TMEM -> +b1, GELU -> bf16 Hs tile (smem, double buffered);  final: Y + b2 + residual
//                             (staging) -> fp32 out
// Every global load goes through the TMA engine one tile ahead of its use, so HBM latency is never exposed;
// the hidden activation never leaves the SM.
//
// DIRECT variant (C = 192, rows too wide to stage twice in shared memory): no input producer / staging; the LayerNorm
// warps read the rows of the NEXT tile straight from global memory (lanes along the row, coalesced) while the MMA /
// epilogue warps work on the current one, and the final epilogue re-reads the residual (L2) one column block ahead
// and stores through a per-warp transposition scratch.  Status: parity-green, but with 4 LayerNorm warps and a single
// A buffer the prologue is not hidden yet — 0.79 ms against 0.73 ms for mlp.cu at M = 483 840 — so it is opt-in
// (SWN_MLP_PERSIST_MAX_C=192); next step: second A buffer + 8 LayerNorm warps.
#include "common.cuh"
#include "kernels.h"

namespace swn {

// profiling aid (SWN_NVCC_EXTRA=-DSWN_MLP_PROFILE=1, tools/mlp_phase_profile.py); compiled out of the product build
#ifndef SWN_MLP_PROFILE
#define SWN_MLP_PROFILE 0
#endif
#if SWN_MLP_PROFILE
#define MP_TIMED(slot, stmt)                                                                          \
  do {                                                                                                \
    if (p.phase_cycles) {                                                                             \
      const long long _t0 = clock64();                                                                \
      stmt;                                                                                           \
      if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.phase_cycles + (slot)), (unsigned long long)(clock64() - _t0)); \
    } else {                                                                                          \
      stmt;                                                                                           \
    }                                                                                                 \
  } while (0)
#else
#define MP_TIMED(slot, stmt) do { stmt; } while (0)
#endif

#ifndef SWN_MP_EPI_SPLIT
#define SWN_MP_EPI_SPLIT 2
#endif
constexpr int MP_EPI_SPLIT = SWN_MP_EPI_SPLIT;  // epilogue warps per TMEM lane group (they split the columns)
constexpr int MP_WARPS = 8 + 4 * MP_EPI_SPLIT;
constexpr int MP_THREADS = MP_WARPS * 32;
constexpr int MP_LN_WARPS = 4;
// Share of the FINAL epilogue (Y + b2 + residual -> out) that the LayerNorm warps take over (staged variant): they are
// busy ~4.5 k of the ~18 k cycles of a C = 96 tile while the 8 epilogue warps carry GELU (6.5 k) AND the final epilogue
// (4 k) on the critical path (profiles/r2a_mlp_role_waits.txt).  1 = the LN warps take two thirds of the 16-column blocks.
// MEASURED SLOWER (C = 96: 0.934 -> 1.077 ms, C = 48: 0.620 -> 0.808 ms, with the Y accumulator double-buffered): the deferred
// final epilogue holds the staging buffer of tile i until LN(i+1) is done, which delays the row loads of tile i+2 (the staging
// would need a third buffer).  Off by default.
#ifndef SWN_MP_FINAL_LN
#define SWN_MP_FINAL_LN 0
#endif
constexpr int MP_EPI_THREADS = 128 * MP_EPI_SPLIT;
// Staged variant with TWO A tiles: LayerNorm of tile i+1 runs while the GEMMs of tile i are in flight instead of between the
// last GEMM1 of tile i and the first of tile i+1.  Needs 2 more k-blocks of shared memory (32 KB at C = 96: only with 64-column
// hidden chunks, -DSWN_MLP96_HC=64).
#ifndef SWN_MP_NA2
#define SWN_MP_NA2 0
#endif

struct MpSmem {
  uint64_t full[8], empty[8];
  uint64_t in_full[2], in_empty[2];
  uint64_t hacc_full[2], hacc_empty[2], hs_full[2], hs_empty[2];
  uint64_t a_full[2], a_empty[2], y_full[2], y_empty[2];    // DIRECT: two A tiles (LayerNorm of tile i+1 overlaps the GEMMs of tile i)
  uint32_t tmem_base;
};

template <int LPR, bool DIRECT>
__global__ void __launch_bounds__(MP_THREADS, 1) mlp_persist_kernel(const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int C = p.C, C16 = (C + 15) & ~15;
  const int KB1 = (C16 + 63) >> 6, steps1 = C16 >> 4;
  const int HC = p.HC, TR = p.TR;
  const int nj = (4 * C) / HC;
  const int nkk = (HC + 63) >> 6, steps2 = HC >> 4;
  const int nT = C16 / TR;
  constexpr int NY = DIRECT ? 1 : 2;          // staged variant: two Y accumulators (the final epilogue of tile i overlaps the GEMMs of tile i+1)
  const int ystride = (C16 + 31) & ~31;
  const int hbase = NY * ystride;
  const int stage_bytes = max(HC, TR) * 128;
  const int w1_bytes = HC * 128, w2_bytes = TR * 128;
  const int rs = p.row_stride;  // staging row stride in bytes (odd number of 16-byte chunks)

  constexpr int NA = (DIRECT || SWN_MP_NA2) ? 2 : 1;   // A-tile buffers
  constexpr int LN_FIRST = DIRECT ? 2 : 4;             // DIRECT: no input producer -> warps 2..7 run the LayerNorm prologue
  constexpr int LN_WARPS = DIRECT ? 6 : MP_LN_WARPS;
  uint8_t* a_smem = smem;
  uint8_t* hs_smem = a_smem + NA * KB1 * A_KBLOCK_BYTES;
  uint8_t* ring = hs_smem + 2 * nkk * A_KBLOCK_BYTES;
  uint8_t* stg = ring + p.stages * stage_bytes;                     // 2 x [128 x rs]
  float* b1s = reinterpret_cast<float*>(stg + 2 * TILE_M * rs);     // [4C]
  float* b2s = b1s + 4 * C;                                         // [C16]
  float* lnw = b2s + C16;                                           // [C16]
  float* lnb = lnw + C16;                                           // [C16]
  MpSmem* sh = reinterpret_cast<MpSmem*>(lnb + C16);
  uint8_t* epi_scratch = reinterpret_cast<uint8_t*>(sh) + ((sizeof(MpSmem) + 15) & ~size_t(15));   // DIRECT: 8 warps x 2 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.M + TILE_M - 1) / TILE_M;
  const int nb = C16 >> 4;                                              // 16-column blocks of Y
  const int nb_ln = (DIRECT || !SWN_MP_FINAL_LN) ? 0 : (2 * nb + 2) / 3;   // ... of which the LayerNorm warps finish [0, nb_ln)
  const int fin_warps = MP_EPI_THREADS / 32 + (nb_ln > 0 ? MP_LN_WARPS : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->in_full[b], 1);
      mbar_init(&sh->in_empty[b], fin_warps);
      mbar_init(&sh->hacc_full[b], 1);
      mbar_init(&sh->hacc_empty[b], MP_EPI_THREADS / 32);
      mbar_init(&sh->hs_full[b], MP_EPI_THREADS / 32);
      mbar_init(&sh->hs_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->a_full[b], LN_WARPS);
      mbar_init(&sh->a_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->y_full[b], 1);
      mbar_init(&sh->y_empty[b], fin_warps);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 4 * C; i += MP_THREADS) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C16; i += MP_THREADS) {
    b2s[i] = p.b2[i];
    lnw[i] = i < C ? p.ln_w[i] : 0.f;
    lnb[i] = i < C ? p.ln_b[i] : 0.f;
  }
  if (warp == 0) tmem_alloc(&sh->tmem_base, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  // final epilogue of the staged variant for the 32 rows of lane group `lg`, 16-column blocks cb0, cb0 + step, ... < cb1:
  // out = Y + b2 + residual is formed in place in the staging row (row-per-lane, conflict free), then stored in the transposed
  // ownership (common.cuh): 8 rows x 64 contiguous bytes per warp instruction
  auto final_blocks = [&](int tile, int s, int lg, int cb0, int step, int cb1) {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((NY == 2 ? s : 0) * ystride);   // Y buffer = tile parity = s
    uint8_t* stile = stg + s * TILE_M * rs;
    uint8_t* res = stile + (lg * 32 + lane) * rs;
    float v[16];
    for (int cb = cb0; cb < cb1; cb += step) {
      tmem_ld16(lane_addr + (uint32_t)(cb * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const int c = cb * 16 + j4;
        if (c < C) {
          const float4 xr = *reinterpret_cast<const float4*>(res + c * 4);
          const float4 bb = *reinterpret_cast<const float4*>(b2s + c);
          *reinterpret_cast<float4*>(res + c * 4) = make_float4(v[j4 + 0] + bb.x + xr.x, v[j4 + 1] + bb.y + xr.y, v[j4 + 2] + bb.z + xr.z, v[j4 + 3] + bb.w + xr.w);
        }
      }
      __syncwarp();
      const int c = cb * 16 + (lane & 3) * 4;
      if (c < C) {
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int rl = lg * 32 + ps * 8 + (lane >> 2);
          const long long mm = (long long)tile * TILE_M + rl;
          if (mm < p.M) *reinterpret_cast<float4*>(p.out + mm * C + c) = *reinterpret_cast<const float4*>(stile + rl * rs + c * 4);
        }
      }
    }
  };

  if (warp == 0) {
    // ===== weight producer: the whole fc1/fc2 stream once per tile =====
    if (lane == 0) {
      RingPos rp{0, 0u};
      auto push = [&](const uint8_t*& src, int bytes) {
        mbar_wait(&sh->empty[rp.s], rp.ph ^ 1u);
        mbar_arrive_expect_tx(&sh->full[rp.s], (uint32_t)bytes);
        bulk_g2s(ring + rp.s * stage_bytes, src, (uint32_t)bytes, &sh->full[rp.s]);
        src += bytes;
        rp.next(p.stages);
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wp);
        for (int kb = 0; kb < KB1; ++kb) push(src, w1_bytes);
        for (int j = 0; j < nj; ++j) {
          if (j + 1 < nj)
            for (int kb = 0; kb < KB1; ++kb) push(src, w1_bytes);
          for (int i = 0; i < nkk * nT; ++i) push(src, w2_bytes);
        }
      }
    }
  } else if (warp == 2 && !DIRECT) {
    // ===== input producer: one bulk copy per token row into the padded staging tile =====
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m0 = (long long)tile * TILE_M;
      const int rows = (int)min((long long)TILE_M, (long long)p.M - m0);
      if (lane == 0) {
        mbar_wait(&sh->in_empty[s], (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->in_full[s], (uint32_t)(rows * C * 4));
      }
      __syncwarp();
      uint8_t* dst = stg + s * TILE_M * rs;
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(dst + r * rs, p.x + (m0 + r) * C, (uint32_t)(C * 4), &sh->in_full[s]);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit =====
    const uint32_t idesc1 = umma_idesc_bf16(TILE_M, (uint32_t)HC);
    const uint32_t idesc2 = umma_idesc_bf16(TILE_M, (uint32_t)TR);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(a_smem));
    const uint64_t hs_desc0 = umma_desc_sw128(smem_u32(hs_smem));
    const uint64_t ring_desc0 = umma_desc_sw128(smem_u32(ring));
    const uint32_t stage_d16 = (uint32_t)(stage_bytes >> 4), kblk_d16 = A_KBLOCK_BYTES >> 4;
    RingPos rp{0, 0u};
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int g0 = it * nj;  // global chunk counter of this tile's first chunk
      const int ab = NA == 2 ? (it & 1) : 0;                       // A buffer of this tile
      const int yb = NY == 2 ? (it & 1) : 0;                       // Y accumulator of this tile
      const uint32_t yph = NY == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u);
      const uint32_t aph = NA == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u);
      auto gemm1 = [&](int j) {
        const int g = g0 + j, buf = g & 1;
        MP_TIMED(2, mbar_wait(&sh->hacc_empty[buf], (((uint32_t)g >> 1) & 1u) ^ 1u));
        const uint32_t d = tmem_base + (uint32_t)(hbase + buf * HC);
        for (int kb = 0; kb < KB1; ++kb) {
          MP_TIMED(1, mbar_wait(&sh->full[rp.s], rp.ph));
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = a_desc0 + (uint64_t)((ab * KB1 + kb) * kblk_d16), bd = ring_desc0 + (uint64_t)(rp.s * stage_d16);
            const int steps = min(4, steps1 - kb * 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < steps) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&sh->empty[rp.s]);
            if (kb == KB1 - 1) {
              umma_commit(&sh->hacc_full[buf]);
              if (j == nj - 1) umma_commit(&sh->a_empty[ab]);  // A tile free once the last GEMM1 retires
            }
          }
          __syncwarp();
          rp.next(p.stages);
        }
      };
      auto gemm2 = [&](int j) {
        const int g = g0 + j, buf = g & 1;
        MP_TIMED(3, mbar_wait(&sh->hs_full[buf], ((uint32_t)g >> 1) & 1u));
        if (j == 0) MP_TIMED(5, mbar_wait(&sh->y_empty[yb], yph ^ 1u));  // the tile that used this Y accumulator is drained
        const uint64_t hd0 = hs_desc0 + (uint64_t)(buf * nkk * kblk_d16);
        for (int kk = 0; kk < nkk; ++kk) {
          const int steps = min(4, steps2 - kk * 4);
          for (int tt = 0; tt < nT; ++tt) {
            MP_TIMED(4, mbar_wait(&sh->full[rp.s], rp.ph));
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = hd0 + (uint64_t)(kk * kblk_d16), bd = ring_desc0 + (uint64_t)(rp.s * stage_d16);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < steps) umma_bf16(tmem_base + (uint32_t)(yb * ystride + tt * TR), ad + 2 * k, bd + 2 * k, idesc2, (j | kk | k) != 0 ? 1u : 0u);
              umma_commit(&sh->empty[rp.s]);
              if (kk == nkk - 1 && tt == nT - 1) {
                umma_commit(&sh->hs_empty[buf]);
                if (j == nj - 1) umma_commit(&sh->y_full[yb]);
              }
            }
            __syncwarp();
            rp.next(p.stages);
          }
        }
      };
      MP_TIMED(0, mbar_wait(&sh->a_full[ab], aph));
      gemm1(0);
      for (int j = 0; j < nj; ++j) {
        if (j + 1 < nj) gemm1(j + 1);
        gemm2(j);
      }
    }
  } else if (warp >= LN_FIRST && warp < LN_FIRST + LN_WARPS) {
    // ===== LayerNorm warps: staging (fp32) -> A tile (bf16, swizzled) =====
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m0 = (long long)tile * TILE_M;
      const int rows = (int)min((long long)TILE_M, (long long)p.M - m0);
      if (DIRECT) {
        // rows straight from global memory, lanes along the row (coalesced), 4 rows in flight per warp
        const int ab = it & 1;
        mbar_wait(&sh->a_empty[ab], (((uint32_t)it >> 1) & 1u) ^ 1u);
        const float* xg = p.x;
        // 6 rows in flight per warp (12 float4 per lane): with 4 the prologue was bound by global-load latency (29 k
        // cycles per tile, profiles/r2_mlp_direct_role_waits.txt)
        build_a_tile<32, (LPR == 32 ? 2 : 1), 6, true>(a_smem + ab * KB1 * A_KBLOCK_BYTES, C, C16, p.ln_w, p.ln_b, p.ln_eps, warp - LN_FIRST, LN_WARPS, lane, [&](int r, int k) {
          const long long mr = m0 + r;
          if (mr >= p.M) return make_float4(0.f, 0.f, 0.f, 0.f);
          return __ldg(reinterpret_cast<const float4*>(xg + mr * C + k));
        });
        fence_proxy_async();
        mbar_arrive_warp(&sh->a_full[ab]);
        continue;
      }
      const int ab = NA == 2 ? (it & 1) : 0;
      const uint32_t a_par = NA == 2 ? ((((uint32_t)it >> 1) & 1u) ^ 1u) : (((uint32_t)it & 1u) ^ 1u);
      uint8_t* a_tile = a_smem + ab * KB1 * A_KBLOCK_BYTES;
      if (warp == 4) { MP_TIMED(12, mbar_wait(&sh->in_full[s], ((uint32_t)it >> 1) & 1u)); MP_TIMED(13, mbar_wait(&sh->a_empty[ab], a_par)); }
      else { mbar_wait(&sh->in_full[s], ((uint32_t)it >> 1) & 1u); mbar_wait(&sh->a_empty[ab], a_par); }
      const long long t_ln0 = SWN_MLP_PROFILE ? clock64() : 0;
      // one thread per row (the staging rows are padded to an odd number of 16-byte chunks, so this is bank
      // conflict free): no shuffles, long independent instruction streams.  Shifted one-pass moments.
      const int row = (warp - 4) * 32 + lane;
      const uint8_t* src = stg + s * TILE_M * rs + row * rs;
      const bool row_ok = row < rows;
      float mean = 0.f, rstd = 0.f;
      if (row_ok) {
        const float x0 = *reinterpret_cast<const float*>(src);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 v = *reinterpret_cast<const float4*>(src + c4 * 16);
          const float d0 = v.x - x0, d1 = v.y - x0, d2 = v.z - x0, d3 = v.w - x0;
          s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
          s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]); s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
        }
        const float inv_c = 1.0f / (float)C;
        const float m1 = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * inv_c;
        const float m2 = ((s2[0] + s2[1]) + (s2[2] + s2[3])) * inv_c;
        mean = x0 + m1;
        rstd = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + p.ln_eps);
      }
      for (int k = 0; k < C16; k += 8) {  // one 16-byte bf16 chunk of the A tile per step
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
        if (row_ok && k < C) {
          float y[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int kk = k + hh * 4;
            if (kk < C) {
              const float4 v = *reinterpret_cast<const float4*>(src + kk * 4);
              const float4 gw = *reinterpret_cast<const float4*>(lnw + kk);
              const float4 gb = *reinterpret_cast<const float4*>(lnb + kk);
              y[hh * 4 + 0] = fmaf((v.x - mean) * rstd, gw.x, gb.x);
              y[hh * 4 + 1] = fmaf((v.y - mean) * rstd, gw.y, gb.y);
              y[hh * 4 + 2] = fmaf((v.z - mean) * rstd, gw.z, gb.z);
              y[hh * 4 + 3] = fmaf((v.w - mean) * rstd, gw.w, gb.w);
            } else {
              y[hh * 4 + 0] = y[hh * 4 + 1] = y[hh * 4 + 2] = y[hh * 4 + 3] = 0.f;
            }
          }
          pk[0] = pack_op(y[0], y[1]); pk[1] = pack_op(y[2], y[3]);
          pk[2] = pack_op(y[4], y[5]); pk[3] = pack_op(y[6], y[7]);
        }
        *reinterpret_cast<uint4*>(a_tile + (k >> 6) * A_KBLOCK_BYTES + sw128_offset(row, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      mbar_arrive_warp(&sh->a_full[ab]);
#if SWN_MLP_PROFILE
      if (p.phase_cycles && warp == 4 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.phase_cycles + 14), (unsigned long long)(clock64() - t_ln0));
#endif
      // this warp's share of the PREVIOUS tile's final epilogue (its accumulator completes while this tile's chunks start)
      if (nb_ln > 0 && it > 0) {
        const int jt = it - 1;
        mbar_wait(&sh->y_full[jt & 1], ((uint32_t)jt >> 1) & 1u);
        tc_fence_after();
        final_blocks(tile - (int)gridDim.x, jt & 1, warp & 3, 0, 1, nb_ln);
        tc_fence_before();
        mbar_arrive_warp(&sh->y_empty[jt & 1]);
        mbar_arrive_warp(&sh->in_empty[jt & 1]);
      }
    }
    if (!DIRECT && nb_ln > 0 && it > 0) {   // the last tile of this CTA
      const int jt = it - 1;
      mbar_wait(&sh->y_full[jt & 1], ((uint32_t)jt >> 1) & 1u);
      tc_fence_after();
      final_blocks(blockIdx.x + jt * (int)gridDim.x, jt & 1, warp & 3, 0, 1, nb_ln);
      tc_fence_before();
      mbar_arrive_warp(&sh->y_empty[jt & 1]);
      mbar_arrive_warp(&sh->in_empty[jt & 1]);
    }
  } else if (warp >= 8) {
    // ===== epilogue warps 8..: MP_EPI_SPLIT warps per TMEM lane group split the 16-column blocks =====
    const int lg = warp & 3;
    const int part = (warp - 8) >> 2;
    const int r = lg * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    int cb_beg = 0, cb_end = 0;   // 16-column blocks of a hidden chunk owned by this warp
    if (steps2 % MP_EPI_SPLIT == 0) {
      cb_beg = part * (steps2 / MP_EPI_SPLIT);
      cb_end = cb_beg + steps2 / MP_EPI_SPLIT;
    } else if (steps2 % 2 == 0) {
      if (part < 2) {
        cb_beg = part * (steps2 / 2);
        cb_end = cb_beg + steps2 / 2;
      }
    } else if (part == 0) {
      cb_end = steps2;
    }
    float v[16];
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const long long m = (long long)tile * TILE_M + r;
      for (int j = 0; j < nj; ++j) {
        const int g = it * nj + j, buf = g & 1;
        const uint32_t ph = ((uint32_t)g >> 1) & 1u;
        if (warp == 8) { MP_TIMED(6, mbar_wait(&sh->hacc_full[buf], ph)); MP_TIMED(7, mbar_wait(&sh->hs_empty[buf], ph ^ 1u)); }
        else { mbar_wait(&sh->hacc_full[buf], ph); mbar_wait(&sh->hs_empty[buf], ph ^ 1u); }
        tc_fence_after();
        const long long t_g0 = SWN_MLP_PROFILE ? clock64() : 0;
        uint8_t* hrow = hs_smem + buf * nkk * A_KBLOCK_BYTES;
        const float* bj = b1s + j * HC;
        const uint32_t t_chunk = lane_addr + (uint32_t)(hbase + buf * HC);
        // all TMEM loads of this thread's blocks are issued before the first GELU (one exposed TMEM round trip per chunk)
        float vv[4][16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (cb_beg + i < cb_end) tmem_ld16(t_chunk + (cb_beg + i) * 16, vv[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cb = cb_beg + i;
          if (cb < cb_end) {
            const int k = cb * 16;
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bb = *reinterpret_cast<const float4*>(bj + k + 4 * q);
              pk[2 * q] = gelu_pack2(vv[i][4 * q] + bb.x, vv[i][4 * q + 1] + bb.y);
              pk[2 * q + 1] = gelu_pack2(vv[i][4 * q + 2] + bb.z, vv[i][4 * q + 3] + bb.w);
            }
            uint8_t* kb_base = hrow + (k >> 6) * A_KBLOCK_BYTES;
            *reinterpret_cast<uint4*>(kb_base + sw128_offset(r, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(kb_base + sw128_offset(r, (k & 63) + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive_warp(&sh->hacc_empty[buf]);
        mbar_arrive_warp(&sh->hs_full[buf]);
#if SWN_MLP_PROFILE
        if (p.phase_cycles && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.phase_cycles + 8), (unsigned long long)(clock64() - t_g0));
#endif
      }
      // final: Y + b2 + residual(staging) -> out
      const int yb = NY == 2 ? (it & 1) : 0;
      const uint32_t yph = NY == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u);
      if (warp == 8) MP_TIMED(9, mbar_wait(&sh->y_full[yb], yph)); else mbar_wait(&sh->y_full[yb], yph);
      tc_fence_after();
      const long long t_f0 = SWN_MLP_PROFILE ? clock64() : 0;
      if (DIRECT) {
        // residual re-read from global (L2) one column block ahead, transposed ownership through the per-warp scratch
        uint8_t* scr = epi_scratch + (warp - 8) * EPI_SCRATCH_BYTES;
        auto load_res = [&](int cb, float4* xr) {
          const int c = cb * 16 + (lane & 3) * 4;
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const long long mm = (long long)tile * TILE_M + lg * 32 + ps * 8 + (lane >> 2);
            xr[ps] = (cb < (C16 >> 4) && c < C && mm < p.M) ? __ldg(reinterpret_cast<const float4*>(p.x + mm * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        // residual rows (L2 hits: the LayerNorm warps read them one tile earlier) are requested TWO column blocks ahead
        float4 cur[4], nx1[4], nx2[4];
        load_res(part, cur);
        load_res(part + MP_EPI_SPLIT, nx1);
        for (int cb = part; cb < (C16 >> 4); cb += MP_EPI_SPLIT) {
          load_res(cb + 2 * MP_EPI_SPLIT, nx2);
          tmem_ld16(lane_addr + (uint32_t)(cb * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = *reinterpret_cast<const float4*>(b2s + cb * 16 + j4 * 4);
            v[j4 * 4] += bb.x; v[j4 * 4 + 1] += bb.y; v[j4 * 4 + 2] += bb.z; v[j4 * 4 + 3] += bb.w;
          }
          epi_scatter16(scr, v, lane);
          const int c = cb * 16 + (lane & 3) * 4;
          if (c < C) {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const long long mm = (long long)tile * TILE_M + lg * 32 + ps * 8 + (lane >> 2);
              if (mm >= p.M) continue;
              const float4 y = epi_gather4(scr, ps, lane);
              *reinterpret_cast<float4*>(p.out + mm * C + c) = make_float4(y.x + cur[ps].x, y.y + cur[ps].y, y.z + cur[ps].z, y.w + cur[ps].w);
            }
          }
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            cur[ps] = nx1[ps];
            nx1[ps] = nx2[ps];
          }
          __syncwarp();
        }
      } else {
      final_blocks(tile, s, lg, nb_ln + part, MP_EPI_SPLIT, nb);
      }
#if SWN_MLP_PROFILE
      if (p.phase_cycles && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.phase_cycles + 10), (unsigned long long)(clock64() - t_f0));
#endif
      tc_fence_before();
      mbar_arrive_warp(&sh->y_empty[yb]);
      if (!DIRECT) mbar_arrive_warp(&sh->in_empty[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int launch_mlp_persist(MlpParams p, int num_sms, cudaStream_t stream) {
  const int C = p.C, C16 = (C + 15) & ~15;
  SWN_CHECK(p.M > 0 && C >= 4 && C % 4 == 0 && C <= 192, "mlp_persist: unsupported C=%d", C);
  const bool direct = C > 96;    // rows too wide to stage twice: LayerNorm warps read global memory directly
  const int nj = (4 * C) / p.HC;
  const int KB1 = (C16 + 63) >> 6, nkk = (p.HC + 63) >> 6;
  int cols = (direct ? 1 : 2) * ((C16 + 31) & ~31) + (nj > 1 ? 2 : 1) * p.HC, tc = 32;
  while (tc < cols) tc <<= 1;
  SWN_CHECK(tc <= 512, "mlp_persist: TMEM overflow");
  {
    const int steps2 = p.HC >> 4;   // 16-column blocks per hidden chunk; an epilogue thread keeps at most 4 in registers
    SWN_CHECK((steps2 % MP_EPI_SPLIT == 0 ? steps2 / MP_EPI_SPLIT : (steps2 % 2 == 0 ? steps2 / 2 : steps2)) <= 4,
              "mlp_persist: hidden chunk HC=%d too wide for the epilogue register tile", p.HC);
  }
  p.tmem_cols = tc;
  const int chunks = C / 4;
  p.row_stride = direct ? 0 : (chunks + ((chunks & 1) ? 0 : 1)) * 16;
  const int stage_bytes = (p.HC > p.TR ? p.HC : p.TR) * 128;
  const int fixed = 1024 + (((direct || SWN_MP_NA2) ? 2 : 1) * KB1 + 2 * nkk) * A_KBLOCK_BYTES + 2 * TILE_M * p.row_stride + (4 * C + 3 * C16) * 4 +
                    (int)sizeof(MpSmem) + 64 + (direct ? 8 * EPI_SCRATCH_BYTES + 16 : 0);
  int stages = (232448 - fixed) / stage_bytes;
  if (stages > 6) stages = 6;
  SWN_CHECK(stages >= 2, "mlp_persist: does not fit in shared memory (C=%d)", C);
  p.stages = stages;
  const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
  const int ntiles = (p.M + TILE_M - 1) / TILE_M;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MP_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  if (direct) return go(mlp_persist_kernel<32, true>);
  if (C <= 16) return go(mlp_persist_kernel<4, false>);
  if (C <= 32) return go(mlp_persist_kernel<8, false>);
  if (C <= 64) return go(mlp_persist_kernel<16, false>);
  return go(mlp_persist_kernel<32, false>);
}

}  // namespace swn
