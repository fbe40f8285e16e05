// Fused Swin MLP on tcgen05 tensor cores:   out = x + fc2( GELU( fc1( LayerNorm(x) ) ) )
// (reference: SwinTransformerBlock.norm2 + mlp + residual, SwinWNet.py:226-234,278).
//
// One CTA (10 warps) owns 128 token rows.  LayerNorm(x) is written once to shared memory as the resident
// bf16 A operand (all warps, see build_a_tile).  The 4C hidden dimension is processed in chunks of HC columns:
//     GEMM1  Hacc[128 x HC]  = A[128 x C] * W1_j^T        (TMEM, double buffered)
//     epilogue-1 (8 warps)   : +b1, GELU (tanh form with a fitted argument, packed half2: common.cuh::gelu_pack2; |d| <= 3e-5
//                              against erf-GELU before the 16-bit rounding), -> 16-bit swizzled smem tile Hs (double buffered)
//     GEMM2  Y[128 x C]     += Hs[128 x HC] * W2_j^T      (TMEM, resident across chunks)
// so the 4C-wide hidden activation never leaves the SM.  GEMM1 of chunk j+1 is issued before GEMM2 of
// chunk j, which keeps the tensor pipe busy while the epilogue warps run GELU on chunk j.  Weights are
// streamed as pre-swizzled tiles through a TMA-engine (cp.async.bulk) mbarrier ring in exactly the
// order the MMA warp consumes them.  Final epilogue: Y + b2 + x -> out (fp32 residual stream).
#include "common.cuh"
#include "kernels.h"

namespace swn {

// profiling aid (build with SWN_NVCC_EXTRA=-DSWN_MLP_PROFILE=1, tools/mlp_phase_profile.py): lane 0 of a warp adds the
// cycles spent in `stmt` to p.phase_cycles[slot]; compiled out of the product build
#ifndef SWN_MLP_PROFILE
#define SWN_MLP_PROFILE 0
#endif
#if SWN_MLP_PROFILE
#define MLP_PROF_ADD(slot, t0)                                                                                             \
  do {                                                                                                                     \
    if (p.phase_cycles && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.phase_cycles + (slot)), (unsigned long long)(clock64() - (t0))); \
  } while (0)
#define MLP_CLOCK() clock64()
#define MLP_TIMED(slot, stmt)                                                                         \
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
#define MLP_PROF_ADD(slot, t0) do { (void)(t0); } while (0)
#define MLP_CLOCK() 0ll
#define MLP_TIMED(slot, stmt) do { stmt; } while (0)
#endif

// epilogue warps per TMEM lane group (they split the 16-column blocks of a hidden chunk / of Y).  The GELU epilogue is
// bound by MUFU.TANH (8 issue cycles per warp-element per SM sub-partition, tools/micro/pipes.cu) and by dependent-issue
// latency: with 2 warps per sub-partition it ran at ~15 cycles per warp-element, with 4 it approaches the MUFU rate.
#ifndef SWN_MLP_EPI_SPLIT
#define SWN_MLP_EPI_SPLIT 4
#endif
// C = 192 runs as TWO co-resident CTAs per SM (SWN_MLP_TWO_CTA, kernels.h): 64-column hidden chunks, one hidden accumulator
// (TMEM 192 + 64 = 256 columns), 96-column fc2 tiles, two ring stages (111 KB of shared memory), 96 registers with 2 epilogue
// warps per lane group — the LayerNorm prologue and the final epilogue of one tile overlap the GEMM / GELU pipeline of the
// other: 0.671 -> 0.627 ms at 484 k rows.
constexpr int mlp_threads(int split) { return (2 + 4 * split) * 32; }

struct MlpSmem {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t hacc_full[2], hacc_empty[2], hs_full[2], hs_empty[2];
  uint64_t a_ready, y_full;
  uint32_t tmem_base;
};

template <int LPR, int KV, int SPLIT, int MINB>
__global__ void __launch_bounds__(mlp_threads(SPLIT), MINB) mlp_kernel(const MlpParams p) {
  constexpr int MLP_EPI_SPLIT = SPLIT, MLP_WARPS = 2 + 4 * SPLIT, MLP_THREADS = MLP_WARPS * 32, MLP_EPI_THREADS = 128 * SPLIT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int C = p.C, C16 = (C + 15) & ~15;
  const int KB1 = (C16 + 63) >> 6, steps1 = C16 >> 4;
  const int HC = p.HC, TR = p.TR;
  const int nj = (4 * C) / HC;
  const int nkk = (HC + 63) >> 6, steps2 = HC >> 4;
  const int nT = C16 / TR;
  const int hbase = (C16 + 31) & ~31;  // TMEM column of the first hidden accumulator
  const bool hacc1 = p.n_hacc == 1;    // single hidden accumulator (TMEM too small for two at this HC): GEMM1(j+1) waits for epilogue(j)
  const int stage_bytes = max(HC, TR) * 128;
  const int w1_bytes = HC * 128, w2_bytes = TR * 128;

  uint8_t* a_smem = smem;
  uint8_t* hs_smem = a_smem + KB1 * A_KBLOCK_BYTES;               // 2 x nkk k-blocks
  uint8_t* ring = hs_smem + 2 * nkk * A_KBLOCK_BYTES;
  float* b1s = reinterpret_cast<float*>(ring + p.stages * stage_bytes);  // [4C]
  float* b2s = b1s + 4 * C;                                              // [C16]
  MlpSmem* sh = reinterpret_cast<MlpSmem*>(b2s + C16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * TILE_M;
  const long long t_cta0 = MLP_CLOCK();

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->hacc_full[b], 1);
      mbar_init(&sh->hacc_empty[b], MLP_EPI_THREADS / 32);
      mbar_init(&sh->hs_full[b], MLP_EPI_THREADS / 32);
      mbar_init(&sh->hs_empty[b], 1);
    }
    mbar_init(&sh->y_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 4 * C; i += MLP_THREADS) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C16; i += MLP_THREADS) b2s[i] = p.b2[i];

  // ---- weight stream bookkeeping (consumption order: G1(0), [G1(j+1), G2(j)]..., G2(nj-1)) ----
  // tile t -> byte size: tiles [0,KB1) are W1; then per j: (j+1<nj ? KB1 W1 tiles : none) + nkk*nT W2 tiles
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.Wp);
  int pt = 0;            // producer tile counter (thread 0 only)
  int pj = -1, pi = 0;   // producer position: pj = -1 -> initial G1(0); else inside iteration pj, item pi
  auto next_bytes = [&]() -> int {   // size of the next tile in stream order, 0 when the stream is exhausted
    while (true) {
      if (pj < 0) {
        if (pi < KB1) { ++pi; return w1_bytes; }
        pj = 0; pi = 0;
      } else if (pj >= nj) {
        return 0;
      } else {
        const int n1 = (pj + 1 < nj) ? KB1 : 0;
        if (pi < n1) { ++pi; return w1_bytes; }
        if (pi < n1 + nkk * nT) { ++pi; return w2_bytes; }
        ++pj; pi = 0;
      }
    }
  };
  // Per-CTA setup runs in warp 0 WHILE warps 1.. build the A tile (the kernel is not persistent; with tcgen05.alloc and
  // the barrier in front of the prologue every CTA paid them serially): first ring fill (issued by the thread that
  // initialised the barriers: program order), then the TMEM allocation.
  if (warp == 0) {
    if (lane == 0) {
      for (; pt < p.stages; ++pt) {
        const int bytes = next_bytes();
        if (bytes == 0) break;
        mbar_arrive_expect_tx(&sh->full[pt], (uint32_t)bytes);
        bulk_g2s(ring + pt * stage_bytes, wsrc, (uint32_t)bytes, &sh->full[pt]);
        wsrc += bytes;
      }
    }
    __syncwarp();
    tmem_alloc(&sh->tmem_base, (uint32_t)p.tmem_cols);
  }

  // ===== prologue: LayerNorm(x) -> resident bf16 A tile (warps 1..) =====
  if (warp > 0) {
    constexpr int UNR = KV == 1 ? 4 : (KV == 2 ? 7 : 2);   // rows in flight per warp (measured: 7 helps at C = 192, hurts at C = 384)
    const float* x = p.x;
    const int M = p.M;
    build_a_tile<LPR, KV, UNR, true>(a_smem, C, C16, p.ln_w, p.ln_b, p.ln_eps, warp - 1, MLP_WARPS - 1, lane, [&](int r, int k) {
      const long long m = m0 + r;
      if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
      return *reinterpret_cast<const float4*>(x + m * C + k);
    });
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();     // barrier inits, TMEM base address, staged biases and the A tile are visible to every role
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  if (warp == 1) MLP_PROF_ADD(12, t_cta0);

  if (warp == 0) {
    // ===== weight producer: rest of the stream =====
    if (lane == 0) {
      while (true) {
        const int bytes = next_bytes();
        if (bytes == 0) break;
        const int s = pt % p.stages;
        mbar_wait(&sh->empty[s], ((uint32_t)(pt / p.stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], (uint32_t)bytes);
        bulk_g2s(ring + s * stage_bytes, wsrc, (uint32_t)bytes, &sh->full[s]);
        wsrc += bytes;
        ++pt;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit =====
    const long long t_mma0 = MLP_CLOCK();
    const uint32_t idesc1 = umma_idesc_bf16(TILE_M, (uint32_t)HC);
    const uint32_t idesc2 = umma_idesc_bf16(TILE_M, (uint32_t)TR);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(a_smem));
    const uint64_t hs_desc0 = umma_desc_sw128(smem_u32(hs_smem));
    const uint64_t ring_desc0 = umma_desc_sw128(smem_u32(ring));
    const uint32_t stage_d16 = (uint32_t)(stage_bytes >> 4), kblk_d16 = A_KBLOCK_BYTES >> 4;
    RingPos rp{0, 0u};
    auto gemm1 = [&](int j) {
      const int buf = hacc1 ? 0 : (j & 1);
      MLP_TIMED(2, mbar_wait(&sh->hacc_empty[buf], ((hacc1 ? (uint32_t)j : ((uint32_t)j >> 1)) & 1u) ^ 1u));
      const uint32_t d = tmem_base + (uint32_t)(hbase + buf * HC);
      for (int kb = 0; kb < KB1; ++kb) {
        MLP_TIMED(1, mbar_wait(&sh->full[rp.s], rp.ph));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = a_desc0 + (uint64_t)(kb * kblk_d16), bd = ring_desc0 + (uint64_t)(rp.s * stage_d16);
          const int steps = min(4, steps1 - kb * 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&sh->empty[rp.s]);
          if (kb == KB1 - 1) umma_commit(&sh->hacc_full[buf]);
        }
        __syncwarp();
        rp.next(p.stages);
      }
    };
    auto gemm2 = [&](int j) {
      const int buf = j & 1;
      MLP_TIMED(3, mbar_wait(&sh->hs_full[buf], ((uint32_t)j >> 1) & 1u));
      const uint64_t hd0 = hs_desc0 + (uint64_t)(buf * nkk * kblk_d16);
      for (int kk = 0; kk < nkk; ++kk) {
        const int steps = min(4, steps2 - kk * 4);
        for (int tt = 0; tt < nT; ++tt) {
          MLP_TIMED(4, mbar_wait(&sh->full[rp.s], rp.ph));
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = hd0 + (uint64_t)(kk * kblk_d16), bd = ring_desc0 + (uint64_t)(rp.s * stage_d16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < steps) umma_bf16(tmem_base + (uint32_t)(tt * TR), ad + 2 * k, bd + 2 * k, idesc2, (j | kk | k) != 0 ? 1u : 0u);
            umma_commit(&sh->empty[rp.s]);
            if (kk == nkk - 1 && tt == nT - 1) {
              umma_commit(&sh->hs_empty[buf]);
              if (j == nj - 1) umma_commit(&sh->y_full);
            }
          }
          __syncwarp();
          rp.next(p.stages);
        }
      }
    };
    gemm1(0);
    for (int j = 0; j < nj; ++j) {
      if (j + 1 < nj) gemm1(j + 1);
      gemm2(j);
    }
    MLP_PROF_ADD(5, t_mma0);
  } else {
    // ===== epilogue warps 2..: thread <-> row; the MLP_EPI_SPLIT warps of a lane group split the columns =====
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;      // column part of this warp, 0 .. MLP_EPI_SPLIT-1
    const int r = lg * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    // 16-column blocks of the hidden chunk owned by this warp
    const int cb_beg = half * steps2 / MLP_EPI_SPLIT;
    const int cb_end = (half + 1) * steps2 / MLP_EPI_SPLIT;
    float v[32];
    for (int j = 0; j < nj; ++j) {
      const int buf = j & 1;                       // hidden smem tile
      const int abuf = hacc1 ? 0 : buf;            // hidden accumulator in TMEM
      const uint32_t ph = ((uint32_t)j >> 1) & 1u;
      const uint32_t aph = hacc1 ? ((uint32_t)j & 1u) : ph;
      if (warp == 2) {
        MLP_TIMED(6, mbar_wait(&sh->hacc_full[abuf], aph));
        MLP_TIMED(7, mbar_wait(&sh->hs_empty[buf], ph ^ 1u));
      } else {
        mbar_wait(&sh->hacc_full[abuf], aph);
        mbar_wait(&sh->hs_empty[buf], ph ^ 1u);
      }
      tc_fence_after();
      const long long t_g0 = MLP_CLOCK();
      uint8_t* hrow = hs_smem + buf * nkk * A_KBLOCK_BYTES;
      const float* bj = b1s + j * HC;
      const uint32_t t_chunk = lane_addr + (uint32_t)(hbase + abuf * HC);
      for (int cb = cb_beg; cb < cb_end; cb += 2) {
        const bool two = cb + 1 < cb_end;
        tmem_ld16(t_chunk + cb * 16, v);
        if (two) tmem_ld16(t_chunk + cb * 16 + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          if (hb == 1 && !two) break;
          const int k = (cb + hb) * 16;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = *reinterpret_cast<const float4*>(bj + k + 4 * i);
            pk[2 * i] = gelu_pack2(v[hb * 16 + 4 * i] + bb.x, v[hb * 16 + 4 * i + 1] + bb.y);
            pk[2 * i + 1] = gelu_pack2(v[hb * 16 + 4 * i + 2] + bb.z, v[hb * 16 + 4 * i + 3] + bb.w);
          }
          uint8_t* kb_base = hrow + (k >> 6) * A_KBLOCK_BYTES;
          *reinterpret_cast<uint4*>(kb_base + sw128_offset(r, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(kb_base + sw128_offset(r, (k & 63) + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive_warp(&sh->hacc_empty[abuf]);
      mbar_arrive_warp(&sh->hs_full[buf]);
      if (warp == 2) MLP_PROF_ADD(8, t_g0);
    }
    // residual rows in the transposed ownership (common.cuh), fetched one column block ahead: the first block is
    // requested before the wait for the last GEMM2, so its latency hides behind the tail of the MMA pipeline
    auto load_res = [&](int cb, float4* xr) {
      const int c = cb * 16 + (lane & 3) * 4;
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const long long mm = m0 + lg * 32 + ps * 8 + (lane >> 2);
        xr[ps] = (cb < (C16 >> 4) && c < C && mm < p.M) ? __ldg(reinterpret_cast<const float4*>(p.x + mm * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 xr_cur[4], xr_nxt[4];
    load_res(half, xr_cur);
    if (warp == 2) MLP_TIMED(9, mbar_wait(&sh->y_full, 0)); else mbar_wait(&sh->y_full, 0);
    const long long t_f0 = MLP_CLOCK();
    tc_fence_after();
    // the hidden tiles are free by now and serve as the per-warp 2 KB transposition scratch
    uint8_t* scr = hs_smem + (warp - 2) * EPI_SCRATCH_BYTES;
    for (int cb = half; cb < (C16 >> 4); cb += MLP_EPI_SPLIT) {
      load_res(cb + MLP_EPI_SPLIT, xr_nxt);
      tmem_ld16(lane_addr + (uint32_t)(cb * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += b2s[cb * 16 + j];
      epi_scatter16(scr, v, lane);
      const int c = cb * 16 + (lane & 3) * 4;
      if (c < C) {
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const long long mm = m0 + lg * 32 + ps * 8 + (lane >> 2);
          if (mm >= p.M) continue;
          const float4 y = epi_gather4(scr, ps, lane);
          *reinterpret_cast<float4*>(p.out + mm * C + c) = make_float4(y.x + xr_cur[ps].x, y.y + xr_cur[ps].y, y.z + xr_cur[ps].z, y.w + xr_cur[ps].w);
        }
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) xr_cur[ps] = xr_nxt[ps];
      __syncwarp();
    }
    if (warp == 2) MLP_PROF_ADD(10, t_f0);
  }
  if (warp == 1) MLP_PROF_ADD(11, t_cta0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int launch_mlp(MlpParams p, cudaStream_t stream) {
  const int C = p.C, C16 = (C + 15) & ~15;
  SWN_CHECK(p.M > 0 && C >= 4 && C % 4 == 0 && C <= 384, "mlp: unsupported C=%d", C);
  SWN_CHECK(p.HC % 16 == 0 && p.HC >= 16 && p.HC <= 256 && (4 * C) % p.HC == 0, "mlp: bad HC=%d for C=%d", p.HC, C);
  SWN_CHECK(p.TR % 16 == 0 && p.TR >= 16 && p.TR <= 256 && C16 % p.TR == 0, "mlp: bad TR=%d for C=%d", p.TR, C);
  const int nj = (4 * C) / p.HC;
  const int KB1 = (C16 + 63) >> 6, nkk = (p.HC + 63) >> 6;
  p.n_hacc = nj > 1 ? 2 : 1;
  if (((C16 + 31) & ~31) + p.n_hacc * p.HC > 512) p.n_hacc = 1;
  const bool two_cta = SWN_MLP_TWO_CTA && C == 192 && p.HC == 64;
  if (two_cta) p.n_hacc = 1;
  int cols = ((C16 + 31) & ~31) + p.n_hacc * p.HC, tc = 32;
  while (tc < cols) tc <<= 1;
  SWN_CHECK(tc <= 512, "mlp: TMEM overflow (C=%d HC=%d)", C, p.HC);
  p.tmem_cols = tc;
  const int stage_bytes = (p.HC > p.TR ? p.HC : p.TR) * 128;
  const int fixed = 1024 + (KB1 + 2 * nkk) * A_KBLOCK_BYTES + (4 * C + C16) * 4 + (int)sizeof(MlpSmem) + 64;
  // aim for >= 2 co-resident CTAs per SM (smem <= ~113 KB, TMEM <= 256 columns) when >= 3 ring stages still fit
  int stages = (tc <= 256) ? (113 * 1024 - fixed) / stage_bytes : 0;
  if (stages < (two_cta ? 2 : 3)) stages = (232448 - fixed) / stage_bytes;
  else if (two_cta) stages = 2;
  if (stages > 6) stages = 6;
  SWN_CHECK(stages >= 2, "mlp: C=%d HC=%d TR=%d does not fit in shared memory", C, p.HC, p.TR);
  p.stages = stages;
  const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
  const long long grid = ((long long)p.M + TILE_M - 1) / TILE_M;
  auto go = [&](auto kern, int threads) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  constexpr int SP = SWN_MLP_EPI_SPLIT, TH = mlp_threads(SP);
  if (two_cta) return go(mlp_kernel<32, 2, 2, 2>, mlp_threads(2));
  if (C <= 16) return go(mlp_kernel<4, 1, SP, 1>, TH);
  if (C <= 32) return go(mlp_kernel<8, 1, SP, 1>, TH);
  if (C <= 64) return go(mlp_kernel<16, 1, SP, 1>, TH);
  if (C <= 128) return go(mlp_kernel<32, 1, SP, 1>, TH);
  if (C <= 256) return go(mlp_kernel<32, 2, SP, 1>, TH);
  return go(mlp_kernel<32, 3, SP, 1>, TH);
}

}  // namespace swn
