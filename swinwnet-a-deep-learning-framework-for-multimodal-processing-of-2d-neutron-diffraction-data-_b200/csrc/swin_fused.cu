// Fused W-MSA / whole SwinTransformerBlock kernel (shift_size = 0, 5x5 windows; SwinWNet.py:183-209, 236-280):
//
//   x -> LN1 -> qkv (tcgen05) -> per-(window, head) softmax(q k^T + rel-pos bias) v (mma.sync on smem tiles)
//     -> proj (tcgen05) -> + x  [-> LN2 -> fc1 (tcgen05) -> GELU -> fc2 (tcgen05) -> +]  -> out
//
// A tile is 5 consecutive windows = 125 tokens (rows wl*25 + t of a 128-row UMMA tile), so window_partition /
// window_reverse / zero padding are pure index math on the cp.async gather and the coalesced write-back; q, k, v,
// the attention probabilities, the attention output and the 4C-wide hidden activation never leave the SM.  HBM
// traffic is one fp32 read and one fp32 write of the token rows (8C bytes / token) for the whole block.
//
// Weights of the block (16-bit, pre-swizzled UMMA K-major SWIZZLE_128B images, packing.py::pack_fused_block) are
// loaded once per persistent CTA with bulk copies and stay resident in shared memory.  All phases of a tile are
// executed by all warps (epilogues split TMEM lane groups x column parts), MMAs are issued by one elected lane of
// warp 0 and tracked with mbarriers; the next tile's rows are prefetched with cp.async while the current tile
// computes.  Latency of the serial phase chain is hidden by co-resident CTAs (C <= 24) or by the prefetch.
//
// Semantics kept from the reference: the window zero padding happens AFTER norm1 (padded tokens are exact zero
// vectors, their q/k/v equal the qkv bias, they take part as keys; SwinWNet.py:242-255), q is scaled by hd^-0.5
// (folded into the packed q rows / bias together with log2(e)), the relative-position bias is added unscaled.
#include "common.cuh"
#include "kernels.h"

#ifndef SWN_FUSED_SPIN
#define SWN_FUSED_SPIN 0
#endif
// 1 = token rows of swin_fused_kernel move through the TMA engine (one cp.async.bulk per row global -> staging and
// staging -> global, issued by the 128 lanes of warps 4-7).  MEASURED SLOWER at these row sizes (48 / 96 / 192 bytes: per-copy
// overhead of the TMA engine; C = 48: 0.88 vs 0.77 ms, C = 12: 4.91 vs 4.15 ms, gpurun_out/r2g_ops.txt), so the 16-byte
// cp.async gather / float4 write-back stays the product path; at 384-byte rows (swin_attn_stream_kernel, C = 96) the
// bulk gather wins 10 % and is unconditional there.
#ifndef SWN_FUSED_BULK
#define SWN_FUSED_BULK 0
#endif

namespace swn {

namespace {

constexpr int FB_WIN = 5, FB_TOK = 25;

__device__ __forceinline__ void fb_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void fb_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// P = 2^(s - m) as a packed 16-bit pair (the A fragment of the P V mma)
__device__ __forceinline__ uint32_t fb_exp2_pack(float a, float b, float m) {
  return pack_op(ex2_approx(a - m), ex2_approx(b - m));
}

struct FbBars {
  uint64_t w, mma, g1[2], g2[2], rows[2];
  uint32_t tmem_base;
};

}  // namespace

// Geometry of the block for channel width C (mirrored by packing.py::fused_block_geometry and the launcher)
template <int C_>
struct FbGeom {
  static constexpr int C = C_;
  static constexpr int K16 = (C + 15) / 16 * 16;
  static constexpr int ONES = (3 * C + 7) / 8 * 8;          // 8 columns of ones behind q|k|v
  static constexpr int NQ = (ONES + 8 + 15) / 16 * 16;
  static constexpr int HC = (4 * C) % 64 == 0 ? 64 : 48;    // hidden chunk (columns of one GEMM1 / k-extent of one GEMM2)
  static constexpr int NJ = (4 * C) / HC;
  static constexpr int RS0 = NQ + 8;
  static constexpr int RS = ((RS0 / 8) & 1) ? RS0 : RS0 + 8;  // qkv row stride (16-bit elements): odd number of 16-B chunks
  static constexpr int RSB = ((C / 4) & 1) ? C * 4 : C * 4 + 16;  // fp32 staging row stride in bytes (odd 16-B chunks)
  static constexpr int KB = (K16 + 63) / 64;
  static_assert((4 * C) % HC == 0 && NQ <= 256 && KB == 1, "unsupported channel width");
};

// One (window, head) pair of the attention core, executed by one warp: S = Q K^T (+ bias image) on mma.sync, softmax,
// O = P V, normalised O -> the 16-bit A tile of the proj GEMM.  qkv rows live in `u_s` as [row][RS] 16-bit with
// q | k | v | ones at columns 0 | C | 2C | ONES; the window's tokens are rows wl*25 .. wl*25+24.
template <int HD, int C, int RS, int ONES>
__device__ __forceinline__ void fb_attention_pair(const uint8_t* u_s, uint8_t* a_s, const float4* biasfrag, int wl, int h,
                                                  int lane) {
  constexpr int KS = HD >= 16 ? HD / 16 : 1;   // k-steps of S = Q K^T
  constexpr int NTO = HD >= 8 ? HD / 8 : 1;    // 8-wide output column tiles of O = P V
  const int g = lane >> 2, t4 = lane & 3;
  const op_t* base = reinterpret_cast<const op_t*>(u_s) + wl * FB_TOK * RS;
  const int qoff = h * HD, koff = C + h * HD, voff = 2 * C + h * HD;
  // accumulators start from the relative-position bias image of this head (log2 domain, -1e30 on the key columns
  // 25..31 that belong to the next window / padding): the softmax argument comes straight out of the mma
  float s[2][4][4];
  {
    const float4* bf = biasfrag + h * 256 + lane;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float4 b4 = bf[(mt * 4 + nt) * 32];
        s[mt][nt][0] = b4.x; s[mt][nt][1] = b4.y; s[mt][nt][2] = b4.z; s[mt][nt][3] = b4.w;
      }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int col = ks * 16 + t4 * 2;
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const op_t* qr = base + (mt * 16 + g) * RS + qoff + col;
      a[mt][0] = col < HD ? *reinterpret_cast<const uint32_t*>(qr) : 0u;
      a[mt][1] = col < HD ? *reinterpret_cast<const uint32_t*>(qr + 8 * RS) : 0u;
      a[mt][2] = col + 8 < HD ? *reinterpret_cast<const uint32_t*>(qr + 8) : 0u;
      a[mt][3] = col + 8 < HD ? *reinterpret_cast<const uint32_t*>(qr + 8 * RS + 8) : 0u;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const op_t* kr = base + (nt * 8 + g) * RS + koff + col;
      const uint32_t b0 = col < HD ? *reinterpret_cast<const uint32_t*>(kr) : 0u;
      const uint32_t b1 = col + 8 < HD ? *reinterpret_cast<const uint32_t*>(kr + 8) : 0u;
      fb_mma(s[0][nt], a[0], b0, b1);
      fb_mma(s[1][nt], a[1], b0, b1);
    }
  }
  // P = 2^(s - rowmax) as 16-bit A fragments; the row sums come out of the P V mma through the ones block of qkv_s
  uint32_t pa[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    float m0 = fmaxf(fmaxf(s[mt][0][0], s[mt][0][1]), fmaxf(s[mt][1][0], s[mt][1][1]));
    m0 = fmaxf(m0, fmaxf(fmaxf(s[mt][2][0], s[mt][2][1]), fmaxf(s[mt][3][0], s[mt][3][1])));
    float m1 = fmaxf(fmaxf(s[mt][0][2], s[mt][0][3]), fmaxf(s[mt][1][2], s[mt][1][3]));
    m1 = fmaxf(m1, fmaxf(fmaxf(s[mt][2][2], s[mt][2][3]), fmaxf(s[mt][3][2], s[mt][3][3])));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    pa[mt][0][0] = fb_exp2_pack(s[mt][0][0], s[mt][0][1], m0);
    pa[mt][0][1] = fb_exp2_pack(s[mt][0][2], s[mt][0][3], m1);
    pa[mt][0][2] = fb_exp2_pack(s[mt][1][0], s[mt][1][1], m0);
    pa[mt][0][3] = fb_exp2_pack(s[mt][1][2], s[mt][1][3], m1);
    pa[mt][1][0] = fb_exp2_pack(s[mt][2][0], s[mt][2][1], m0);
    pa[mt][1][1] = fb_exp2_pack(s[mt][2][2], s[mt][2][3], m1);
    // key columns 24..31 (n-tile 3): the odd ones (25, 27, 29, 31) are padding for EVERY lane (bias image -1e30 -> P = 0):
    // no exponential is issued for them (the attention core is MUFU-bound: 12.5 % fewer ex2 per pair)
    pa[mt][1][2] = pack_op(ex2_approx(s[mt][3][0] - m0), 0.f);
    pa[mt][1][3] = pack_op(ex2_approx(s[mt][3][2] - m1), 0.f);
  }
  // O = P V  (+ one extra 8-column tile of ones: its accumulator is the softmax denominator of the row)
  const int vcol0 = voff & ~7;
  float o[2][NTO][4], od[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    od[mt][0] = od[mt][1] = od[mt][2] = od[mt][3] = 0.f;
#pragma unroll
    for (int n = 0; n < NTO; ++n) o[mt][n][0] = o[mt][n][1] = o[mt][n][2] = o[mt][n][3] = 0.f;
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint32_t vrow = smem_u32(base + (ks * 16 + (lane & 15)) * RS);
    uint32_t b0, b1;
#pragma unroll
    for (int n = 0; n < NTO; ++n) {
      fb_ldsm_x2_trans(b0, b1, vrow + (vcol0 + n * 8) * 2);
      fb_mma(o[0][n], pa[0][ks], b0, b1);
      fb_mma(o[1][n], pa[1][ks], b0, b1);
    }
    fb_ldsm_x2_trans(b0, b1, vrow + ONES * 2);
    fb_mma(od[0], pa[0][ks], b0, b1);
    fb_mma(od[1], pa[1][ks], b0, b1);
  }
  float inv[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    inv[mt][0] = rcp_approx(od[mt][0]);
    inv[mt][1] = rcp_approx(od[mt][2]);
  }
  // normalised O -> A tile (16-bit, swizzled), rows wl*25 + i, columns h*HD + d
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < NTO; ++n) {
      const int dcol = n * 8 + t4 * 2 - (voff - vcol0);
      if (dcol >= 0 && dcol < HD) {
        const int k = qoff + dcol;
        const int i0 = mt * 16 + g, i1 = i0 + 8;
        uint8_t* kb = a_s + (k >> 6) * A_KBLOCK_BYTES;
        if (i0 < FB_TOK)
          *reinterpret_cast<uint32_t*>(kb + sw128_offset(wl * FB_TOK + i0, k & 63)) = pack_op(o[mt][n][0] * inv[mt][0], o[mt][n][1] * inv[mt][0]);
        if (i1 < FB_TOK)
          *reinterpret_cast<uint32_t*>(kb + sw128_offset(wl * FB_TOK + i1, k & 63)) = pack_op(o[mt][n][2] * inv[mt][1], o[mt][n][3] * inv[mt][1]);
      }
    }
}

template <int CC, int NH, int NT, int MINB, bool PROF>
__global__ void __launch_bounds__(NT, MINB) swin_fused_kernel(const FusedBlockParams p) {
  using G = FbGeom<CC>;
  constexpr int HD = CC / NH;
  constexpr int NP = NT / 128;               // column parts per TMEM lane group
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  constexpr int C = G::C, K16 = G::K16, NQ = G::NQ, HC = G::HC, nj = G::NJ, nH = NH, RS = G::RS, RSB = G::RSB;
  constexpr int C4 = C >> 2, KB = G::KB, ksteps = K16 >> 4, ONES = G::ONES;
  constexpr int HCp = (HC + 31) & ~31;
  constexpr int NHS = C == 24 ? 1 : nj;                              // hidden-tile buffers (C = 24: smem for 2 CTAs / SM)
  constexpr int REG0 = ((NQ > nj * HCp ? NQ : nj * HCp) + 31) / 32 * 32;
  constexpr int TMY = REG0;                                          // TMEM: [0, REG0) qkv / hidden accumulators, then Y
  constexpr int TMEM_COLS = REG0 + K16 <= 32 ? 32 : REG0 + K16 <= 64 ? 64 : REG0 + K16 <= 128 ? 128 : REG0 + K16 <= 256 ? 256 : 512;

  // ---- shared memory carve-up (offsets computed by the launcher) ----
  uint8_t* wq_s = smem;                                   // [KB][NQ x 64]
  uint8_t* wp_s = wq_s + KB * NQ * 128;                   // [KB][K16 x 64]
  uint8_t* w1_s = wp_s + KB * K16 * 128;                  // [nj][KB][HC x 64]
  uint8_t* w2_s = w1_s + nj * KB * HC * 128;              // [nj][K16 x 64]
  uint8_t* a_s = smem + p.off_a;                          // A tile: [KB][128 x 64]
  uint8_t* u_s = smem + p.off_u;                          // qkv rows [132][RS]  |  hidden tiles [n_hs][128 x 64]
  uint8_t* stg = smem + p.off_stage;                      // [2][128][RSB] fp32 token rows (x, then x1, then out)
  float* f_s = reinterpret_cast<float*>(smem + p.off_f);
  const float* bqkv = f_s;                                // [NQ]  qkv bias with norm1's affine folded in (+ ones block)
  const float* bqkv_pad = bqkv + NQ;                      // [NQ]  plain qkv bias: q/k/v of zero-padded tokens
  const float* bproj = bqkv_pad + NQ;                     // [K16]
  const float* b1 = bproj + K16;                          // [4C]  fc1 bias with norm2's affine folded in
  const float* b2 = b1 + 4 * C;                           // [K16]
  const float4* biasfrag = reinterpret_cast<const float4*>(b2 + K16);   // [nH][8][32] accumulator-fragment bias images
  int* tok_s = reinterpret_cast<int*>(smem + p.off_misc);             // [3][128]
  float2* red = reinterpret_cast<float2*>(tok_s + 3 * 128);           // [NP][128]
  FbBars* bars = reinterpret_cast<FbBars*>(red + NP * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7;

  if (tid == 0) {
    mbar_init(&bars->w, 1);
    mbar_init(&bars->mma, 1);
    mbar_init(&bars->g1[0], 1);
    mbar_init(&bars->g1[1], 1);
    mbar_init(&bars->g2[0], 1);
    mbar_init(&bars->g2[1], 1);
    mbar_init(&bars->rows[0], 128);
    mbar_init(&bars->rows[1], 128);
    fence_barrier_init();
  }
  for (int i = tid * 16; i < KB * A_KBLOCK_BYTES; i += NT * 16) *reinterpret_cast<uint4*>(a_s + i) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid * 16; i < p.u_bytes; i += NT * 16) *reinterpret_cast<uint4*>(u_s + i) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < p.nf; i += NT) f_s[i] = p.fpk[i];
  if (warp == 0) tmem_alloc(&bars->tmem_base, (uint32_t)TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  if (tid == 0) {
    // resident weights: one expect_tx, bulk copies of <= 16 KB
    mbar_arrive_expect_tx(&bars->w, (uint32_t)p.w_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wpk);
    for (int o = 0; o < p.w_bytes; o += 16384) bulk_g2s(smem + o, src + o, (uint32_t)min(16384, p.w_bytes - o), &bars->w);
  }

  const uint32_t idesc_q = umma_idesc_bf16(TILE_M, (uint32_t)(NQ > 256 ? NQ / 2 : NQ));
  const uint32_t idesc_c = umma_idesc_bf16(TILE_M, (uint32_t)K16);
  const uint32_t idesc_h = umma_idesc_bf16(TILE_M, (uint32_t)HC);
  const uint64_t a_desc = umma_desc_sw128(smem_u32(a_s));
  const int nWin2 = p.nWy * p.nWx;

  auto calc_tok = [&](int tile, int slot) {
    if (tid < 128) {
      int tok = -1;
      const int wl = tid / FB_TOK, t = tid - wl * FB_TOK;
      const long long w = (long long)tile * FB_WIN + wl;
      if (wl < FB_WIN && w < p.n_windows) {
        const int b = (int)(w / nWin2);
        const int wr = (int)(w - (long long)b * nWin2);
        const int wy = wr / p.nWx, wx = wr - wy * p.nWx;
        const int Y = wy * 5 + t / 5, X = wx * 5 + t % 5;
        if (Y < p.H && X < p.W) tok = (b * p.H + Y) * p.W + X;
      }
      tok_s[slot * 128 + tid] = tok;
    }
  };
  // 16-byte chunk q = tid + i*NT of the [128 x C] tile <-> (row, chunk-in-row), advanced without divisions
  const int ch_r0 = tid / C4, ch_c0 = tid - ch_r0 * C4;
  constexpr int ch_rstep = NT / C4, ch_cstep = NT - ch_rstep * C4;
  auto issue_loads = [&](int slot, int buf) {
#if SWN_FUSED_BULK
    if (warp >= 4 && warp < 8) {
      const int r = tid - 128;
      bulk_wait_read();            // this lane's write-back of the tile that used the buffer has left shared memory
      const int tok = tok_s[slot * 128 + r];
      if (tok >= 0) {
        mbar_arrive_expect_tx(&bars->rows[buf], (uint32_t)(C * 4));
        bulk_g2s(stg + buf * 128 * RSB + r * RSB, p.x + (long long)tok * C, (uint32_t)(C * 4), &bars->rows[buf]);
      } else {
        mbar_arrive(&bars->rows[buf]);   // rows of invalid tokens keep stale bytes: every consumer masks them (tok < 0)
      }
    }
    return;
#endif
    const uint32_t dst0 = smem_u32(stg + buf * 128 * RSB);
    int r = ch_r0, c4 = ch_c0;
    while (r < 128) {
      const int tok = tok_s[slot * 128 + r];
      const float* src = tok >= 0 ? p.x + (long long)tok * C + c4 * 4 : p.x;
      cp_async16(dst0 + r * RSB + c4 * 16, src, tok >= 0 ? 16u : 0u);
      c4 += ch_cstep;
      r += ch_rstep;
      if (c4 >= C4) {
        c4 -= C4;
        ++r;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // (x - mean) * rstd of the staged rows -> 16-bit A tile (the LayerNorm affine is folded into the packed weights;
  // rows of invalid tokens are exact zeros)
  auto normalize_rows = [&](const uint8_t* srow, float mean, float rstd, bool valid) {
    const float nm = -mean * rstd;
    _Pragma("unroll")
    for (int u_i = 0; u_i < ((K16 >> 3) + NP - 1) / NP; ++u_i) {
      const int u = part + u_i * NP;
      if (u >= (K16 >> 3)) break;
      uint32_t pk[4] = {0u, 0u, 0u, 0u};
      if (valid) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int c = u * 8 + hh * 4;
          if (c < C) {
            const float4 v = *reinterpret_cast<const float4*>(srow + c * 4);
            pk[hh * 2] = pack_op(fmaf(v.x, rstd, nm), fmaf(v.y, rstd, nm));
            pk[hh * 2 + 1] = pack_op(fmaf(v.z, rstd, nm), fmaf(v.w, rstd, nm));
          }
        }
      }
      const int k = u * 8;
      *reinterpret_cast<uint4*>(a_s + (k >> 6) * A_KBLOCK_BYTES + sw128_offset(row, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  };

  // phase parities of the mbarriers, one bit each (0: mma, 1-2: g1[0..1], 3-4: g2[0..1]); every completed phase of
  // every barrier is waited for exactly once by every thread, so the bits stay in step with the barriers
  uint32_t phases = 0;
  auto wait_bar = [&](uint64_t* bar, int bit) {   // warp 0 polls the mbarrier, the CTA barrier releases everybody else
#if SWN_FUSED_SPIN
    if (warp == 0) mbar_wait_spin(bar, (phases >> bit) & 1u);
#else
    if (warp == 0) mbar_wait(bar, (phases >> bit) & 1u);
#endif
    phases ^= 1u << bit;
  };
  constexpr float inv_c = 1.0f / (float)C;
  int it = 0;
  // optional phase profile (tools/phase_profile.py): thread 0 accumulates the cycles between consecutive marks
  __shared__ long long ph_acc[PROF ? 16 : 1];
  long long ph_t = 0;
  const bool prof = PROF && tid == 0;
  auto mark = [&](int i) {
    if constexpr (PROF) if (prof) {
      const long long t = clock64();
      ph_acc[i] += t - ph_t;
      ph_t = t;
    }
  };
  if constexpr (PROF) if (prof) {
    for (int i = 0; i < 16; ++i) ph_acc[i] = 0;
    ph_t = clock64();
  }
  const bool prefetch = p.n_stage == 2;
  uint32_t rows_ph = 0;        // phase parity of rows[0] / rows[1], one bit each
  calc_tok(blockIdx.x, 0);     // grid <= ntiles
  __syncthreads();
  if (prefetch) issue_loads(0, 0);

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = prefetch ? (it & 1) : 0, slot = it % 3;
    const int next = tile + gridDim.x;
    const bool has_next = next < p.ntiles;
    if (has_next) calc_tok(next, (it + 1) % 3);
    __syncthreads();   // previous tile's write-back finished everywhere; tok of the next tile visible
    mark(0);
#if SWN_FUSED_BULK
    if (prefetch) {
      if (has_next) issue_loads((it + 1) % 3, buf ^ 1);
    } else {
      issue_loads(slot, 0);
    }
    mbar_wait(&bars->rows[buf], (rows_ph >> buf) & 1u);   // this tile's rows are in stg[buf]
    rows_ph ^= 1u << buf;
#else
    if (prefetch) {
      if (has_next) {
        issue_loads((it + 1) % 3, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    } else {
      issue_loads(slot, 0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();   // this tile's rows are in stg[buf]
#endif
    mark(1);

    uint8_t* stile = stg + buf * 128 * RSB;
    uint8_t* srow = stile + row * RSB;
    const int tokr = tok_s[slot * 128 + row];
    const float x0 = *reinterpret_cast<const float*>(srow);

    // ---------------- LN1 ----------------
    {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < (C4 + NP - 1) / NP; ++ci) {
        const int c = (part + ci * NP) * 4;
        if (c >= C) break;
        const float4 v = *reinterpret_cast<const float4*>(srow + c * 4);
        const float d0 = v.x - x0, d1 = v.y - x0, d2 = v.z - x0, d3 = v.w - x0;
        s1 += (d0 + d1) + (d2 + d3);
        s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
      }
      red[part * 128 + row] = make_float2(s1, s2);
    }
    __syncthreads();
    mark(2);
    {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) {
        const float2 r2 = red[pp * 128 + row];
        t1 += r2.x;
        t2 += r2.y;
      }
      const float m1 = t1 * inv_c;
      const float rstd = rsqrtf(fmaxf(t2 * inv_c - m1 * m1, 0.f) + p.eps);
      normalize_rows(srow, x0 + m1, rstd, tokr >= 0);
    }
    fence_proxy_async();
    __syncthreads();
    mark(3);

    // ---------------- qkv = LN1(x) Wqkv^T  (TMEM columns [0, NQ)) ----------------
    if (warp == 0) {
      if (it == 0) mbar_wait(&bars->w, 0u);
      tc_fence_after();
      if (elect_one()) {
        const int nchunk = NQ > 256 ? 2 : 1, ncols = NQ / nchunk;
        for (int ch = 0; ch < nchunk; ++ch)
#pragma unroll
  #pragma unroll
        for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = a_desc + (uint64_t)((k >> 2) * (A_KBLOCK_BYTES >> 4) + (k & 3) * 2);
            const uint64_t bd = umma_desc_sw128(smem_u32(wq_s + (k >> 2) * NQ * 128 + ch * ncols * 128)) + (uint64_t)((k & 3) * 2);
            umma_bf16(tmem_base + (uint32_t)(ch * ncols), ad, bd, idesc_q, k != 0 ? 1u : 0u);
          }
        umma_commit(&bars->mma);
      }
      __syncwarp();
    }
    wait_bar(&bars->mma, 0);
    __syncthreads();
    mark(4);
    tc_fence_after();
    {
      op_t* qrow = reinterpret_cast<op_t*>(u_s) + row * RS;
      float v[16];
      _Pragma("unroll")
      for (int cu_i = 0; cu_i < ((NQ >> 4) + NP - 1) / NP; ++cu_i) {
        const int cu = part + cu_i * NP;
        if (cu >= (NQ >> 4)) break;
        tmem_ld16(lane_addr + (uint32_t)(cu * 16), v);
        tmem_ld_wait();
        const float* bb = (tokr >= 0 ? bqkv : bqkv_pad) + cu * 16;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // bias as float4: scalar broadcast loads cost a shared-memory wavefront each
          const float4 b4 = *reinterpret_cast<const float4*>(bb + 4 * i);
          pk[2 * i] = pack_op(v[4 * i] + b4.x, v[4 * i + 1] + b4.y);
          pk[2 * i + 1] = pack_op(v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
        }
        *reinterpret_cast<uint4*>(qrow + cu * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(qrow + cu * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
    __syncthreads();
    mark(5);

    // ---------------- window attention core: one (window, head) pair per warp pass ----------------
    for (int pr = warp; pr < FB_WIN * nH; pr += NT / 32) {
      const int wl = pr / nH, h = pr - wl * nH;
      if ((long long)tile * FB_WIN + wl >= p.n_windows) continue;
      fb_attention_pair<HD, C, RS, ONES>(u_s, a_s, biasfrag, wl, h, lane);
    }
    fence_proxy_async();
    __syncthreads();
    mark(6);

    // ---------------- proj: Y = O Wproj^T (TMEM columns [tm_y, tm_y + K16)) ----------------
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = a_desc + (uint64_t)((k >> 2) * (A_KBLOCK_BYTES >> 4) + (k & 3) * 2);
          const uint64_t bd = umma_desc_sw128(smem_u32(wp_s + (k >> 2) * K16 * 128)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem_base + (uint32_t)TMY, ad, bd, idesc_c, k != 0 ? 1u : 0u);
        }
        umma_commit(&bars->mma);
      }
      __syncwarp();
    }
    wait_bar(&bars->mma, 0);
    __syncthreads();
    mark(7);
    tc_fence_after();
    {
      // x1 = x + proj + bias, written back to the staging row; partial LN2 moments on the fly
      float s1 = 0.f, s2 = 0.f;
      float v[16];
      _Pragma("unroll")
      for (int cu_i = 0; cu_i < ((K16 >> 4) + NP - 1) / NP; ++cu_i) {
        const int cu = part + cu_i * NP;
        if (cu >= (K16 >> 4)) break;
        tmem_ld16(lane_addr + (uint32_t)(TMY + cu * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const int c = cu * 16 + j4;
          if (c < C) {
            const float4 xr = *reinterpret_cast<const float4*>(srow + c * 4);
            const float4 bb = *reinterpret_cast<const float4*>(bproj + c);
            float4 o4;
            o4.x = xr.x + v[j4 + 0] + bb.x;
            o4.y = xr.y + v[j4 + 1] + bb.y;
            o4.z = xr.z + v[j4 + 2] + bb.z;
            o4.w = xr.w + v[j4 + 3] + bb.w;
            *reinterpret_cast<float4*>(srow + c * 4) = o4;
            const float d0 = o4.x - x0, d1 = o4.y - x0, d2 = o4.z - x0, d3 = o4.w - x0;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
          }
        }
      }
      red[part * 128 + row] = make_float2(s1, s2);
    }
    tc_fence_before();
#if SWN_FUSED_BULK
    if (!p.do_mlp) fence_proxy_async();   // the staged rows are the result: read by the bulk write-back (async proxy)
#endif
    __syncthreads();
    mark(8);

    if (p.do_mlp) {
      // ---------------- LN2 ----------------
      {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
          const float2 r2 = red[pp * 128 + row];
          t1 += r2.x;
          t2 += r2.y;
        }
        const float m1 = t1 * inv_c;
        const float rstd = rsqrtf(fmaxf(t2 * inv_c - m1 * m1, 0.f) + p.eps);
        normalize_rows(srow, x0 + m1, rstd, true);
      }
      fence_proxy_async();
      __syncthreads();
      mark(9);

      // ---------------- MLP: all GEMM1 chunks are issued at once (TMEM columns [j*HCp, j*HCp + HC)); the GELU epilogue
      // of chunk j fills hidden tile j and GEMM2(j) accumulates Y behind it, so only two MMA round trips are exposed ----
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < nj; ++j)
#pragma unroll
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = a_desc + (uint64_t)((k >> 2) * (A_KBLOCK_BYTES >> 4) + (k & 3) * 2);
              const uint64_t bd = umma_desc_sw128(smem_u32(w1_s + (j * KB + (k >> 2)) * HC * 128)) + (uint64_t)((k & 3) * 2);
              umma_bf16(tmem_base + (uint32_t)(j * HCp), ad, bd, idesc_h, k != 0 ? 1u : 0u);
            }
          umma_commit(&bars->g1[0]);
        }
        __syncwarp();
      }
      wait_bar(&bars->g1[0], 1);
      __syncthreads();
      mark(10);
      tc_fence_after();
      if constexpr (NHS == nj && nj > 1) {
        // one hidden tile per chunk: all GELU epilogues back to back (TMEM loads of every chunk in flight before the
        // first GELU), one barrier, then all GEMM2 chunks
        constexpr int UPT = ((HC >> 4) + NP - 1) / NP;     // 16-column units per thread and chunk
        float v[nj * UPT][16];
#pragma unroll
        for (int j = 0; j < nj; ++j)
#pragma unroll
          for (int ui = 0; ui < UPT; ++ui) {
            const int cu = part + ui * NP;
            if (cu < (HC >> 4)) tmem_ld16(lane_addr + (uint32_t)(j * HCp + cu * 16), v[j * UPT + ui]);
          }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < nj; ++j)
#pragma unroll
          for (int ui = 0; ui < UPT; ++ui) {
            const int cu = part + ui * NP;
            if (cu < (HC >> 4)) {
              const float* bj = b1 + j * HC + cu * 16;
              const float* vv = v[j * UPT + ui];
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b4 = *reinterpret_cast<const float4*>(bj + 4 * i);
                pk[2 * i] = gelu_pack2(vv[4 * i] + b4.x, vv[4 * i + 1] + b4.y);
                pk[2 * i + 1] = gelu_pack2(vv[4 * i + 2] + b4.z, vv[4 * i + 3] + b4.w);
              }
              uint8_t* hs = u_s + j * A_KBLOCK_BYTES;
              *reinterpret_cast<uint4*>(hs + sw128_offset(row, cu * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(hs + sw128_offset(row, cu * 16 + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        mark(11);
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < nj; ++j) {
              const uint64_t hd0 = umma_desc_sw128(smem_u32(u_s + j * A_KBLOCK_BYTES));
              const uint64_t bd0 = umma_desc_sw128(smem_u32(w2_s + j * K16 * 128));
#pragma unroll
              for (int k = 0; k < (HC >> 4); ++k)
                umma_bf16(tmem_base + (uint32_t)TMY, hd0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), idesc_c, (j | k) != 0 ? 1u : 0u);
            }
            umma_commit(&bars->g2[0]);
          }
          __syncwarp();
        }
      } else {
#pragma unroll
      for (int j = 0; j < nj; ++j) {
        if (j > 0) {   // single hidden tile: wait until GEMM2(j-1) has consumed it
          wait_bar(&bars->g2[0], 3);
          __syncthreads();
          mark(12);
        }
        uint8_t* hs = u_s;
        {
          float v[16];
          const float* bj = b1 + j * HC;
          _Pragma("unroll")
          for (int cu_i = 0; cu_i < ((HC >> 4) + NP - 1) / NP; ++cu_i) {
            const int cu = part + cu_i * NP;
            if (cu >= (HC >> 4)) break;
            tmem_ld16(lane_addr + (uint32_t)(j * HCp + cu * 16), v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = *reinterpret_cast<const float4*>(bj + cu * 16 + 4 * i);
              pk[2 * i] = gelu_pack2(v[4 * i] + b4.x, v[4 * i + 1] + b4.y);
              pk[2 * i + 1] = gelu_pack2(v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
            }
            *reinterpret_cast<uint4*>(hs + sw128_offset(row, cu * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(hs + sw128_offset(row, cu * 16 + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        mark(13);
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
            const uint64_t hd0 = umma_desc_sw128(smem_u32(hs));
            const uint64_t bd0 = umma_desc_sw128(smem_u32(w2_s + j * K16 * 128));
#pragma unroll
            for (int k = 0; k < (HC >> 4); ++k)
              umma_bf16(tmem_base + (uint32_t)TMY, hd0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), idesc_c, (j | k) != 0 ? 1u : 0u);
            umma_commit(&bars->g2[0]);
          }
          __syncwarp();
        }
      }
      }
      wait_bar(&bars->g2[0], 3);
      __syncthreads();
      mark(14);
      tc_fence_after();
      {
        float v[16];
        _Pragma("unroll")
        for (int cu_i = 0; cu_i < ((K16 >> 4) + NP - 1) / NP; ++cu_i) {
          const int cu = part + cu_i * NP;
          if (cu >= (K16 >> 4)) break;
          tmem_ld16(lane_addr + (uint32_t)(TMY + cu * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const int c = cu * 16 + j4;
            if (c < C) {
              const float4 xr = *reinterpret_cast<const float4*>(srow + c * 4);
              const float4 bb = *reinterpret_cast<const float4*>(b2 + c);
              *reinterpret_cast<float4*>(srow + c * 4) =
                  make_float4(xr.x + v[j4 + 0] + bb.x, xr.y + v[j4 + 1] + bb.y, xr.z + v[j4 + 2] + bb.z, xr.w + v[j4 + 3] + bb.w);
            }
          }
        }
      }
      tc_fence_before();
#if SWN_FUSED_BULK
      fence_proxy_async();
#endif
      __syncthreads();
      mark(15);
    }

    // ---------------- write-back of the valid token rows ----------------
#if SWN_FUSED_BULK
    if (warp >= 4 && warp < 8) {
      const int r = tid - 128;
      const int tok = tok_s[slot * 128 + r];
      if (tok >= 0) bulk_s2g(p.out + (long long)tok * C, stile + r * RSB, (uint32_t)(C * 4));
      bulk_commit();
    }
#else
    {
      int r = ch_r0, c4 = ch_c0;
      while (r < 128) {
        const int tok = tok_s[slot * 128 + r];
        if (tok >= 0)
          *reinterpret_cast<float4*>(p.out + (long long)tok * C + c4 * 4) = *reinterpret_cast<const float4*>(stile + r * RSB + c4 * 16);
        c4 += ch_cstep;
        r += ch_rstep;
        if (c4 >= C4) {
          c4 -= C4;
          ++r;
        }
      }
    }
#endif
  }

  if constexpr (PROF) if (prof)
    for (int i = 0; i < 16; ++i) p.phase_cycles[(long long)blockIdx.x * 16 + i] += ph_acc[i];
#if SWN_FUSED_BULK
  bulk_wait_all();             // outstanding write-backs read this CTA's shared memory
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)TMEM_COLS);
  }
}

// =================================================================================================================
// Streamed-weight variant for C = 96 (heads 3 or 6): the attention half  out = x + proj(W-MSA(LN1(x)))  only (the MLP
// half runs in mlp_persist.cu).  At this width the weights (100 KB) and the q|k|v rows (82 KB) do not fit next to each
// other, so the six weight tiles of a token tile [Wqkv rows 0-159 | 160-303] x [k-block 0 | 1], Wproj x [kb 0 | 1]
// stream through a 4-slot ring.  The thread that issues the MMAs also issues the bulk copies, at the two points of
// the tile where program order proves the slots free (after the qkv MMAs and after the proj MMAs have retired), so
// the ring needs no "empty" barriers and the next tile's qkv weights are in flight during the attention phase.
// The q|k|v buffer doubles as the fp32 staging buffer of the token rows: it is free from the end of the attention
// phase of tile i to the qkv epilogue of tile i+1, which is when the rows of tile i+1 are gathered into it with
// cp.async (behind the proj MMA and epilogue of tile i).  The residual of tile i is re-read from global memory (L2
// hit) into registers before the proj MMA is waited for, and the result is written straight from the epilogue
// registers.  The qkv bias rides in the GEMM: A carries two indicator columns (valid token / zero-padded token)
// behind its 96 channels and the packed weights carry the folded bias / the plain bias in those two k positions.
// =================================================================================================================
namespace {
struct FsBars {
  uint64_t full[4], mma, rows;
  uint32_t tmem_base;
};
constexpr int FS_C = 96, FS_K16 = 96, FS_ONES = 288, FS_NQ = 304, FS_NQ0 = 160, FS_NQ1 = 144, FS_RS = 312;
constexpr int FS_STAGE = FS_NQ0 * 128, FS_NSTG = 4, FS_TMY = 320, FS_TMEM = 512;
constexpr int FS_RSX = FS_C * 4 + 16;          // fp32 staging row stride in bytes (25 16-byte chunks: odd)
constexpr int FS_KQ = 112;                     // k extent of the qkv GEMM: 96 channels + the two bias indicator columns
constexpr int FS_U_BYTES = (132 * FS_RS * 2 + 1023) / 1024 * 1024;
__host__ __device__ constexpr int fs_w_tile_bytes(int t) { return (t < 2 ? FS_NQ0 : (t < 4 ? FS_NQ1 : FS_C)) * 128; }
__host__ __device__ constexpr int fs_w_tile_off(int t) {
  return (t < 2 ? t * FS_NQ0 : (t < 4 ? 2 * FS_NQ0 + (t - 2) * FS_NQ1 : 2 * FS_NQ0 + 2 * FS_NQ1 + (t - 4) * FS_C)) * 128;
}
}  // namespace

template <int NH, bool PROF>
__global__ void __launch_bounds__(512, 1) swin_attn_stream_kernel(const FusedBlockParams p) {
  constexpr int C = FS_C, HD = C / NH, NT = 512, NP = 4, RS = FS_RS, NQ = FS_NQ;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* a_s = smem;                                     // A tile: 2 k-blocks [128 x 64]
  uint8_t* u_s = a_s + 2 * A_KBLOCK_BYTES;                 // q|k|v|ones rows [132][RS]
  uint8_t* ring = u_s + FS_U_BYTES;                        // FS_NSTG weight slots
  float* f_s = reinterpret_cast<float*>(ring + FS_NSTG * FS_STAGE);
  const float* bproj = f_s;
  const float4* biasfrag = reinterpret_cast<const float4*>(bproj + FS_K16);
  int* tok_s = reinterpret_cast<int*>(f_s + FS_K16 + NH * 1024);   // [2][128]
  float2* red = reinterpret_cast<float2*>(tok_s + 2 * 128);                  // [4][128] (mean, M2) partials
  FsBars* bars = reinterpret_cast<FsBars*>(red + 4 * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7;
  if (tid == 0) {
    for (int i = 0; i < FS_NSTG; ++i) mbar_init(&bars->full[i], 1);
    mbar_init(&bars->mma, 1);
    mbar_init(&bars->rows, 128);
    fence_barrier_init();
  }
  for (int i = tid * 16; i < 2 * A_KBLOCK_BYTES + FS_U_BYTES; i += NT * 16) *reinterpret_cast<uint4*>(a_s + i) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < FS_K16 + NH * 1024; i += NT) f_s[i] = p.fpk[i];
  if (warp == 0) tmem_alloc(&bars->tmem_base, FS_TMEM);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint64_t a_desc = umma_desc_sw128(smem_u32(a_s));
  const uint32_t idesc_q0 = umma_idesc_bf16(TILE_M, FS_NQ0), idesc_q1 = umma_idesc_bf16(TILE_M, FS_NQ1);
  const uint32_t idesc_c = umma_idesc_bf16(TILE_M, FS_K16);
  const int nWin2 = p.nWy * p.nWx;

  // weight stream bookkeeping (warp 0 only): n_push / n_use count weight tiles since kernel start; slot = n % 4,
  // full-barrier parity = (n / 4) & 1; the tile inside the 6-tile cycle is n % 6.
  uint32_t n_push = 0, n_use = 0;
  auto push = [&](int count) {   // called by every thread at the same program points; warp 0 keeps the counters
    if (warp == 0) {
      if (lane == 0) {
        for (int i = 0; i < count; ++i) {
          const uint32_t n = n_push + (uint32_t)i;
          const int t = (int)(n % 6u), slot = (int)(n & 3u);
          mbar_arrive_expect_tx(&bars->full[slot], (uint32_t)fs_w_tile_bytes(t));
          bulk_g2s(ring + slot * FS_STAGE, reinterpret_cast<const uint8_t*>(p.Wpk) + fs_w_tile_off(t), (uint32_t)fs_w_tile_bytes(t),
                   &bars->full[slot]);
        }
      }
      __syncwarp();
      n_push += (uint32_t)count;
    }
  };
  auto calc_tok = [&](int tile, int slot) {
    if (tid < 128) {
      int tok = -1;
      const int wl = tid / FB_TOK, t = tid - wl * FB_TOK;
      const long long w = (long long)tile * FB_WIN + wl;
      if (wl < FB_WIN && w < p.n_windows) {
        const int b = (int)(w / nWin2);
        const int wr = (int)(w - (long long)b * nWin2);
        const int wy = wr / p.nWx, wx = wr - wy * p.nWx;
        const int Y = wy * 5 + t / 5, X = wx * 5 + t % 5;
        if (Y < p.H && X < p.W) tok = (b * p.H + Y) * p.W + X;
      }
      tok_s[slot * 128 + tid] = tok;
    }
  };
  // gather the fp32 rows of a tile into the staging view of u_s ([128][FS_RSX] bytes): ONE bulk copy (TMA engine) per
  // token row, issued by the 128 lanes of warps 4-7 — no LSU instructions, no address math in the other warps (the
  // 16-byte cp.async gather this replaces cost 15 % of a tile: 3072 LDGSTS per tile through the LSU pipe).  Every lane
  // arrives once on `rows` (count 128), valid rows add their 384 bytes to the transaction count; rows of invalid tokens
  // keep stale bytes, which every consumer masks with tok >= 0.
  auto issue_loads = [&](int slot) {
    if (warp >= 4 && warp < 8) {
      const int r = tid - 128;
      const int tok = tok_s[slot * 128 + r];
      if (tok >= 0) {
        mbar_arrive_expect_tx(&bars->rows, (uint32_t)(C * 4));
        bulk_g2s(u_s + r * FS_RSX, p.x + (long long)tok * C, (uint32_t)(C * 4), &bars->rows);
      } else {
        mbar_arrive(&bars->rows);
      }
    }
  };
  uint32_t ph_rows = 0;
  uint32_t ph_mma = 0;
  push(4);                      // the first tile's qkv weights
  calc_tok(blockIdx.x, 0);      // grid <= ntiles
  __syncthreads();
  issue_loads(0);

  // optional phase profile: thread 0 accumulates the cycles between consecutive marks
  __shared__ long long ph_acc[PROF ? 16 : 1];
  long long ph_t = 0;
  const bool prof = PROF && tid == 0;
  auto mark = [&](int i) {
    if constexpr (PROF) if (prof) {
      const long long t = clock64();
      ph_acc[i] += t - ph_t;
      ph_t = t;
    }
  };
  if constexpr (PROF) if (prof) {
    for (int i = 0; i < 16; ++i) ph_acc[i] = 0;
    ph_t = clock64();
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int slot = it & 1;
    const int next = tile + gridDim.x;
    const bool has_next = next < p.ntiles;
    const int tokr = tok_s[slot * 128 + row];
    // ---------------- LN1 from the staged rows: this thread owns the 16-column units {part, part + 4} ----------
    mbar_wait(&bars->rows, ph_rows);
    ph_rows ^= 1u;
    mark(0);
    float xv[2][16];
    constexpr int NU = C / 16;   // 6 units per row
#pragma unroll
    for (int ui = 0; ui < 2; ++ui) {
      const int cu = part + ui * NP;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cu < NU) v4 = *reinterpret_cast<const float4*>(u_s + row * FS_RSX + (cu * 16 + j4 * 4) * 4);
        xv[ui][j4 * 4 + 0] = v4.x; xv[ui][j4 * 4 + 1] = v4.y; xv[ui][j4 * 4 + 2] = v4.z; xv[ui][j4 * 4 + 3] = v4.w;
      }
    }
    if (has_next) calc_tok(next, slot ^ 1);
    {
      const int nloc = part + NP < NU ? 32 : 16;
      float sm = 0.f;
#pragma unroll
      for (int ui = 0; ui < 2; ++ui)
#pragma unroll
        for (int i = 0; i < 16; ++i) sm += xv[ui][i];     // absent units are zeros
      const float mloc = sm / (float)nloc;
      float m2 = 0.f;
#pragma unroll
      for (int ui = 0; ui < 2; ++ui)
        if (part + ui * NP < NU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) m2 = fmaf(xv[ui][i] - mloc, xv[ui][i] - mloc, m2);
        }
      red[part * 128 + row] = make_float2(mloc, m2);
    }
    __syncthreads();
    mark(1);
    {
      // Chan et al. combination of the four partial (mean, M2) pairs (sizes 32, 32, 16, 16)
      float mean = 0.f;
      float2 r2[NP];
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) {
        r2[pp] = red[pp * 128 + row];
        mean += r2[pp].x * (pp + NP < NU ? 32.f : 16.f);
      }
      mean *= (1.0f / C);
      float m2 = 0.f;
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) m2 += r2[pp].y + (pp + NP < NU ? 32.f : 16.f) * (r2[pp].x - mean) * (r2[pp].x - mean);
      const float rstd = rsqrtf(m2 * (1.0f / C) + p.eps);
      const float nm = -mean * rstd;
#pragma unroll
      for (int ui = 0; ui < 2; ++ui) {
        const int cu = part + ui * NP;
        if (cu < NU) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = tokr >= 0 ? pack_op(fmaf(xv[ui][2 * i], rstd, nm), fmaf(xv[ui][2 * i + 1], rstd, nm)) : 0u;
          const int k = cu * 16;
          uint8_t* kb = a_s + (k >> 6) * A_KBLOCK_BYTES;
          *reinterpret_cast<uint4*>(kb + sw128_offset(row, k & 63)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(kb + sw128_offset(row, (k & 63) + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
    if (part == 2) {   // bias indicator columns 96 (valid token) / 97 (zero-padded token): k-block 1, k = 32 / 33
      const uint32_t one_lo = pack_op(1.0f, 0.0f), one_hi = pack_op(0.0f, 1.0f);
      *reinterpret_cast<uint4*>(a_s + A_KBLOCK_BYTES + sw128_offset(row, 32)) = make_uint4(tokr >= 0 ? one_lo : one_hi, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(a_s + A_KBLOCK_BYTES + sw128_offset(row, 40)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    __syncthreads();
    mark(2);

    // ---------------- qkv = LN1(x) Wqkv^T: TMEM columns [0, 160) and [160, 304) ----------------
    if (warp == 0) {
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 4; ++t, ++n_use) {
        const int wslot = n_use & 3;
        mbar_wait_spin(&bars->full[wslot], (n_use >> 2) & 1u);
        mark(12 + t);      // profile build: time until weight tile t has landed (+ issue of the previous tile's MMAs)
        tc_fence_after();
        if (elect_one()) {
          const int ch = t >> 1, kb = t & 1;
          const uint64_t bd0 = umma_desc_sw128(smem_u32(ring + wslot * FS_STAGE));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (kb * 4 + ks < FS_KQ / 16)
              umma_bf16(tmem_base + (uint32_t)(ch * FS_NQ0), a_desc + (uint64_t)(kb * (A_KBLOCK_BYTES >> 4) + ks * 2), bd0 + (uint64_t)(ks * 2),
                        ch == 0 ? idesc_q0 : idesc_q1, (kb | ks) != 0 ? 1u : 0u);
          }
          if (t == 3) umma_commit(&bars->mma);
        }
        __syncwarp();
      }
      mbar_wait_spin(&bars->mma, ph_mma);
    }
    ph_mma ^= 1u;
    push(has_next ? 4 : 2);     // proj weights of this tile (+ the first half of the next tile's qkv weights)
    __syncthreads();
    mark(3);
    tc_fence_after();
    {
      op_t* qrow = reinterpret_cast<op_t*>(u_s) + row * RS;
      float v[16];
#pragma unroll
      for (int ci = 0; ci < (NQ / 16 + NP - 1) / NP; ++ci) {
        const int cu = part + ci * NP;
        if (cu >= NQ / 16) break;
        tmem_ld16(lane_addr + (uint32_t)(cu * 16), v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_op(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(qrow + cu * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(qrow + cu * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
    __syncthreads();
    mark(4);

    // ---------------- window attention core ----------------
    for (int pr = warp; pr < FB_WIN * NH; pr += NT / 32) {
      const int wl = pr / NH, h = pr - wl * NH;
      if ((long long)tile * FB_WIN + wl >= p.n_windows) continue;
      fb_attention_pair<HD, C, RS, FS_ONES>(u_s, a_s, biasfrag, wl, h, lane);
    }
    fence_proxy_async();
    __syncthreads();
    mark(5);

    // ---------------- proj ----------------
    if (warp == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb, ++n_use) {
        const int wslot = n_use & 3;
        mbar_wait_spin(&bars->full[wslot], (n_use >> 2) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bd0 = umma_desc_sw128(smem_u32(ring + wslot * FS_STAGE));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (kb * 4 + ks < FS_K16 / 16)
              umma_bf16(tmem_base + (uint32_t)FS_TMY, a_desc + (uint64_t)(kb * (A_KBLOCK_BYTES >> 4) + ks * 2), bd0 + (uint64_t)(ks * 2), idesc_c,
                        (kb | ks) != 0 ? 1u : 0u);
          }
          if (kb == 1) umma_commit(&bars->mma);
        }
        __syncwarp();
      }
    }
    mark(8);
    // u_s is free until the next qkv epilogue: gather the next tile's rows into it; fetch this tile's residual (L2)
    if (has_next) issue_loads(slot ^ 1);
    // (transposed ownership for coalescing: in pass ps a lane handles 16-byte chunk lane&3 of row ps*8 + lane/4 of the
    // warp's 32 rows, so one warp instruction touches 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes)
    const int wrow0 = (warp & 3) * 32;
    float4 xres[2][4];
#pragma unroll
    for (int ui = 0; ui < 2; ++ui)
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int cu = part + ui * NP;
        const int tok = tok_s[slot * 128 + wrow0 + ps * 8 + (lane >> 2)];
        xres[ui][ps] = (cu < NU && tok >= 0) ? __ldg(reinterpret_cast<const float4*>(p.x + (long long)tok * C + cu * 16 + (lane & 3) * 4))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }

    mark(9);
    if (warp == 0) {
      mbar_wait_spin(&bars->mma, ph_mma);
      mark(10);
    }
    ph_mma ^= 1u;
    if (has_next) push(2);      // second half of the next tile's qkv weights
    __syncthreads();
    mark(6);
    tc_fence_after();
    {
      // Y + bias goes through a per-warp 2 KB scratch (the A tile is free once the proj MMA has retired) so that the
      // residual add and the global store run in the coalesced transposed ownership; 16-byte chunks are XOR-swizzled
      // with (row >> 1) & 3: conflict free for the row-per-lane write and the 2-rows-per-phase read
      uint8_t* scratch = a_s + warp * 2048;
      float v[16];
#pragma unroll
      for (int ui = 0; ui < 2; ++ui) {
        const int cu = part + ui * NP;       // warp-uniform: the .sync.aligned TMEM load stays converged
        if (cu < NU) {
          tmem_ld16(lane_addr + (uint32_t)(FS_TMY + cu * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = *reinterpret_cast<const float4*>(bproj + cu * 16 + j4 * 4);
            *reinterpret_cast<float4*>(scratch + lane * 64 + ((j4 ^ ((lane >> 1) & 3)) << 4)) =
                make_float4(v[j4 * 4] + bb.x, v[j4 * 4 + 1] + bb.y, v[j4 * 4 + 2] + bb.z, v[j4 * 4 + 3] + bb.w);
          }
          __syncwarp();
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int rl = ps * 8 + (lane >> 2), c = lane & 3;
            const int tok = tok_s[slot * 128 + wrow0 + rl];
            const float4 y = *reinterpret_cast<const float4*>(scratch + rl * 64 + ((c ^ ((rl >> 1) & 3)) << 4));
            const float4 x4 = xres[ui][ps];
            if (tok >= 0)
              *reinterpret_cast<float4*>(p.out + (long long)tok * C + cu * 16 + c * 4) = make_float4(x4.x + y.x, x4.y + y.y, x4.z + y.z, x4.w + y.w);
          }
          __syncwarp();
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    mark(7);
  }
  if constexpr (PROF) if (prof)
    for (int i = 0; i < 16; ++i) p.phase_cycles[(long long)blockIdx.x * 16 + i] += ph_acc[i];
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FS_TMEM);
  }
}

static int launch_swin_attn_stream(FusedBlockParams p, int num_sms, cudaStream_t stream) {
  SWN_CHECK(p.C == FS_C && (p.nH == 3 || p.nH == 6) && !p.do_mlp, "swin_attn_stream: C=96, 3 or 6 heads, attention half only");
  p.nWy = (p.H + 4) / 5;
  p.nWx = (p.W + 4) / 5;
  p.n_windows = (long long)p.B * p.nWy * p.nWx;
  p.ntiles = (int)((p.n_windows + FB_WIN - 1) / FB_WIN);
  const size_t smem = 1024 + 2 * A_KBLOCK_BYTES + FS_U_BYTES + FS_NSTG * FS_STAGE + (size_t)(FS_K16 + p.nH * 1024) * 4 +
                      2 * 128 * 4 + 4 * 128 * 8 + sizeof(FsBars) + 64;
  SWN_CHECK(smem <= 232448, "swin_attn_stream: shared memory overflow (%zu)", smem);
  const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 512, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  if (p.phase_cycles) return p.nH == 3 ? go(swin_attn_stream_kernel<3, true>) : go(swin_attn_stream_kernel<6, true>);
  return p.nH == 3 ? go(swin_attn_stream_kernel<3, false>) : go(swin_attn_stream_kernel<6, false>);
}

int launch_swin_fused(FusedBlockParams p, int num_sms, cudaStream_t stream) {
  const int C = p.C;
  SWN_CHECK((long long)p.B * p.H * p.W < (1ll << 31), "swin_fused: token count overflows int32");
  if (C == FS_C) return launch_swin_attn_stream(p, num_sms, stream);
  SWN_CHECK(p.B > 0 && p.H > 0 && p.W > 0 && C >= 4 && C % 4 == 0 && C <= 48 && p.nH > 0 && C % p.nH == 0,
            "swin_fused: unsupported C=%d nH=%d", C, p.nH);
  const int hd = C / p.nH;
  (void)hd;
  SWN_CHECK((long long)p.B * p.H * p.W < (1ll << 31), "swin_fused: token count overflows int32");
  auto up = [](int v, int a) { return (v + a - 1) / a * a; };
  // geometry shared with packing.py::fused_block_geometry
  p.K16 = up(C, 16);
  p.ones_col = up(3 * C, 8);                 // 8 columns of ones behind q|k|v (softmax denominators via the P V mma)
  p.NQ = up(p.ones_col + 8, 16);
  SWN_CHECK(p.NQ <= 256, "swin_fused: NQ too large");
  p.HC = (4 * C) % 64 == 0 ? 64 : 48;
  SWN_CHECK((4 * C) % p.HC == 0, "swin_fused: hidden width %d not divisible into chunks", 4 * C);
  p.nj = (4 * C) / p.HC;
  p.n_hs = C == 24 ? 1 : p.nj;
  p.RS = p.NQ + 8;
  if (((p.RS / 8) & 1) == 0) p.RS += 8;       // odd number of 16-byte chunks per row: conflict-free row-per-thread access
  const int chunks = C / 4;
  p.RSB = (chunks + ((chunks & 1) ? 0 : 1)) * 16;
  p.nWy = (p.H + 4) / 5;
  p.nWx = (p.W + 4) / 5;
  p.n_windows = (long long)p.B * p.nWy * p.nWx;
  p.ntiles = (int)((p.n_windows + FB_WIN - 1) / FB_WIN);
  const int KB = (p.K16 + 63) >> 6;
  p.w_bytes = (KB * p.NQ + KB * p.K16 + p.nj * KB * p.HC + p.nj * p.K16) * 128;
  p.nf = 2 * p.NQ + 2 * p.K16 + 4 * C + p.nH * 1024;
  const int qkv_bytes = 132 * p.RS * 2, hs_bytes = p.n_hs * A_KBLOCK_BYTES;
  p.u_bytes = up(qkv_bytes > hs_bytes ? qkv_bytes : hs_bytes, 1024);
  const int HCp = (p.HC + 31) & ~31;
  const int hacc_cols = p.nj * HCp;
  const int reg0 = up(p.NQ > hacc_cols ? p.NQ : hacc_cols, 32);
  p.tm_y = reg0;
  int tc = 32;
  while (tc < reg0 + p.K16) tc <<= 1;
  p.tmem_cols = tc;
  const int threads = C >= 48 ? 512 : 256;
  const int want = C >= 48 ? 1 : (C >= 24 ? 2 : 3);     // co-resident CTAs per SM hide the serial phase chain
  size_t smem = 0;
  auto layout = [&](int n_stage) {
    p.n_stage = n_stage;
    p.off_a = up(p.w_bytes, 1024);
    p.off_u = p.off_a + KB * A_KBLOCK_BYTES;
    p.off_stage = p.off_u + p.u_bytes;
    p.off_f = up(p.off_stage + n_stage * 128 * p.RSB, 16);
    p.off_misc = up(p.off_f + p.nf * 4, 16);
    smem = (size_t)p.off_misc + 3 * 128 * 4 + (threads / 128) * 128 * 8 + sizeof(FbBars) + 1024;
  };
  layout(2);
  if ((smem + 1024) * want > 233472) layout(1);   // give up the prefetch buffer before giving up a co-resident CTA
  SWN_CHECK(tc <= 512 && smem <= 232448, "swin_fused: C=%d does not fit (smem %zu, tmem %d)", C, smem, tc);
  int occ = want;
  while (occ > 1 && ((smem + 1024) * occ > 233472 || occ * p.tmem_cols > 512)) --occ;
  long long grid = (long long)num_sms * occ;
  if (grid > p.ntiles) grid = p.ntiles;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  if (C == 12 && p.nH == 3) return p.phase_cycles ? go(swin_fused_kernel<12, 3, 256, 3, true>) : go(swin_fused_kernel<12, 3, 256, 3, false>);
  if (C == 24 && p.nH == 3) return p.phase_cycles ? go(swin_fused_kernel<24, 3, 256, 2, true>) : go(swin_fused_kernel<24, 3, 256, 2, false>);
  if (C == 48 && p.nH == 3) return p.phase_cycles ? go(swin_fused_kernel<48, 3, 512, 1, true>) : go(swin_fused_kernel<48, 3, 512, 1, false>);
  if (C == 48 && p.nH == 6) return p.phase_cycles ? go(swin_fused_kernel<48, 6, 512, 1, true>) : go(swin_fused_kernel<48, 6, 512, 1, false>);
  SWN_CHECK(false, "swin_fused: no kernel instance for C=%d nH=%d (built: 12/3, 24/3, 48/3, 48/6)", C, p.nH);
  return 1;
}

}  // namespace swn
