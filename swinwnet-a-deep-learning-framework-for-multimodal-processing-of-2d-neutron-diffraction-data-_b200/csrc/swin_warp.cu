// Whole SwinTransformerBlocks (shift 0) for the narrow layers — C = 12 at 500x960 and C = 24 at 250x480 tokens per
// diffraction (UpscalingHead, SwinWNet.py:656-678), C = 48 at 125x240 (first encoder stage), 3 (or 6) heads,
// SwinWNet.py:236-280 — with ONE WARP PER WINDOW and no block-level synchronisation at all:
//
//   x rows (25 tokens, padded to two 16-row mma tiles) -> LN1 -> q | k | v^T -> per head S = q k^T + bias, softmax, P v
//     -> proj + x -> LN2 -> fc1 -> GELU -> fc2 + -> out
//
// Every intermediate lives in the registers of the warp as mma.sync (m16n8k16 / m16n8k8) fragments: the accumulator
// fragment of one GEMM is, after packing to 16 bit, the A fragment of the next (rows stay with their lanes), the k
// accumulators are the B fragments of q k^T, and v is produced TRANSPOSED (v^T = Wv xn^T, the LayerNorm fragments
// serving as the B operand) so that its accumulators are the B fragments of P v.  Shared memory holds only the weights
// (as ready-made fragments, 5 / 17 KB) and the relative-position bias image; HBM traffic is one fp32 read and one fp32
// write of the row.  At these widths a 128-row tcgen05 tile is > 75 % padding and the tcgen05 block kernel
// (swin_fused.cu) spent its time in the ~10 CTA-wide barriers / MMA round trips per tile (39 % issue utilisation,
// profiles/r2_ncu_lines_n_f12.txt); here the only waits are the warp's own instruction latencies, hidden by the other
// 11-15 warps of the SM.  Biases ride on two spare k columns of the padded operands (column C is 1 for every row, column
// C+1 is 1 for real tokens only: LayerNorm's beta contribution must vanish for the zero-padded window tokens,
// SwinWNet.py:242,254), LayerNorm's gamma is folded into the weights, head_dim^-0.5 log2(e) into q.
// C = 48 has no spare k column (48 = 3 x 16): there the biases initialise the accumulators instead (LayerNorm's beta folded
// into them for real tokens, the plain bias for zero-padded ones), and the weight fragments (70 KB per block) are shared by
// the 8 warps of one 256-thread CTA per SM.
#include "common.cuh"
#include "kernels.h"

namespace swn {

namespace {

#ifndef SWN_WB_MINB12
#define SWN_WB_MINB12 4      // C = 12: CTAs of 4 warps per SM: 128 registers per thread
#endif
#ifndef SWN_WB_MINB24
#define SWN_WB_MINB24 3      // C = 24: 168 registers per thread
#endif

__device__ __forceinline__ void wb_mma16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void wb_mma8(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
#ifndef SWN_WB_THR48
#define SWN_WB_THR48 256     // C = 48: one CTA per SM (70 KB of weight fragments per block); 256 threads = 255 registers per thread
#endif
#ifndef SWN_WB_EXP16
#define SWN_WB_EXP16 0       // 1: softmax exponentials as ex2.approx.f16x2 of the max-subtracted pair (measured slower: 2.49 -> 2.60 ms)
#endif
// 2^a, 2^b (a, b <= 0: row maximum already subtracted in fp32) as a packed 16-bit pair — the A fragment of P v
__device__ __forceinline__ uint32_t wb_exp2_pair(float a, float b) {
#if SWN_WB_EXP16 && !SWN_OPERAND_BF16
  uint32_t r;
  const uint32_t x = pack_op(a, b);    // saturating: the -1e30 of masked key columns becomes -65504 -> 2^x = 0
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
#else
  return pack_op(ex2_approx(a), ex2_approx(b));
#endif
}
__device__ __forceinline__ uint32_t wb_exp2_single(float a) {   // (2^a, 0): the odd key columns 25..31 are padding
#if SWN_WB_EXP16 && !SWN_OPERAND_BF16
  return wb_exp2_pair(a, -65504.f);
#else
  return pack_op(ex2_approx(a), 0.f);
#endif
}
__device__ __forceinline__ float wb_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

}  // namespace

// Geometry for channel width C / NH heads (mirrored by packing.py::warp_block_geometry / pack_warp_block)
template <int C_, int NH_>
struct WbGeom {
  static constexpr int C = C_, NH = NH_, HD = C / NH;
  static constexpr int K16 = (C + 15) / 16 * 16, KT = K16 / 16;
  static constexpr int NJ = (C + 7) / 8;    // 8-column tiles that hold real channels
  static constexpr int NT8 = K16 / 8;       // 8-column tiles of the padded row
  static constexpr int MTV = KT;            // 16-row tiles of v^T
  static constexpr int NH1 = C / 2;         // 8-column tiles of the hidden row (4C / 8)
  static constexpr int KT2 = C / 4;         // k-steps of fc2 (4C / 16)
  static constexpr bool BIASCOL = K16 >= C + 2;              // biases ride on the k columns C, C+1 (else: accumulator init)
  static constexpr int ONE_J = C / 8, ONE_T = (C % 8) / 2;   // tile / lane%4 that hold the bias columns C, C+1
  static constexpr int G8 = HD >= 8 ? HD / 8 : 1;            // 8-column tiles per head group
  static constexpr int HPG = HD >= 8 ? 1 : 8 / HD;           // heads per group (two 4-wide heads share a tile)
  static constexpr int NG = NJ / G8;                         // head groups
  // 16-bit weight fragments, element offsets
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + KT * NJ * 128;
  static constexpr int OFF_V = OFF_K + KT * NJ * 128;
  static constexpr int OFF_P = OFF_V + MTV * KT * 256;
  static constexpr int OFF_1 = OFF_P + KT * NJ * 128;
  static constexpr int OFF_2 = OFF_1 + KT * NH1 * 128;
  static constexpr int W_ELEMS = OFF_2 + KT2 * NJ * 128;
  // fp32 block: b2 [K16] | bias fragments [NH][2][4][32][4] | (no bias columns:) bq' bk' bv' | bq bk bv | bproj | b1'
  static constexpr int F_BIAS = K16;
  static constexpr int F_X = K16 + NH * 1024;
  static constexpr int F_QF = F_X, F_KF = F_X + C, F_VF = F_X + 2 * C, F_QP = F_X + 3 * C, F_KP = F_X + 4 * C, F_VP = F_X + 5 * C;
  static constexpr int F_BP = F_X + 6 * C, F_B1 = F_X + 7 * C;
  static constexpr int F_ELEMS = F_X + (BIASCOL ? 0 : 11 * C);
  static_assert(C % 4 == 0 && C % NH == 0 && (HD == 4 || HD == 8 || HD == 16) && NJ % G8 == 0 && F_ELEMS % 4 == 0 && (BIASCOL || C % 16 == 0),
                "unsupported channel width / head count");
};

template <int C, int NH, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) swin_warp_block_kernel(const WarpBlockParams p) {
  using G = WbGeom<C, NH>;
  constexpr int NJ = G::NJ, NT8 = G::NT8, KT = G::KT, HD = G::HD, G8 = G::G8, HPG = G::HPG;
  constexpr bool BIASCOL = G::BIASCOL;
  extern __shared__ __align__(16) uint8_t wb_smem[];
  uint32_t* w_s = reinterpret_cast<uint32_t*>(wb_smem);
  float* f_s = reinterpret_cast<float*>(wb_smem + p.depth * G::W_ELEMS * 2);
  for (int i = threadIdx.x; i < p.depth * (G::W_ELEMS / 8); i += NTHR)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wpk) + i);
  for (int i = threadIdx.x; i < p.depth * (G::F_ELEMS / 4); i += NTHR)
    reinterpret_cast<float4*>(f_s)[i] = __ldg(reinterpret_cast<const float4*>(p.fpk) + i);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint32_t ONE_ZERO = pack_op(1.f, 0.f), ONE_ONE = pack_op(1.f, 1.f);

  // row slot s = 0..3 of this lane: token i = 8 s + g of the window (i >= 25: padding of the second mma tile)
  int dy[4], dx[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int i = 8 * s + g;
    dy[s] = i < 25 ? i / 5 : 1 << 20;
    dx[s] = i - (i / 5) * 5;
  }
  bool colv[NJ];   // does this lane's column pair of tile j hold real channels?
#pragma unroll
  for (int j = 0; j < NJ; ++j) colv[j] = (C % 8 == 0) || (8 * j + 2 * t < C);
  const float inv_c = 1.0f / C;

  // (x - mean) * rstd of the four row slots as 16-bit A fragments.  first: rows that are not real tokens (zero-padded window
  // tokens, SwinWNet.py:254) are zero AFTER the norm; with bias columns, column C+1 is their indicator
  auto layer_norm = [&](const float (&v)[2][NJ][4], uint32_t (&a)[2][KT][4], const bool (&real)[4], bool first) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int mt = s >> 1, e0 = (s & 1) * 2;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) sum += v[mt][j][e0] + v[mt][j][e0 + 1];
      const float mean = quad_sum(sum) * inv_c;
      float d[NJ][2], q = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        d[j][0] = colv[j] ? v[mt][j][e0] - mean : 0.f;
        d[j][1] = colv[j] ? v[mt][j][e0 + 1] - mean : 0.f;
        q = fmaf(d[j][0], d[j][0], q);
        q = fmaf(d[j][1], d[j][1], q);
      }
      float rstd = wb_rsqrt(quad_sum(q) * inv_c + p.eps);
      if (first && !real[s]) rstd = 0.f;
      uint32_t pk[NT8];
#pragma unroll
      for (int j = 0; j < NT8; ++j) pk[j] = j < NJ ? pack_op(d[j < NJ ? j : 0][0] * rstd, d[j < NJ ? j : 0][1] * rstd) : 0u;
      if (BIASCOL && t == G::ONE_T) pk[G::ONE_J < NT8 ? G::ONE_J : 0] = (first && real[s]) ? ONE_ONE : ONE_ZERO;
#pragma unroll
      for (int j = 0; j < NT8; ++j) a[mt][j >> 1][(j & 1) * 2 + (s & 1)] = pk[j];
    }
  };

  const int nWin2 = p.nWy * p.nWx;
  const int nwarps = gridDim.x * (NTHR / 32);
  for (int win = blockIdx.x * (NTHR / 32) + warp; win < p.n_windows; win += nwarps) {
    const int b = win / nWin2, wr = win - b * nWin2;
    const int wy5 = (wr / p.nWx) * 5, wx5 = (wr - (wr / p.nWx) * p.nWx) * 5;
    // ---- rows of the window: accumulator-fragment layout (row slot, 8-column tile, column pair 2t) ----
    float x[2][NJ][4];
    int tok[4];
    bool real[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int Y = wy5 + dy[s], X = wx5 + dx[s];
      real[s] = Y < p.H && X < p.W;
      tok[s] = real[s] ? (b * p.H + Y) * p.W + X : -1;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float2 r = make_float2(0.f, 0.f);
        if (real[s] && colv[j]) r = __ldg(reinterpret_cast<const float2*>(p.x + (long long)tok[s] * C + 8 * j + 2 * t));
        x[s >> 1][j][(s & 1) * 2] = r.x;
        x[s >> 1][j][(s & 1) * 2 + 1] = r.y;
      }
    }
    // does the window hold zero-padded tokens (H or W not a multiple of 5)?  Only then the two bias vectors differ per row
    const bool padded = !BIASCOL && (wy5 + 5 > p.H || wx5 + 5 > p.W);
    // the blocks of a BasicLayer all use shift 0 (SwinWNet.py:328), i.e. the SAME window partition: the rows of the window
    // run through `depth` consecutive blocks without leaving the registers
    for (int blk = 0; blk < p.depth; ++blk) {
      const uint32_t* wb = w_s + blk * (G::W_ELEMS / 2);
      const float* fb = f_s + blk * G::F_ELEMS;
      const uint2* wq = reinterpret_cast<const uint2*>(wb + G::OFF_Q / 2) + lane;
      const uint2* wk = reinterpret_cast<const uint2*>(wb + G::OFF_K / 2) + lane;
      const uint4* wv = reinterpret_cast<const uint4*>(wb + G::OFF_V / 2) + lane;
      const uint2* wp = reinterpret_cast<const uint2*>(wb + G::OFF_P / 2) + lane;
      const uint2* w1 = reinterpret_cast<const uint2*>(wb + G::OFF_1 / 2) + lane;
      const uint2* w2 = reinterpret_cast<const uint2*>(wb + G::OFF_2 / 2) + lane;
      const float4* biasfrag = reinterpret_cast<const float4*>(fb + G::F_BIAS) + lane;
      uint32_t a1[2][KT][4];
      layer_norm(x, a1, real, true);
      // accumulator start of a q / k column tile (no bias columns): the bias of this lane's column pair, per row slot the one
      // with LayerNorm's beta folded in (real token) or the plain one (zero-padded token)
      auto qk_init = [&](float (&acc)[2][4], int off_fold, int off_plain, int n) {
        const float2 bf = *reinterpret_cast<const float2*>(fb + off_fold + 8 * n + 2 * t);
        float2 bp = bf;
        if (padded) bp = *reinterpret_cast<const float2*>(fb + off_plain + 8 * n + 2 * t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const bool r0 = !padded || real[2 * mt], r1 = !padded || real[2 * mt + 1];
          acc[mt][0] = r0 ? bf.x : bp.x; acc[mt][1] = r0 ? bf.y : bp.y;
          acc[mt][2] = r1 ? bf.x : bp.x; acc[mt][3] = r1 ? bf.y : bp.y;
        }
      };
      // ---- v^T = Wv xn^T for one 16-row tile of v channels: rows = channels, columns = the 32 tokens; packed = B operand of P v
      uint32_t vb[4][2];
      auto v_transposed = [&](int mv) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float vc[4] = {0.f, 0.f, 0.f, 0.f};
          if (!BIASCOL) {
            const float f0 = fb[G::F_VF + 16 * mv + g], f1 = fb[G::F_VF + 16 * mv + g + 8];
            vc[0] = vc[1] = f0;
            vc[2] = vc[3] = f1;
            if (padded) {
              const float p0 = fb[G::F_VP + 16 * mv + g], p1 = fb[G::F_VP + 16 * mv + g + 8];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = 8 * nt + 2 * t + e;        // token of this accumulator column
                const bool rl = i >= 25 || (wy5 + i / 5 < p.H && wx5 + i % 5 < p.W);
                if (!rl) { vc[e] = p0; vc[2 + e] = p1; }
              }
            }
          }
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            const uint4 av = wv[(mv * KT + kt) * 32];
            wb_mma16(vc, &av.x, a1[nt >> 1][kt][nt & 1], a1[nt >> 1][kt][2 + (nt & 1)]);
          }
          vb[nt][0] = pack_op(vc[0], vc[1]);
          vb[nt][1] = pack_op(vc[2], vc[3]);
        }
      };

      // ---- attention, head group by head group; normalised output pairs (row slot, tile) collect in opk ----
      uint32_t opk[4][NT8];
      float keep[2][4];   // two heads per tile (head_dim 4): the first head's normalised output until the second is done
#pragma unroll
      for (int gi = 0; gi < G::NG; ++gi) {
        uint32_t qa[G8][2][2], kb[G8][4];
#pragma unroll
        for (int gt = 0; gt < G8; ++gt) {
          const int n = gi * G8 + gt;
          float qc[2][4] = {}, kc[2][4] = {};
          if (!BIASCOL) {
            qk_init(qc, G::F_QF, G::F_QP, n);
            qk_init(kc, G::F_KF, G::F_KP, n);
          }
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            const uint2 bq = wq[(kt * NJ + n) * 32], bk = wk[(kt * NJ + n) * 32];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              wb_mma16(qc[mt], a1[mt][kt], bq.x, bq.y);
              wb_mma16(kc[mt], a1[mt][kt], bk.x, bk.y);
            }
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            qa[gt][mt][0] = pack_op(qc[mt][0], qc[mt][1]);
            qa[gt][mt][1] = pack_op(qc[mt][2], qc[mt][3]);
            kb[gt][2 * mt] = pack_op(kc[mt][0], kc[mt][1]);
            kb[gt][2 * mt + 1] = pack_op(kc[mt][2], kc[mt][3]);
          }
        }
#pragma unroll
        for (int hh = 0; hh < HPG; ++hh) {
          const int h = gi * HPG + hh;
          if (h >= NH) break;
          if (((h * HD) & 15) == 0) v_transposed((h * HD) >> 4);     // first head of a 16-channel tile of v^T
          // S = q_h k_h^T on top of the relative-position bias image (log2 domain, -1e30 on key columns 25..31)
          float sc[2][4][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const float4 b4 = biasfrag[(h * 8 + mt * 4 + nt) * 32];
              sc[mt][nt][0] = b4.x; sc[mt][nt][1] = b4.y; sc[mt][nt][2] = b4.z; sc[mt][nt][3] = b4.w;
            }
          if constexpr (G8 == 2) {        // head_dim 16: one k16 step over the two tiles of the head
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t af[4] = {qa[0][mt][0], qa[0][mt][1], qa[1][mt][0], qa[1][mt][1]};
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) wb_mma16(sc[mt][nt], af, kb[0][nt], kb[1][nt]);
            }
          } else {
            const bool mine = HPG == 1 || (t >> 1) == hh;    // two heads share a tile: this lane's columns belong to head hh
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t q0 = mine ? qa[0][mt][0] : 0u, q1 = mine ? qa[0][mt][1] : 0u;
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) wb_mma8(sc[mt][nt], q0, q1, kb[0][nt]);
            }
          }
          // P = 2^(s - rowmax) as A fragments of P v; no exponential for key columns 25, 27, 29, 31 (padding in every lane)
          uint32_t pa[2][2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            float m0 = fmaxf(fmaxf(sc[mt][0][0], sc[mt][0][1]), fmaxf(sc[mt][1][0], sc[mt][1][1]));
            m0 = fmaxf(m0, fmaxf(fmaxf(sc[mt][2][0], sc[mt][2][1]), sc[mt][3][0]));
            float m1 = fmaxf(fmaxf(sc[mt][0][2], sc[mt][0][3]), fmaxf(sc[mt][1][2], sc[mt][1][3]));
            m1 = fmaxf(m1, fmaxf(fmaxf(sc[mt][2][2], sc[mt][2][3]), sc[mt][3][2]));
            m0 = quad_max(m0);
            m1 = quad_max(m1);
            pa[mt][0][0] = wb_exp2_pair(sc[mt][0][0] - m0, sc[mt][0][1] - m0);
            pa[mt][0][1] = wb_exp2_pair(sc[mt][0][2] - m1, sc[mt][0][3] - m1);
            pa[mt][0][2] = wb_exp2_pair(sc[mt][1][0] - m0, sc[mt][1][1] - m0);
            pa[mt][0][3] = wb_exp2_pair(sc[mt][1][2] - m1, sc[mt][1][3] - m1);
            pa[mt][1][0] = wb_exp2_pair(sc[mt][2][0] - m0, sc[mt][2][1] - m0);
            pa[mt][1][1] = wb_exp2_pair(sc[mt][2][2] - m1, sc[mt][2][3] - m1);
            pa[mt][1][2] = wb_exp2_single(sc[mt][3][0] - m0);
            pa[mt][1][3] = wb_exp2_single(sc[mt][3][2] - m1);
          }
          // row sums (P times a tile of ones) and O = P v_h, both on the tensor cores
          float od[2][4] = {};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) wb_mma16(od[mt], pa[mt][ks], ONE_ONE, ONE_ONE);
          float inv[2][2];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            inv[mt][0] = rcp_approx(od[mt][0]);
            inv[mt][1] = rcp_approx(od[mt][2]);
          }
#pragma unroll
          for (int ot = 0; ot < G8; ++ot) {
            const int half = ((h * HD + 8 * ot) >> 3) & 1;
            float o[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) wb_mma16(o[mt], pa[mt][ks], vb[2 * ks][half], vb[2 * ks + 1][half]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              o[mt][0] *= inv[mt][0]; o[mt][1] *= inv[mt][0]; o[mt][2] *= inv[mt][1]; o[mt][3] *= inv[mt][1];
            }
            if (HPG == 2 && hh == 0 && h + 1 < NH) {
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int e = 0; e < 4; ++e) keep[mt][e] = o[mt][e];
            } else {
              if (HPG == 2 && hh == 1) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[mt][e] = (t >> 1) == 0 ? keep[mt][e] : o[mt][e];
              }
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                opk[2 * mt][gi * G8 + ot] = pack_op(o[mt][0], o[mt][1]);
                opk[2 * mt + 1][gi * G8 + ot] = pack_op(o[mt][2], o[mt][3]);
              }
            }
          }
        }
      }
      if (BIASCOL) {   // bias column of the proj operand (and zero padding behind it)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
          for (int j = NJ; j < NT8; ++j) opk[s][j] = 0u;
          if (t == G::ONE_T) opk[s][G::ONE_J < NT8 ? G::ONE_J : 0] = ONE_ZERO;
          else if (G::ONE_J < NJ && 8 * G::ONE_J + 2 * t >= C) opk[s][G::ONE_J < NT8 ? G::ONE_J : 0] = 0u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float2 bp = *reinterpret_cast<const float2*>(fb + G::F_BP + 8 * j + 2 * t);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            x[mt][j][0] += bp.x; x[mt][j][1] += bp.y; x[mt][j][2] += bp.x; x[mt][j][3] += bp.y;
          }
        }
      }
      // ---- x1 = x + proj(attn) + b, accumulated in place on the residual ----
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t ao[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ao[mt][0] = opk[2 * mt][2 * kt];
          ao[mt][1] = opk[2 * mt + 1][2 * kt];
          ao[mt][2] = opk[2 * mt][2 * kt + 1];
          ao[mt][3] = opk[2 * mt + 1][2 * kt + 1];
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const uint2 bw = wp[(kt * NJ + j) * 32];
          wb_mma16(x[0][j], ao[0], bw.x, bw.y);
          wb_mma16(x[1][j], ao[1], bw.x, bw.y);
        }
      }
      // ---- LN2, fc1 -> GELU -> fc2 in 16-column steps of the hidden row, accumulated on x1 + b2 ----
      uint32_t a2[2][KT][4];
      layer_norm(x, a2, real, false);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float2 b2 = *reinterpret_cast<const float2*>(fb + 8 * j + 2 * t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          x[mt][j][0] += b2.x; x[mt][j][1] += b2.y; x[mt][j][2] += b2.x; x[mt][j][3] += b2.y;
        }
      }
#pragma unroll
      for (int u = 0; u < G::KT2; ++u) {
        float h0[2][4] = {}, h1[2][4] = {};
        if (!BIASCOL) {
          const float2 c0 = *reinterpret_cast<const float2*>(fb + G::F_B1 + 16 * u + 2 * t);
          const float2 c1 = *reinterpret_cast<const float2*>(fb + G::F_B1 + 16 * u + 8 + 2 * t);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            h0[mt][0] = h0[mt][2] = c0.x; h0[mt][1] = h0[mt][3] = c0.y;
            h1[mt][0] = h1[mt][2] = c1.x; h1[mt][1] = h1[mt][3] = c1.y;
          }
        }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const uint2 b0 = w1[(kt * G::NH1 + 2 * u) * 32], b1 = w1[(kt * G::NH1 + 2 * u + 1) * 32];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            wb_mma16(h0[mt], a2[mt][kt], b0.x, b0.y);
            wb_mma16(h1[mt], a2[mt][kt], b1.x, b1.y);
          }
        }
        uint32_t ah[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ah[mt][0] = gelu_pack2(h0[mt][0], h0[mt][1]);
          ah[mt][1] = gelu_pack2(h0[mt][2], h0[mt][3]);
          ah[mt][2] = gelu_pack2(h1[mt][0], h1[mt][1]);
          ah[mt][3] = gelu_pack2(h1[mt][2], h1[mt][3]);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const uint2 bw = w2[(u * NJ + j) * 32];
          wb_mma16(x[0][j], ah[0], bw.x, bw.y);
          wb_mma16(x[1][j], ah[1], bw.x, bw.y);
        }
      }
    }   // blocks
    // ---- write the rows back ----
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (real[s]) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (colv[j])
            *reinterpret_cast<float2*>(p.out + (long long)tok[s] * C + 8 * j + 2 * t) =
                make_float2(x[s >> 1][j][(s & 1) * 2], x[s >> 1][j][(s & 1) * 2 + 1]);
      }
  }
}

int launch_swin_warp_block(WarpBlockParams p, int num_sms, cudaStream_t stream) {
  SWN_CHECK((p.nH == 3 && (p.C == 12 || p.C == 24 || p.C == 48)) || (p.nH == 6 && p.C == 48),
            "swin_block_warp: no kernel instance for C=%d nH=%d (built: 12/3, 24/3, 48/3, 48/6)", p.C, p.nH);
  SWN_CHECK(p.B > 0 && p.H > 0 && p.W > 0, "swin_block_warp: empty input");
  SWN_CHECK(p.depth >= 1 && p.depth <= 4, "swin_block_warp: depth %d not in 1..4", p.depth);
  p.nWy = (p.H + 4) / 5;
  p.nWx = (p.W + 4) / 5;
  const long long nw = (long long)p.B * p.nWy * p.nWx;
  SWN_CHECK(nw * 25 < (1ll << 31), "swin_block_warp: token count overflows int32");
  p.n_windows = (int)nw;
  auto go = [&](auto kern, int nthr, size_t smem) -> int {
    SWN_CHECK(smem <= 232448, "swin_block_warp: %d blocks of C=%d do not fit in shared memory (%zu bytes)", p.depth, p.C, smem);
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SWN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthr, smem));
    SWN_CHECK(occ > 0, "swin_block_warp: kernel does not fit on an SM");
    long long grid = (long long)num_sms * occ;
    const long long need = (nw + nthr / 32 - 1) / (nthr / 32);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, nthr, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  auto bytes = [&](auto geom) { return (size_t)p.depth * (decltype(geom)::W_ELEMS * 2 + decltype(geom)::F_ELEMS * 4); };
  if (p.C == 12) return go(swin_warp_block_kernel<12, 3, 128, SWN_WB_MINB12>, 128, bytes(WbGeom<12, 3>{}));
  if (p.C == 24) return go(swin_warp_block_kernel<24, 3, 128, SWN_WB_MINB24>, 128, bytes(WbGeom<24, 3>{}));
  if (p.nH == 3) return go(swin_warp_block_kernel<48, 3, SWN_WB_THR48, 1>, SWN_WB_THR48, bytes(WbGeom<48, 3>{}));
  return go(swin_warp_block_kernel<48, 6, SWN_WB_THR48, 1>, SWN_WB_THR48, bytes(WbGeom<48, 6>{}));
}

}  // namespace swn
