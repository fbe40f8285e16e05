// Window attention core (W-MSA / SW-MSA) for 5x5 windows:  softmax(q k^T * hd^-0.5 + rel_pos_bias [+ shift
// mask]) v, reading the token-ordered qkv tensor produced by the fused norm1+qkv GEMM and writing the
// token-ordered attention output consumed by the proj GEMM.  window_partition / window_reverse, the
// zero padding to a multiple of 5, torch.roll and compute_mask of the reference (SwinWNet.py:86-149,
// 183-206, 246-272) are pure index math here: no window tensor is ever materialised.
//
// A CTA stages the q|k|v rows of a few windows in shared memory (25 tokens padded to 32 rows); each warp
// then takes (window, head) pairs: S = Q K^T and O = P V run on the tensor cores (mma.sync m16n8k16 bf16,
// fp32 accumulate; the 25x25 problem is far below a tcgen05 tile), softmax runs on the accumulator
// fragments in fp32, O overwrites the pair's Q slot in shared memory and the CTA finally writes whole
// token rows, coalesced.
//
// Padding semantics: the reference pads AFTER norm1, so a padded token is an exact zero vector whose
// q/k/v equal the qkv bias; such tokens take part as keys and their outputs are dropped.
// Shift semantics (shift>0, beyond what the reference can execute): standard Swin —
// windows are taken on the grid rolled by -shift, tokens attend only inside their region id.
#include "common.cuh"
#include "kernels.h"

namespace swn {

constexpr int WA_THREADS = 128;
constexpr int WS = 5, WN = 25, WR = 32;  // window side, tokens per window, padded rows per window

__device__ __forceinline__ void wa_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void wa_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int HD>
__global__ void __launch_bounds__(WA_THREADS) window_attn_kernel(const WinAttnParams p, int nWy, int nWx, int wpb,
                                                                  long long n_windows, int RS) {
  constexpr int KS = HD >= 16 ? HD / 16 : 1;   // k-steps of S = Q K^T
  constexpr int NTO = HD >= 8 ? HD / 8 : 1;    // 8-wide output column tiles of O = P V
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = p.C, nH = p.nH;
  op_t* qkv_s = reinterpret_cast<op_t*>(smem_raw);                       // [wpb][32][RS]
  float* tab_s = reinterpret_cast<float*>(smem_raw + (size_t)wpb * WR * RS * 2);          // [81*nH]
  long long* tok_s = reinterpret_cast<long long*>(tab_s + ((81 * nH + 1) & ~1));         // [wpb][32]
  int* rid_s = reinterpret_cast<int*>(tok_s + wpb * WR);                                  // [wpb][32]

  const long long w0 = (long long)blockIdx.x * wpb;
  // relative-position table, transposed to [head][81] and pre-scaled into the log2 domain
  for (int i = threadIdx.x; i < 81 * nH; i += WA_THREADS) tab_s[i] = p.rpb_table[(i % 81) * nH + i / 81] * 1.4426950408889634f;
  for (int i = threadIdx.x; i < wpb * WR; i += WA_THREADS) {
    const long long w = w0 + i / WR;
    const int t = i % WR;
    long long tok = -2;  // -2: padding row of the 32-row tile (zero), -1: padded token (takes the bias)
    int rid = 0;
    if (w < n_windows && t < WN) {
      tok = -1;
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
  // stage q|k|v (3C bf16 per token), rows in window order: one warp per row, lanes along the row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const bool vec8 = (3 * C) % 8 == 0;  // 16-byte vectors when the row length allows
  for (int slot = warp; slot < wpb * WR; slot += WA_THREADS / 32) {
    const long long tok = tok_s[slot];
    op_t* dst = qkv_s + (size_t)slot * RS;
    if (tok >= 0) {
      const op_t* src = p.qkv + tok * (3 * C);
      // cp.async: every row of the CTA is in flight at once, so DRAM latency is paid once per CTA
      if (vec8) {
        for (int c = lane * 8; c < 3 * C; c += 256)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + c)), "l"(src + c) : "memory");
      } else {
        for (int c = lane * 4; c < 3 * C; c += 128)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + c)), "l"(src + c) : "memory");
      }
    } else {
      for (int c = lane * 4; c < 3 * C; c += 128) {
        uint2 val = make_uint2(0u, 0u);
        if (tok == -1) {
          const float4 bv = *reinterpret_cast<const float4*>(p.qkv_bias + c);
          val = make_uint2(pack_op(bv.x, bv.y), pack_op(bv.z, bv.w));
        }
        *reinterpret_cast<uint2*>(dst + c) = val;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const float sl2 = rsqrtf((float)HD) * 1.4426950408889634f;  // scores are kept in log2 units
  // relative-position index = R[i] - R[j] + 40 with R[t] = (t/5)*9 + t%5
  int Rrow[4], Rcol[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = min((q >> 1) * 16 + g + (q & 1) * 8, WN - 1);
    Rrow[q] = (i / WS) * 9 + i % WS + 40;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = min((q >> 1) * 8 + t4 * 2 + (q & 1), WN - 1);
    Rcol[q] = (j / WS) * 9 + j % WS;
  }

  for (int pr = warp; pr < wpb * nH; pr += WA_THREADS / 32) {
    const int wl = pr / nH, h = pr - wl * nH;
    if (w0 + wl >= n_windows) continue;
    op_t* base = qkv_s + (size_t)wl * WR * RS;
    const int qoff = h * HD, koff = C + h * HD, voff = 2 * C + h * HD;
    // ---- S = Q K^T ----
    float s[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
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
        wa_mma(s[0][nt], a[0], b0, b1);
        wa_mma(s[1][nt], a[1], b0, b1);
      }
    }
    // ---- softmax over the 25 keys (fp32, log2 domain) ----
    const int* rid = rid_s + wl * WR;
    const float* tabh = tab_s + h * 81;
    float inv[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float mx[2] = {-1e30f, -1e30f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = nt * 8 + t4 * 2 + (e & 1);
          const int qr = mt * 2 + (e >> 1);
          float val = fmaf(s[mt][nt][e], sl2, tabh[Rrow[qr] - Rcol[nt * 2 + (e & 1)]]);
          if (p.shift > 0) {
            const int i = mt * 16 + g + (e >> 1) * 8;
            if (rid[min(i, WR - 1)] != rid[j]) val -= 100.0f * 1.4426950408889634f;
          }
          if (nt == 3) val = j < WN ? val : -1e30f;   // keys 25..31 are padding (only the last 8-key tile)
          s[mt][nt][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = ex2_approx(s[mt][nt][e] - mx[e >> 1]);
          s[mt][nt][e] = pv;
          sum[e >> 1] += pv;
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
        inv[mt][r] = 1.0f / sum[r];
      }
    }
    // ---- O = P V ----
    const int vcol0 = voff & ~7;          // 16-byte aligned column of the V tile (HD=4: head sits at +0 or +4)
    float o[2][NTO][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < NTO; ++n) o[mt][n][0] = o[mt][n][1] = o[mt][n][2] = o[mt][n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t pa[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        pa[mt][0] = pack_op(s[mt][2 * ks][0], s[mt][2 * ks][1]);
        pa[mt][1] = pack_op(s[mt][2 * ks][2], s[mt][2 * ks][3]);
        pa[mt][2] = pack_op(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]);
        pa[mt][3] = pack_op(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3]);
      }
      const uint32_t vrow = smem_u32(base + (ks * 16 + (lane & 15)) * RS + vcol0);
#pragma unroll
      for (int n = 0; n < NTO; ++n) {
        uint32_t b0, b1;
        wa_ldsm_x2_trans(b0, b1, vrow + n * 16);
        wa_mma(o[0][n], pa[0], b0, b1);
        wa_mma(o[1][n], pa[1], b0, b1);
      }
    }
    // ---- O (normalised, bf16) overwrites this pair's Q slot ----
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < NTO; ++n) {
        const int dcol = n * 8 + t4 * 2 - (voff - vcol0);   // column inside the head
        if (dcol >= 0 && dcol < HD) {
          op_t* orow = base + (mt * 16 + g) * RS + qoff + dcol;
          *reinterpret_cast<uint32_t*>(orow) = pack_op(o[mt][n][0] * inv[mt][0], o[mt][n][1] * inv[mt][0]);
          *reinterpret_cast<uint32_t*>(orow + 8 * RS) = pack_op(o[mt][n][2] * inv[mt][1], o[mt][n][3] * inv[mt][1]);
        }
      }
  }
  __syncthreads();
  // ---- write whole token rows (C bf16), coalesced ----
  for (int i = warp; i < wpb * WN; i += WA_THREADS / 32) {
    const int wl = i / WN, t = i - wl * WN;
    const long long tok = tok_s[wl * WR + t];
    if (tok < 0) continue;
    const op_t* src = qkv_s + (size_t)(wl * WR + t) * RS;
    op_t* dst = p.out + tok * C;
    if (C % 8 == 0) {
      for (int c = lane * 8; c < C; c += 256) *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
    } else {
      for (int c = lane * 4; c < C; c += 128) *reinterpret_cast<uint2*>(dst + c) = *reinterpret_cast<const uint2*>(src + c);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// shift == 0, head_dim 16 / 32 (every W-MSA of the shipped models at C >= 192): ONE WARP PER WINDOW, no shared-memory
// staging and no block barrier.  The warp walks over the heads of its window; q, k and v of a head are loaded from global
// memory straight into mma.sync fragments with ONE 16-byte (head_dim 32) or 8-byte (16) load per row and operand — possible
// because the contraction index of q k^T and the column index of P v may be permuted (see ldv below; 4-byte loads in the
// canonical fragment order were measured slower than the staged kernel: 8 sectors per instruction saturate the L1 tag
// stage) — v is turned into the B operand of P v with movmatrix.trans, the row sums come from P x ones, and the
// relative-position bias image of every head sits in shared memory in accumulator-fragment order (built once per
// persistent CTA from the [81, nH] table).  The staged kernel above paid three CTA barriers, the table transpose and the
// token index math per group of 1-2 windows and reached 35 % of the HBM roofline.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WW_THREADS = 256;

__device__ __forceinline__ uint32_t ww_movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void ww_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
__global__ void __launch_bounds__(WW_THREADS, 2) window_attn_warp_kernel(const WinAttnParams p, int nWy, int nWx, int n_units, int hsplit) {
  constexpr int KS = HD / 16, NR = HD / 8, VEC = HD / 4;   // k-steps of q k^T, registers (column pairs) / columns per lane and row
  extern __shared__ __align__(16) uint8_t ww_smem[];
  float* bias_f = reinterpret_cast<float*>(ww_smem);           // [nH][2 mt][4 nt][32 lanes][4]
  const int C = p.C, nH = p.nH, C3 = 3 * p.C;
  if (p.bias_frags) {   // ready-made images (built once per block on the host side of the C ABI): 4 KB per head, vector copy
    for (int i = threadIdx.x; i < nH * 256; i += WW_THREADS)
      reinterpret_cast<float4*>(bias_f)[i] = __ldg(reinterpret_cast<const float4*>(p.bias_frags) + i);
  } else {              // from the [81, nH] table: ~40 instructions per entry, 22 us per CTA at 24 heads — a third of the
                        // whole launch at 16 x 30 tokens (measured), which is why the images can be passed in
    for (int i = threadIdx.x; i < nH * 1024; i += WW_THREADS) {
      const int e = i & 3, ln = (i >> 2) & 31, tile = (i >> 7) & 7, h = i >> 10;
      const int row = (tile >> 2) * 16 + (ln >> 2) + (e >> 1) * 8, key = (tile & 3) * 8 + (ln & 3) * 2 + (e & 1);
      float v = 0.f;
      if (key >= WN) v = -1e30f;                                  // key columns 25..31: padding of the mma tile
      else if (row < WN) v = __ldg(p.rpb_table + ((row / WS - key / WS + WS - 1) * (2 * WS - 1) + (row % WS - key % WS + WS - 1)) * nH + h) * 1.4426950408889634f;
      bias_f[i] = v;
    }
  }
  __syncthreads();
  const float4* bias4 = reinterpret_cast<const float4*>(bias_f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const float sl2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const uint32_t ONE_ONE = pack_op(1.f, 1.f);
  const int nWin2 = nWy * nWx;
  // work unit = (window, group of nH / hsplit heads): small inputs are split by heads so that every warp slot has work
  const int hpg = nH / hsplit;
  for (int unit = blockIdx.x * (WW_THREADS / 32) + warp; unit < n_units; unit += gridDim.x * (WW_THREADS / 32)) {
    const int win = unit / hsplit, h_beg = (unit - win * hsplit) * hpg;
    const int b = win / nWin2, wr = win - b * nWin2;
    const int wy = wr / nWx, wx = wr - wy * nWx;
    // row slot s = 0..3 of this lane: token i = 8 s + g of the window.  kind 0: real token, 1: zero-padded token (q/k/v =
    // the qkv bias, SwinWNet.py:254), 2: padding of the second mma tile (zeros; masked as a key by the bias image)
    const op_t* rowp[4];
    long long otok[4];
    int kind[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int i = 8 * s + g;
      kind[s] = 2;
      rowp[s] = p.qkv;
      otok[s] = 0;
      if (i < WN) {
        const int Y = wy * WS + i / WS, X = wx * WS + i % WS;
        kind[s] = 1;
        if (Y < p.H && X < p.W) {
          kind[s] = 0;
          otok[s] = ((long long)b * p.H + Y) * p.W + X;
          rowp[s] = p.qkv + otok[s] * C3;
        }
      }
    }
    // NR consecutive column pairs (8 or 4 columns: one 16- or 8-byte load) of row slot s.  The k index of q k^T and the
    // column index of P v may be permuted freely as long as both operands agree, so lane (g, t) takes columns
    // [VEC t, VEC t + VEC) of its rows: register r is the pair VEC t + 2r, the mma k slots (2t, 2t+1) and (2t+8, 2t+9) of
    // k-step ks are registers 2ks and 2ks+1 — for q and k alike
    auto ldv = [&](int s, int col, uint32_t (&r)[NR]) {
      if (kind[s] == 0) {
        if constexpr (NR == 4) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(rowp[s] + col));
          r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(rowp[s] + col));
          r[0] = v.x; r[1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NR; ++i)
          r[i] = kind[s] == 1 ? pack_op(__ldg(p.qkv_bias + col + 2 * i), __ldg(p.qkv_bias + col + 2 * i + 1)) : 0u;
      }
    };
    for (int h = h_beg; h < h_beg + hpg; ++h) {
      const int qc = h * HD + VEC * t4, kc = C + qc, vc = 2 * C + qc;
      // ---- q, k, v rows straight from global memory (one vector load per row slot and operand) ----
      uint32_t qr[4][NR], kr[4][NR], vr[4][NR];
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        ldv(s4, qc, qr[s4]);
        ldv(s4, kc, kr[s4]);
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) ldv(s4, vc, vr[s4]);
      uint32_t qa[2][KS][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          qa[mt][ks][0] = qr[2 * mt][2 * ks];
          qa[mt][ks][1] = qr[2 * mt + 1][2 * ks];
          qa[mt][ks][2] = qr[2 * mt][2 * ks + 1];
          qa[mt][ks][3] = qr[2 * mt + 1][2 * ks + 1];
        }
      // ---- S = q k^T * hd^-0.5 log2(e) + bias image ----
      float sc[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          sc[mt][nt][0] = sc[mt][nt][1] = sc[mt][nt][2] = sc[mt][nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) ww_mma(sc[mt][nt], qa[mt][ks], kr[nt][2 * ks], kr[nt][2 * ks + 1]);
          const float4 b4 = bias4[(h * 8 + mt * 4 + nt) * 32 + lane];
          sc[mt][nt][0] = fmaf(sc[mt][nt][0], sl2, b4.x);
          sc[mt][nt][1] = fmaf(sc[mt][nt][1], sl2, b4.y);
          sc[mt][nt][2] = fmaf(sc[mt][nt][2], sl2, b4.z);
          sc[mt][nt][3] = fmaf(sc[mt][nt][3], sl2, b4.w);
        }
      // ---- P = 2^(s - rowmax) as A fragments (no exponential for the key columns 25, 27, 29, 31: padding in every lane) ----
      uint32_t pa[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float m0 = fmaxf(fmaxf(sc[mt][0][0], sc[mt][0][1]), fmaxf(sc[mt][1][0], sc[mt][1][1]));
        m0 = fmaxf(m0, fmaxf(fmaxf(sc[mt][2][0], sc[mt][2][1]), sc[mt][3][0]));
        float m1 = fmaxf(fmaxf(sc[mt][0][2], sc[mt][0][3]), fmaxf(sc[mt][1][2], sc[mt][1][3]));
        m1 = fmaxf(m1, fmaxf(fmaxf(sc[mt][2][2], sc[mt][2][3]), sc[mt][3][2]));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        pa[mt][0][0] = pack_op(ex2_approx(sc[mt][0][0] - m0), ex2_approx(sc[mt][0][1] - m0));
        pa[mt][0][1] = pack_op(ex2_approx(sc[mt][0][2] - m1), ex2_approx(sc[mt][0][3] - m1));
        pa[mt][0][2] = pack_op(ex2_approx(sc[mt][1][0] - m0), ex2_approx(sc[mt][1][1] - m0));
        pa[mt][0][3] = pack_op(ex2_approx(sc[mt][1][2] - m1), ex2_approx(sc[mt][1][3] - m1));
        pa[mt][1][0] = pack_op(ex2_approx(sc[mt][2][0] - m0), ex2_approx(sc[mt][2][1] - m0));
        pa[mt][1][1] = pack_op(ex2_approx(sc[mt][2][2] - m1), ex2_approx(sc[mt][2][3] - m1));
        pa[mt][1][2] = pack_op(ex2_approx(sc[mt][3][0] - m0), 0.f);
        pa[mt][1][3] = pack_op(ex2_approx(sc[mt][3][2] - m1), 0.f);
      }
      // ---- O = P v, row sums = P x ones ----
      float od[2][4] = {};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        ww_mma(od[0], pa[0][ks], ONE_ONE, ONE_ONE);
        ww_mma(od[1], pa[1][ks], ONE_ONE, ONE_ONE);
      }
      float inv[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        inv[mt][0] = rcp_approx(od[mt][0]);
        inv[mt][1] = rcp_approx(od[mt][2]);
      }
      // v register r of key tile j is an 8x8 matrix [key][column VEC (c / 2) + 2r + c % 2]; transposed it is the B operand of
      // output tile r, whose accumulator columns (2t, 2t+1) are the columns VEC t + 2r, + 1: the NR tiles of a lane are
      // VEC consecutive columns of its rows -> one vector store per row
      uint32_t orow[4][NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float o[2][4] = {};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t b0 = ww_movm_trans(vr[2 * ks][r]), b1 = ww_movm_trans(vr[2 * ks + 1][r]);
          ww_mma(o[0], pa[0][ks], b0, b1);
          ww_mma(o[1], pa[1][ks], b0, b1);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          orow[2 * mt][r] = pack_op(o[mt][0] * inv[mt][0], o[mt][1] * inv[mt][0]);
          orow[2 * mt + 1][r] = pack_op(o[mt][2] * inv[mt][1], o[mt][3] * inv[mt][1]);
        }
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4)
        if (kind[s4] == 0) {
          op_t* dst = p.out + otok[s4] * C + qc;
          if constexpr (NR == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(orow[s4][0], orow[s4][1], orow[s4][2], orow[s4][3]);
          else *reinterpret_cast<uint2*>(dst) = make_uint2(orow[s4][0], orow[s4][1]);
        }
    }
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
#ifndef SWN_WA_WARP
#define SWN_WA_WARP 1
#endif
  if (SWN_WA_WARP && p.shift == 0 && (hd == 16 || hd == 32) && (size_t)p.nH * 4096 <= 100 * 1024 && n_windows * p.nH < (1ll << 31)) {
    const size_t smem_w = (size_t)p.nH * 4096;
    auto go_w = [&](auto kern) -> int {
      SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
      int dev = 0, sms = 148, occ = 0;
      SWN_CUDA(cudaGetDevice(&dev));
      SWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      SWN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WW_THREADS, smem_w));
      SWN_CHECK(occ > 0, "window_attn: warp kernel does not fit on an SM");
      long long grid = (long long)sms * occ;
      int hsplit = 1;                      // split the heads of a window over several warps until every warp slot has ~3 units
      while (hsplit < p.nH && n_windows * hsplit < 3 * grid * (WW_THREADS / 32)) {
        ++hsplit;
        while (p.nH % hsplit) ++hsplit;
      }
      const long long n_units = n_windows * hsplit;
      const long long need = (n_units + WW_THREADS / 32 - 1) / (WW_THREADS / 32);
      if (grid > need) grid = need;
      kern<<<(unsigned)grid, WW_THREADS, smem_w, stream>>>(p, nWy, nWx, (int)n_units, hsplit);
      SWN_CUDA(cudaGetLastError());
      return 0;
    };
    return hd == 16 ? go_w(window_attn_warp_kernel<16>) : go_w(window_attn_warp_kernel<32>);
  }
  // padded row stride (bf16 elements): multiple of 8 (16-byte rows for ldmatrix), (RS/2) % 8 == 4 when possible
  int RS = ((3 * p.C + 7) & ~7) + 8;
  if (((RS / 2) & 7) == 0) RS += 8;
  // windows per CTA: enough (window, head) pairs for the 4 warps, but at most ~44 KB of staging so that >= 4 CTAs
  // (16+ warps) are resident per SM
  int wpb = 12 / p.nH;
  if (wpb < 1) wpb = 1;
  if (wpb > 8) wpb = 8;
  while (wpb > 1 && wpb * WR * RS * 2 > 44 * 1024) --wpb;
  const size_t smem = (size_t)wpb * WR * RS * 2 + (size_t)((81 * p.nH + 1) & ~1) * 4 + (size_t)wpb * WR * 12 + 16;
  const long long grid = (n_windows + wpb - 1) / wpb;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, WA_THREADS, smem, stream>>>(p, nWy, nWx, wpb, n_windows, RS);
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
