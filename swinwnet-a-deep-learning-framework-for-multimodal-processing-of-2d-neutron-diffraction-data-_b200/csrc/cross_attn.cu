// Global cross-attention core between the two branches (CrossAttentionBlock / nn.MultiheadAttention,
// SwinWNet.py:764-783): O = softmax(Q K^T / sqrt(hd)) V per (batch, head), flash-style (online softmax,
// scores never leave the SM).  Q/K/V are the bf16 projections produced by the rowgemm kernel (norm_q /
// norm_kv + in_proj fused there); the out_proj + gamma residual is fused into the following rowgemm.
// Tensor-core path: warp-level mma.sync m16n8k16 bf16 with fp32 accumulation (this op is ~3 % of the
// model FLOPs; the projections around it run on tcgen05).
#include "common.cuh"
#include "kernels.h"

namespace swn {

// 8 warps x 16 query rows per CTA: every CTA streams the whole K/V of its (batch, head) from L2, so the query tile
// height sets the L2 traffic (BQ = 64 moved 2.8 GB per call at L = 1920 and was bound by it); K/V tiles are
// double-buffered with cp.async so the next tile lands while the current one is in the tensor cores.
constexpr int CA_THREADS = 256;
constexpr int CA_BQ = 128, CA_BK = 64;

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." SWN_MMA_T "." SWN_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int HD>
__global__ void __launch_bounds__(CA_THREADS) cross_attn_kernel(const CrossAttnParams p) {
  constexpr int LDS = HD + 8;  // padded smem row (bf16 elements)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  op_t* q_s = reinterpret_cast<op_t*>(smem_raw);
  op_t* kv_s = q_s + CA_BQ * LDS;   // [2 stages][K tile | V tile][CA_BK][LDS]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * CA_BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int C = p.C;
  const op_t* qg = p.q + ((long long)b * p.Lq) * C + h * HD;
  const op_t* kg = p.kv + ((long long)b * p.Lk) * (2 * C) + h * HD;

  constexpr int VPR = HD / 8;  // 16-byte vectors per row
  for (int i = threadIdx.x; i < CA_BQ * VPR; i += CA_THREADS) {
    const int r = i / VPR, c = (i % VPR) * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < p.Lq) val = *reinterpret_cast<const uint4*>(qg + (long long)(q0 + r) * C + c);
    *reinterpret_cast<uint4*>(q_s + r * LDS + c) = val;
  }
  __syncthreads();
  // Q fragments stay in registers for the whole KV sweep
  uint32_t qf[HD / 16][4];
  {
    const op_t* qr = q_s + (warp * 16 + g) * LDS + t4 * 2;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      qf[ks][0] = *reinterpret_cast<const uint32_t*>(qr + ks * 16);
      qf[ks][1] = *reinterpret_cast<const uint32_t*>(qr + 8 * LDS + ks * 16);
      qf[ks][2] = *reinterpret_cast<const uint32_t*>(qr + ks * 16 + 8);
      qf[ks][3] = *reinterpret_cast<const uint32_t*>(qr + 8 * LDS + ks * 16 + 8);
    }
  }
  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run[2] = {-1e30f, -1e30f}, l_run[2] = {0.f, 0.f};
  const float sl2 = rsqrtf((float)HD) * 1.4426950408889634f;

  auto issue_kv = [&](int k0, int stage) {
    op_t* k_dst = kv_s + stage * 2 * CA_BK * LDS;
    op_t* v_dst = k_dst + CA_BK * LDS;
    for (int i = threadIdx.x; i < CA_BK * VPR; i += CA_THREADS) {
      const int r = i / VPR, c = (i % VPR) * 8;
      const bool ok = k0 + r < p.Lk;
      const op_t* ksrc = kg + (long long)(ok ? k0 + r : 0) * (2 * C) + c;
      const uint32_t n = ok ? 16u : 0u;     // zero-fill the key tail
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(k_dst + r * LDS + c)), "l"(ksrc), "r"(n) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(v_dst + r * LDS + c)), "l"(ksrc + C), "r"(n) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_kv(0, 0);
  int stage = 0;
  for (int k0 = 0; k0 < p.Lk; k0 += CA_BK, stage ^= 1) {
    if (k0 + CA_BK < p.Lk) {
      issue_kv(k0 + CA_BK, stage ^ 1);     // the other stage was released by the barrier at the end of the last iteration
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const op_t* k_s = kv_s + stage * 2 * CA_BK * LDS;
    const op_t* v_s = k_s + CA_BK * LDS;

    float s[CA_BK / 8][4];
#pragma unroll
    for (int n = 0; n < CA_BK / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      const op_t* kr = k_s + (n * 8 + g) * LDS + t4 * 2;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
        mma_bf16_16816(s[n], qf[ks], b0, b1);
      }
    }
    // scale, mask the key tail, online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {-1e30f, -1e30f};
#pragma unroll
    for (int n = 0; n < CA_BK / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + n * 8 + t4 * 2 + (e & 1);
        const float val = key < p.Lk ? s[n][e] * sl2 : -1e30f;
        s[n][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = ex2_approx(m_run[r] - m_new);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < CA_BK / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = ex2_approx(s[n][e] - m_run[e >> 1]);
        s[n][e] = pv;
        rs[e >> 1] += pv;
      }
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      o[n][0] *= corr[0];
      o[n][1] *= corr[0];
      o[n][2] *= corr[1];
      o[n][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int ks = 0; ks < CA_BK / 16; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_op(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_op(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_op(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_op(s[2 * ks + 1][2], s[2 * ks + 1][3]);
      const uint32_t vrow = smem_u32(v_s + (ks * 16 + (lane & 15)) * LDS);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, vrow + n * 16);
        mma_bf16_16816(o[n], pa, b0, b1);
      }
    }
    __syncthreads();   // this stage may be overwritten by the prefetch issued in the next iteration
  }
  // finalize: quad-reduce the row sums, normalise, store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int qa = q0 + warp * 16 + g, qb = qa + 8;
  op_t* og = p.out + ((long long)b * p.Lq) * C + h * HD + t4 * 2;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    if (qa < p.Lq) *reinterpret_cast<uint32_t*>(og + (long long)qa * C + n * 8) = pack_op(o[n][0] * inv0, o[n][1] * inv0);
    if (qb < p.Lq) *reinterpret_cast<uint32_t*>(og + (long long)qb * C + n * 8) = pack_op(o[n][2] * inv1, o[n][3] * inv1);
  }
}

int launch_cross_attn(CrossAttnParams p, cudaStream_t stream) {
  SWN_CHECK(p.C % p.nH == 0, "cross_attn: C %% nH != 0");
  const int hd = p.C / p.nH;
  SWN_CHECK(hd == 64 || hd == 128, "cross_attn: unsupported head_dim %d (embed_dim must be 48)", hd);
  SWN_CHECK(p.Lq > 0 && p.Lk > 0 && p.B > 0 && p.B <= 65535, "cross_attn: bad sizes");
  dim3 grid((p.Lq + CA_BQ - 1) / CA_BQ, p.nH, p.B);
  const size_t smem = (size_t)(CA_BQ + 4 * CA_BK) * (hd + 8) * 2;
  auto go = [&](auto kern) -> int {
    SWN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, CA_THREADS, smem, stream>>>(p);
    SWN_CUDA(cudaGetLastError());
    return 0;
  };
  return hd == 64 ? go(cross_attn_kernel<64>) : go(cross_attn_kernel<128>);
}

}  // namespace swn
