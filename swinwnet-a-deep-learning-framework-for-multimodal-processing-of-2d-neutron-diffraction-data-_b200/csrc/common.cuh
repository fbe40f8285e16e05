// Shared device helpers for the sm_100a SwinWNet kernels: mbarrier / bulk-copy (TMA engine) /
// tcgen05 (UMMA + TMEM) PTX wrappers, the K-major SWIZZLE_128B shared-memory tile layout, and
// small math utilities.  Everything here is hand-written for sm_100a; no CUTLASS dependency.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace swn {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define SWN_CHECK(cond, ...)                     \
  do {                                           \
    if (!(cond)) {                               \
      swn::set_error(__VA_ARGS__);               \
      return 1;                                  \
    }                                            \
  } while (0)
#define SWN_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      swn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// generic device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---------------------------------------------------------------------------------------------
// Tensor-core operand type.  All MMA operands (LayerNorm outputs, qkv, attention probabilities / outputs,
// GELU outputs, weights) are 16-bit with fp32 accumulation.  Default is IEEE fp16 (10-bit mantissa): almost every
// operand on this path is bounded (post-LayerNorm, softmax-weighted averages), the rest saturates (pack_op), and the
// 8x smaller rounding error than bf16 is what keeps the end-to-end error well inside the 2e-2 parity budget
// (PyTorch's own bf16 autocast of the reference sits at 1.9e-2).
// build.py variant "bf16" (-DSWN_OPERAND_BF16=1) builds the same kernels with bf16 operands; both variants run the
// whole GPU suite (tests/test_gpu_variants.py).
// ---------------------------------------------------------------------------------------------
#ifndef SWN_OPERAND_BF16
#define SWN_OPERAND_BF16 0
#endif
#if SWN_OPERAND_BF16
#define SWN_MMA_T "bf16"
using op_t = __nv_bfloat16;
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float op_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float op_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
#else
#define SWN_MMA_T "f16"
using op_t = __half;
// fp32 pair -> packed fp16 pair, SATURATING (one F2FP.SATFINITE instruction): values beyond +-65504 clamp instead of
// becoming inf (and NaN after the next subtraction).  Operands that are not bounded by construction — the raw residual
// stream in the A_F32 prologues (PatchExpanding.expand, decoder linears), GELU outputs, folded LN-gamma * W products —
// therefore degrade gracefully on outlier channels instead of poisoning the row (tests/test_gpu_ops.py range tests).
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float op_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float op_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
#endif

// exact-erf GELU (nn.GELU default).  erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7; 6.7e-7 on GELU in
// fp32) with the SFU approximations ex2.approx / rcp.approx: 12 FMA/ALU-pipe + 2 MUFU instructions per element,
// three orders of magnitude below the bf16 rounding that follows it.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-0.5f * 1.4426950408889634f * x * x);  // exp(-(x/sqrt2)^2)
  const float r = fmaf(-p, e, 1.0f);                                 // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), r, hx);                                     // 0.5x(1 + sign(x) r)
}

// Fast GELU for the tensor-core epilogues (the 4C-wide hidden activation makes GELU the dominant CUDA-core cost of
// the MLP kernels).  Phi(x) = 0.5 (1 + tanh(u(x))) with u(x) = x (a0 + a1 x^2 + a2 x^4) least-squares fitted to
// atanh(erf(x / sqrt 2)): |gelu_fast - gelu_erf| <= 3.0e-5 in exact arithmetic (the textbook tanh form is 4.7e-4
// off), plus the MUFU.TANH approximation error (2^-11 relative on tanh) — below the 16-bit rounding that follows.
// 7 FMA-pipe + 1 MUFU instruction per element (gelu_erf: 12 + 2).  x^2 is clamped where the quartic would turn over.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 49.0f);
  const float pz = fmaf(fmaf(-3.58732362e-4f, x2, 3.70503451e-2f), x2, 7.97458471e-1f);
  const float t = tanh_approx(x * pz);
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// GELU of two pre-activations, packed as two 16-bit MMA operands.  fp16 build: the whole evaluation runs on packed
// half2 arithmetic (1 cvt + 6 HFMA2-class + the two halves of tanh.approx.f16x2 for TWO elements, against 2 x 8 + pack
// in fp32); the result is rounded to fp16 anyway, and a float64 replay of this exact op sequence (tests/test_host.py)
// stays within 1.7x the rms error of "fp32 GELU, then one rounding" — a few 1e-4 on O(1) activations.  The bf16 build
// (8 mantissa bits) keeps the fp32 evaluation.  Opt out with -DSWN_GELU_FP32=1.
#ifndef SWN_GELU_FP32
#define SWN_GELU_FP32 0
#endif
__device__ __forceinline__ uint32_t gelu_pack2(float a, float b) {
#if SWN_OPERAND_BF16 || SWN_GELU_FP32
  return pack_op(gelu_fast(a), gelu_fast(b));
#else
  const uint32_t xr = pack_op(a, b);     // saturating conversion
  const __half2 x = *reinterpret_cast<const __half2*>(&xr);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(49.0f));
  __half2 pz = __hfma2(__float2half2_rn(-3.58732362e-4f), x2, __float2half2_rn(3.70503451e-2f));
  pz = __hfma2(pz, x2, __float2half2_rn(7.97458471e-1f));
  const __half2 u = __hmul2(x, pz);
  uint32_t ur = *reinterpret_cast<const uint32_t*>(&u), tr;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tr) : "r"(ur));
  const __half2 t = *reinterpret_cast<const __half2*>(&tr);
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 g = __hfma2(hx, t, hx);
  return *reinterpret_cast<const uint32_t*>(&g);
#endif
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP (the barrier is initialised with the number of warps, not threads): every lane has issued its own
// fences (fence.proxy.async / tcgen05.fence::before_thread_sync) for its own writes, __syncwarp orders them before lane
// 0's releasing arrive.  Measured reason (profiles/r2a_mlp_role_waits.txt): with 256 per-thread arrivals on two barriers
// per hidden chunk the epilogue warps of the MLP kernels spent ~1.6 k cycles per chunk on ~340 useful instructions —
// mbarrier arrivals on one address serialise.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or ~1 ms passes)
// instead of spinning — spinning waiters were measured to eat ~45 % of the issue slots of the working warps.
#ifndef SWN_MBAR_HINT_NS
#define SWN_MBAR_HINT_NS 1000000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SWN_MBAR_HINT_NS)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as
// a hung GPU.  Every legitimate wait in these kernels is microseconds; the retry budget is seconds.
#ifndef SWN_MBAR_BACKOFF_NS
#define SWN_MBAR_BACKOFF_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t tries = 0;
  while (!mbar_try_wait(bar, parity)) {
#if SWN_MBAR_BACKOFF_NS > 0
    __nanosleep(SWN_MBAR_BACKOFF_NS);   // waiting roles must not compete with the working warps for issue slots
#endif
    if ((++tries & 63u) == 0 && clock64() - t0 > 8000000000LL) {
      printf("swn: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Pure polling wait (mbarrier.test_wait, never suspends): for the single warp that watches an MMA-completion barrier
// while the rest of the CTA is parked at a CTA barrier, the wake-up latency of the suspending try_wait form is on the
// critical path of every tile; spinning costs nothing there.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0, tries = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && (++tries & 0xfffffu) == 0 && tries > 0x40000000u) {
      printf("swn: mbarrier spin timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!ok);
}

// ---------------------------------------------------------------------------------------------
// async proxy: bulk copy global -> shared (TMA engine, 1-D), proxy fence
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// bulk copy shared -> global (TMA engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (the buffer may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA, commit, loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 64 bf16 (=128 B),
// 8-row groups 1024 B apart (SBO), version 1 (sm_100).  `addr` = byte address in shared::cta.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffffu) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO = 1024 B  [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                    // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32, both K-major, M x N tile.
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)                              // D format = F32
         | ((uint32_t)SWN_OPERAND_BF16 << 7)    // A format: 0 = F16, 1 = BF16
         | ((uint32_t)SWN_OPERAND_BF16 << 10)   // B format
         | ((N >> 3) << 17)      // N / 8
         | ((M >> 4) << 24);     // M / 16
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp (deterministic for the full mask).  The MMA-issuing warps keep their loops
// warp-uniform and predicate only the tcgen05 instructions on this: a loop under `if (lane == 0)` forces the
// compiler through R2UR/vote sequences per descriptor and was measured to make the single issuing thread
// (~150 cycles of scalar overhead per UMMA) the bottleneck of the whole CTA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// ring-buffer cursor (stage index + phase parity) advanced without integer division
struct RingPos {
  int s;
  uint32_t ph;
  __device__ __forceinline__ void next(int stages) {
    if (++s == stages) {
      s = 0;
      ph ^= 1u;
    }
  }
};
// arrive on `bar` once every previously issued tcgen05 op of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  __syncwarp();  // .sync.aligned: the warp must be converged
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// same, 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// K-major SWIZZLE_128B tile addressing.  A "k-block" is [rows x 64] bf16: row r occupies 128 B at
// r*128; its eight 16-byte chunks are XOR-permuted with (r & 7) (Swizzle<3,4,3>).  Tile bases are
// 1024-byte aligned.  `k` is the element index inside the k-block (0..63).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ row) & 7u) << 4) + ((k & 7u) << 1);
}

// ---------------------------------------------------------------------------------------------
// Coalesced epilogue I/O.  After tcgen05.ld a lane holds 16 consecutive fp32 columns of ITS row, so a direct
// 16-byte global access per lane touches 32 different rows per warp instruction: 32 half-used sectors, and the SM
// injects only ~1 sector request per cycle (measured: this pattern alone cost ~45 % of the proj-GEMM time).  The warp
// therefore transposes the 32 x 16 block through 2 KB of shared memory (16-byte chunks XOR-swizzled with
// (row >> 1) & 3: conflict free both ways); afterwards lane l owns chunk l&3 of rows ps*8 + l/4, ps = 0..3, i.e. one
// warp instruction covers 8 rows x 64 contiguous bytes (fully used sectors).  All lanes of the warp must call these.
// ---------------------------------------------------------------------------------------------
constexpr int EPI_SCRATCH_BYTES = 2048;   // per warp
__device__ __forceinline__ void epi_scatter16(uint8_t* scratch, const float* v, int lane) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4)
    *reinterpret_cast<float4*>(scratch + lane * 64 + ((j4 ^ ((lane >> 1) & 3)) << 4)) =
        make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
  __syncwarp();
}
// chunk (lane & 3) of row ps*8 + (lane >> 2) of the block written by epi_scatter16
__device__ __forceinline__ float4 epi_gather4(const uint8_t* scratch, int ps, int lane) {
  const int rl = ps * 8 + (lane >> 2), c = lane & 3;
  return *reinterpret_cast<const float4*>(scratch + rl * 64 + ((c ^ ((rl >> 1) & 3)) << 4));
}

constexpr int TILE_M = 128;                 // rows per CTA tile = UMMA M
constexpr int KBLK = 64;                    // bf16 elements per swizzle row
constexpr int A_KBLOCK_BYTES = TILE_M * 128;  // one [128 x 64] bf16 k-block

// ---------------------------------------------------------------------------------------------
// Resident A-tile builder shared by the row-tile kernels.  Rows of the 128-row tile are distributed over
// `nwarps` warps; LPR lanes cooperate on one row (so a warp covers 32/LPR rows at once — small channel
// counts keep their lanes busy) and UNR row groups are in flight per warp (all global loads are issued
// before the first reduction, which hides DRAM latency).  `load(row, k)` returns 4 consecutive fp32
// source values (zeros outside the matrix); optional LayerNorm over K; bf16 result is written in the
// K-major SWIZZLE_128B image; columns K..K16 and invalid rows are zero filled.
// ---------------------------------------------------------------------------------------------
template <int LPR, int KV, int UNR, bool LN, class Load>
__device__ __forceinline__ void build_a_tile(uint8_t* a_smem, int K, int K16, const float* __restrict__ ln_w,
                                             const float* __restrict__ ln_b, float eps, int warp, int nwarps, int lane,
                                             Load load, int row_begin = 0, int row_end = TILE_M) {
  constexpr int RPW = 32 / LPR;          // rows per warp pass (row_begin / row_end must be multiples of it)
  const int sub = lane / LPR, sl = lane % LPR;
  const float inv_k = 1.0f / (float)K;
  for (int g0 = row_begin / RPW + warp * UNR; g0 < row_end / RPW; g0 += nwarps * UNR) {
    float4 v[UNR][KV];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int r = (g0 + u) * RPW + sub;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int k = (i * LPR + sl) * 4;
        v[u][i] = (r < row_end && k < K) ? load(r, k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (LN) {
      float mean[UNR], rstd[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < KV; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean[u] = s * inv_k;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          const int k = (i * LPR + sl) * 4;
          if (k < K) {
            const float a = v[u][i].x - mean[u], b = v[u][i].y - mean[u], c = v[u][i].z - mean[u], d = v[u][i].w - mean[u];
            q += (a * a + b * b) + (c * c + d * d);
          }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd[u] = rsqrtf(q * inv_k + eps);
      }
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int k = (i * LPR + sl) * 4;
        if (k < K) {
          const float4 gw = *reinterpret_cast<const float4*>(ln_w + k);
          const float4 gb = *reinterpret_cast<const float4*>(ln_b + k);
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            v[u][i].x = (v[u][i].x - mean[u]) * rstd[u] * gw.x + gb.x;
            v[u][i].y = (v[u][i].y - mean[u]) * rstd[u] * gw.y + gb.y;
            v[u][i].z = (v[u][i].z - mean[u]) * rstd[u] * gw.z + gb.z;
            v[u][i].w = (v[u][i].w - mean[u]) * rstd[u] * gw.w + gb.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int r = (g0 + u) * RPW + sub;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int k = (i * LPR + sl) * 4;
        if (r < row_end && k < K16) {
          const uint2 o = (k < K) ? make_uint2(pack_op(v[u][i].x, v[u][i].y), pack_op(v[u][i].z, v[u][i].w))
                                  : make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(a_smem + (k >> 6) * A_KBLOCK_BYTES + sw128_offset(r, k & 63)) = o;
        }
      }
    }
  }
}

}  // namespace swn
