// Micro-benchmark of the CUDA-core / TMEM primitives that bound the epilogues of the tcgen05 kernels on B200:
// MUFU.TANH (f32, f16x2), MUFU.EX2, the packed-half GELU of common.cuh, and tcgen05.ld.  One CTA per SM, W warps per
// CTA; prints issue cycles per warp-instruction per SM sub-partition.   nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define SWN_OPERAND_BF16 0
#include "../../swinwnet-a-deep-learning-framework-for-multimodal-processing-of-2d-neutron-diffraction-data-_b200/csrc/common.cuh"
using namespace swn;
namespace swn { void set_error(const char*, ...) {} }

constexpr int ITERS = 2000, UNROLL = 8;

template <int MODE>
__global__ void k_alu(float* out, long long* cyc) {
  float x[UNROLL];
  uint32_t h[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { x[i] = 0.001f * (threadIdx.x + i); h[i] = pack_op(x[i], -x[i]); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (MODE == 0) x[i] = tanh_approx(x[i]);
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) x[i] = ex2_approx(x[i]);
      if (MODE == 3) h[i] = gelu_pack2(x[i], __uint_as_float(h[i]) * 1e-30f + 0.5f) ^ h[i];
      if (MODE == 4) x[i] = gelu_fast(x[i]);
      if (MODE == 5) h[i] = pack_op(fmaf(x[i], 1.0001f, 0.5f), x[i]) + h[i];
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < UNROLL; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int X32>
__global__ void k_tmem(float* out, long long* cyc) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t lane_addr = base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float v[32];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t col = (uint32_t)(((it * 4 + i) * (X32 ? 32 : 16) + (warp >> 2) * 64) & 511 & ~(X32 ? 31 : 15));
      if (X32) tmem_ld32(lane_addr + col, v); else tmem_ld16(lane_addr + col, v);
    }
    tmem_ld_wait();
    acc += v[0] + v[15];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base_s, 512); }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const char* names[] = {"tanh.approx.f32", "tanh.approx.f16x2", "ex2.approx.f32", "gelu_pack2 (2 elem)", "gelu_fast f32 (1 elem)", "fma+pack (F2FP)"};
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      auto launch = [&](auto kern) { kern<<<148, warps * 32>>>(out, cyc); };
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) launch(k_alu<0>); if (mode == 1) launch(k_alu<1>); if (mode == 2) launch(k_alu<2>);
        if (mode == 3) launch(k_alu<3>); if (mode == 4) launch(k_alu<4>); if (mode == 5) launch(k_alu<5>);
      }
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per = (double)h[0] / (ITERS * UNROLL) / (warps / 4.0);   // cycles per warp-op per SMSP
      printf("%-24s warps/SM %2d: %7.2f cycles per warp-op per SMSP  (%s)\n", names[mode], warps, per, cudaGetErrorString(cudaGetLastError()));
    }
  for (int x32 = 0; x32 < 2; ++x32)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) { if (x32) k_tmem<1><<<148, warps * 32>>>(out, cyc); else k_tmem<0><<<148, warps * 32>>>(out, cyc); }
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per = (double)h[0] / (ITERS * 4) / (warps / 4.0);
      const double bytes = 32.0 * (x32 ? 32 : 16) * 4;
      printf("tcgen05.ld x%-2d            warps/SM %2d: %7.2f cycles per ld per SMSP -> %.1f B/clk/SM  (%s)\n", x32 ? 32 : 16, warps, per, bytes / per * 4,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
