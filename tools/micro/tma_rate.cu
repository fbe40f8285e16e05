// Micro-benchmark: throughput of 1-D bulk copies (cp.async.bulk global -> shared) as a function of the copy size.
// One CTA per SM streams a tile of `rows x row_bytes` (contiguous in global memory) per iteration into a double-buffered
// shared-memory staging area, either as `rows` copies of row_bytes (padded destination rows) or as rows/group copies of
// group*row_bytes.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tma_rate.cu -o gpurun_out/tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, long long tile_bytes, int ntiles, int rows, int row_bytes, int group,
                                            int dst_stride, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2];
  uint8_t* stg = smem;
  const int buf_bytes = rows * dst_stride;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  };
  if (warp == 0) {
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (lane == 0) {
        wait(&empty[s], ((it >> 1) & 1) ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(rows * row_bytes) : "memory");
      }
      __syncwarp();
      const uint8_t* g = src + (long long)tile * tile_bytes;
      const int ncopies = rows / group;
      for (int c = lane; c < ncopies; c += 32)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(stg + s * buf_bytes + c * group * dst_stride)),
                     "l"(g + (long long)c * group * row_bytes), "r"(group * row_bytes), "r"(s32(&full[s]))
                     : "memory");
    }
  } else if (warp == 1) {
    int it = 0;
    unsigned long long acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      wait(&full[s], (it >> 1) & 1);
      acc += *reinterpret_cast<const unsigned long long*>(stg + s * buf_bytes + lane * 8);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
    if (acc == 0x1234567) sink[0] = acc;
  }
}

int main() {
  const int rows = 128, ntiles = 15000;
  uint8_t* src;
  unsigned long long* sink;
  cudaMalloc(&src, (size_t)ntiles * rows * 768);
  cudaMemset(src, 1, (size_t)ntiles * rows * 768);
  cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int row_bytes_list[] = {48, 96, 192, 384, 768};
  for (int rb : row_bytes_list)
    for (int group : {1, 2, 4, 8, 32, 128}) {
      const int dst_stride = group == 1 ? rb + 16 : rb;
      const size_t smem = 2 * (size_t)rows * dst_stride;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<148, 128, smem>>>(src, (long long)rows * rb, ntiles, rows, rb, group, dst_stride, sink);
      cudaEventRecord(e0);
      k<<<148, 128, smem>>>(src, (long long)rows * rb, ntiles, rows, rb, group, dst_stride, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)ntiles * rows * rb;
      const double copies = (double)ntiles * rows / group;
      printf("row %4d B  x%3d rows per copy (%6d B): %7.3f ms  %7.1f GB/s  %6.1f ns per copy per SM  err=%d\n", rb, group, rb * group, ms,
             bytes / ms / 1e6, ms * 1e6 / (copies / 148), (int)cudaGetLastError());
    }
  return 0;
}
