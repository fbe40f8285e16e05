// Training-side kernels of the drop-in: one multi-tensor AdamW launch for all parameters of the
// model (torch.optim.AdamW semantics as built by the reference trainers: FullModel_supervised_trainer.py:85-92,
// Segmentator_pretrain.py:124-131 — decoupled weight decay, bias-corrected moments, eps added after the sqrt), and a
// bucket pack / unpack pair for the data-parallel gradient all-reduce (dist.py).  HBM-bound: 16 B read + 12 B written
// per parameter, float4 accesses, one CTA per 4096-element chunk of one tensor (chunk table built once on the host).
#include "common.cuh"
#include "kernels.h"

namespace swn {

namespace {

constexpr int TO_CHUNK = 4096, TO_THREADS = 256;

__global__ void __launch_bounds__(TO_THREADS) adamw_multi_kernel(const AdamWTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                                                 float lr, float beta1, float beta2, float eps, float wd,
                                                                 float grad_scale) {
  const int2 ch = chunks[blockIdx.x];           // (tensor index, element offset)
  const AdamWTensor t = tab[ch.x];
  if (t.g == nullptr) return;                   // parameter without a gradient this step (frozen / unused branch)
  const long long n = t.n;
  const long long base = ch.y;
  const float step_size = lr / t.bc1, decay = 1.0f - lr * wd, bc2_sqrt = t.bc2_sqrt;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= grad_scale;
    p *= decay;
    m = fmaf(beta1, m, (1.0f - beta1) * g);
    v = fmaf(beta2, v, (1.0f - beta2) * g * g);
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step_size * (m / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  if (vec) {
    for (long long i = base + threadIdx.x * 4; i < base + TO_CHUNK && i + 3 < n; i += TO_THREADS * 4) {
      float4 p = *reinterpret_cast<float4*>(t.p + i), m = *reinterpret_cast<float4*>(t.m + i), v = *reinterpret_cast<float4*>(t.v + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
    }
    // tail of the tensor (n % 4 elements) belongs to the chunk that contains it
    const long long tail0 = n & ~3ll;
    if (tail0 >= base && tail0 < base + TO_CHUNK) {
      const long long i = tail0 + threadIdx.x;
      if (i < n) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
    }
  } else {
    for (long long i = base + threadIdx.x; i < base + TO_CHUNK && i < n; i += TO_THREADS) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  }
}

// gradients <-> flat all-reduce bucket (mode 0: pack (missing gradient -> zeros), mode 1: unpack * scale)
__global__ void __launch_bounds__(TO_THREADS) bucket_copy_kernel(const AdamWTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                                                 float* __restrict__ flat, int mode, float scale) {
  const int2 ch = chunks[blockIdx.x];
  const AdamWTensor t = tab[ch.x];
  float* f = flat + t.flat_off;
  for (long long i = ch.y + threadIdx.x; i < ch.y + TO_CHUNK && i < t.n; i += TO_THREADS) {
    if (mode == 0) f[i] = t.g ? t.g[i] : 0.f;
    else if (t.g) t.g[i] = f[i] * scale;
  }
}

}  // namespace

int launch_adamw_multi(const AdamWTensor* tab, const int2* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                       float wd, float grad_scale, cudaStream_t s) {
  SWN_CHECK(tab && chunks && n_chunks > 0, "adamw: bad arguments");
  adamw_multi_kernel<<<n_chunks, TO_THREADS, 0, s>>>(tab, chunks, lr, beta1, beta2, eps, wd, grad_scale);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

int launch_bucket_copy(const AdamWTensor* tab, const int2* chunks, int n_chunks, float* flat, int mode, float scale, cudaStream_t s) {
  SWN_CHECK(tab && chunks && flat && n_chunks > 0 && (mode == 0 || mode == 1), "bucket_copy: bad arguments");
  bucket_copy_kernel<<<n_chunks, TO_THREADS, 0, s>>>(tab, chunks, flat, mode, scale);
  SWN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace swn
