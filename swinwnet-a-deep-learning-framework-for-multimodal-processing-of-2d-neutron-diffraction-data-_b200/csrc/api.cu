// extern "C" boundary of libswinwnet_b200.so (declared in include/swinwnet_b200.h).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/swinwnet_b200.h"
#ifndef SWN_TUNING_HOOKS
#define SWN_TUNING_HOOKS 0
#endif
#include "common.cuh"
#include "kernels.h"

namespace swn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static void mlp_config(int C, int* HC, int* TR) {
  const int C16 = (C + 15) & ~15, Hd = 4 * C, hbase = (C16 + 31) & ~31;
  // C <= 96: HC = 64 keeps smem/TMEM small enough for two co-resident CTAs per SM (these widths are bound by the
  // CUDA-core epilogue, not the tensor pipe); wider layers use HC = 128 (full-rate N) when TMEM allows.
  // C > 96: HC = 128 (full-rate N for GEMM1; at C = 384 TMEM then holds only ONE hidden accumulator next to the 384 Y
  // columns — measured 13 % faster than two 64-column accumulators: 0.60 -> 0.52 ms at M = 122 880)
  (void)hbase;
#if SWN_MLP_TWO_CTA
  if (C == 192) { *HC = 64; *TR = 96; return; }     // two co-resident CTAs per SM (mlp.cu)
#endif
  if (C > 96 && Hd % 128 == 0) *HC = 128;
#ifdef SWN_MLP96_HC
  else if (C == 96) *HC = SWN_MLP96_HC;
#else
  else if (C == 96) *HC = 96;              // persistent kernel: 4 chunks of 96 measured 5 % faster than 6 of 64
#endif
  else if (Hd % 64 == 0) *HC = 64;
  else *HC = Hd;
  *TR = C16 <= 256 ? C16 : (C16 % 128 == 0 ? 128 : C16 / 2);
#if SWN_TUNING_HOOKS
  // tuning build only (build.py variant "prof", tools/bench_ops.py sweeps): SWN_MLP_HC / SWN_MLP_TR override the tiling
  if (C >= 96) {
    if (const char* e = getenv("SWN_MLP_HC")) { const int v = atoi(e); if (v >= 16 && v % 16 == 0 && Hd % v == 0) *HC = v; }
    if (const char* e = getenv("SWN_MLP_TR")) { const int v = atoi(e); if (v >= 16 && v % 16 == 0 && C16 % v == 0) *TR = v; }
  }
#endif
}
static long long* g_phase_cycles = nullptr;
#ifndef SWN_MLP_PERSIST_MAX_C_DEFAULT
#define SWN_MLP_PERSIST_MAX_C_DEFAULT 96   // 192 routes C = 192 through the DIRECT (global-read LayerNorm) persistent variant
#endif
static int mlp_persist_max_c() {
#if SWN_TUNING_HOOKS
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SWN_MLP_PERSIST_MAX_C");
    v = e ? atoi(e) : SWN_MLP_PERSIST_MAX_C_DEFAULT;
  }
  return v;
#else
  return SWN_MLP_PERSIST_MAX_C_DEFAULT;
#endif
}
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
}  // namespace swn

using namespace swn;

extern "C" {

const char* swn_last_error(void) { return g_err; }
int swn_abi_version(void) { return SWN_ABI_VERSION; }
#ifndef SWN_BUILD_DIGEST
#define SWN_BUILD_DIGEST "unknown"
#endif
const char* swn_build_digest(void) { return SWN_BUILD_DIGEST; }
int swn_sizeof_rowgemm_args(void) { return (int)sizeof(swn_rowgemm_args); }
int swn_operand_is_bf16(void) { return SWN_OPERAND_BF16; }

int swn_mlp_config(int C, int* HC, int* TR) {
  SWN_CHECK(C >= 4 && C % 4 == 0 && C <= 384 && HC && TR, "mlp_config: unsupported C=%d", C);
  mlp_config(C, HC, TR);
  return 0;
}

int swn_rowgemm(const swn_rowgemm_args* a, void* stream) {
  SWN_CHECK(a != nullptr, "rowgemm: null args");
  RowGemmParams p{};
  p.A = a->A; p.a_mode = a->a_mode; p.M = a->M; p.K = a->K; p.lda = a->lda;
  p.ln_w = a->ln_w; p.ln_b = a->ln_b; p.ln_eps = a->ln_eps;
  p.gH = a->gH; p.gW = a->gW; p.gC = a->gC; p.gHo = a->gHo; p.gWo = a->gWo;
  p.Wp = reinterpret_cast<const op_t*>(a->Wp);
  p.NT = a->NT; p.nchunks = a->nchunks; p.n_valid = a->n_valid;
  p.e_mode = a->e_mode; p.bias = a->bias; p.out = a->out; p.ldo = a->ldo;
  p.res = a->res; p.ldres = a->ldres; p.alpha = a->alpha;
  p.xH = a->xH; p.xW = a->xW; p.xHs = a->xHs; p.xWs = a->xWs; p.ln2_w = a->ln2_w; p.ln2_b = a->ln2_b;
  SWN_CHECK(p.A && p.Wp && p.out, "rowgemm: null pointer");
  SWN_CHECK(p.a_mode >= 0 && p.a_mode <= 3 && p.e_mode >= 0 && p.e_mode <= 2, "rowgemm: bad mode");
  if (p.a_mode == A_F32_LN || p.a_mode == A_MERGE_LN) SWN_CHECK(p.ln_w && p.ln_b, "rowgemm: LayerNorm params missing");
  if (p.e_mode == E_EXPAND) SWN_CHECK(p.ln2_w && p.ln2_b && p.nchunks == 4, "rowgemm: expand needs 4 chunks + LN params");
  SWN_CHECK(p.M > 0 && p.K > 0 && p.K % 4 == 0 && p.NT >= 16 && p.NT <= 256 && p.NT % 16 == 0 && p.n_valid > 0 &&
                p.n_valid <= p.NT && p.n_valid % 4 == 0 && p.nchunks >= 1 && p.ldo % 4 == 0,
            "rowgemm: bad shape arguments");
  int rc = launch_expand_warp(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
  if (rc >= 0) return rc;
  rc = launch_rowgemm_persist(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
  if (rc >= 0) return rc;
  return launch_rowgemm(p, reinterpret_cast<cudaStream_t>(stream));
}

int swn_mlp(const float* x, float* out, int M, int C, const float* ln_w, const float* ln_b, float ln_eps, const void* Wp,
            const float* b1, const float* b2, void* stream) {
  SWN_CHECK(x && out && ln_w && ln_b && Wp && b1 && b2, "mlp: null pointer");
  SWN_CHECK(C >= 4 && C % 4 == 0 && C <= 384, "mlp: unsupported C=%d", C);
  MlpParams p{};
  p.x = x; p.out = out; p.M = M; p.C = C; p.ln_w = ln_w; p.ln_b = ln_b; p.ln_eps = ln_eps;
  p.Wp = reinterpret_cast<const op_t*>(Wp); p.b1 = b1; p.b2 = b2;
  mlp_config(C, &p.HC, &p.TR);
  p.phase_cycles = g_phase_cycles;
  // persistent kernel: staged rows for C <= 96; SWN_MLP_PERSIST_MAX_C=192 opts into its direct-LayerNorm variant
  if (C <= mlp_persist_max_c()) return launch_mlp_persist(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
  return launch_mlp(p, reinterpret_cast<cudaStream_t>(stream));
}

int swn_swin_block_small(const float* x, float* out, int B, int H, int W, int C, int num_heads, int shift, float eps,
                         const float* const* w, void* stream) {
  SWN_CHECK(x && out && w, "swin_block_small: null pointer");
  for (int i = 0; i < 13; ++i) SWN_CHECK(w[i] != nullptr, "swin_block_small: null parameter pointer %d", i);
  SmallBlockParams p{x, out, B, H, W, C, num_heads, shift, eps,
                     w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], w[9], w[10], w[11], w[12]};
  return launch_swin_block_small(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}

int swn_swin_block_fused(const float* x, float* out, int B, int H, int W, int C, int num_heads, float eps,
                         const void* Wpk, const float* fpk, int do_mlp, void* stream) {
  SWN_CHECK(x && out && Wpk && fpk, "swin_block_fused: null pointer");
  SWN_CHECK(x != out, "swin_block_fused: out must not alias x");
  FusedBlockParams p{};
  p.x = x; p.out = out; p.B = B; p.H = H; p.W = W; p.C = C; p.nH = num_heads; p.eps = eps;
  p.Wpk = reinterpret_cast<const op_t*>(Wpk); p.fpk = fpk; p.do_mlp = do_mlp;
  p.phase_cycles = g_phase_cycles;
  return launch_swin_fused(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}

int swn_swin_block_warp(const float* x, float* out, int B, int H, int W, int C, int num_heads, float eps, const void* Wpk,
                        const float* fpk, int depth, void* stream) {
  SWN_CHECK(x && out && Wpk && fpk, "swin_block_warp: null pointer");
  SWN_CHECK(x != out, "swin_block_warp: out must not alias x");
  WarpBlockParams p{};
  p.x = x; p.out = out; p.B = B; p.H = H; p.W = W; p.C = C; p.nH = num_heads; p.eps = eps;
  p.Wpk = reinterpret_cast<const op_t*>(Wpk); p.fpk = fpk; p.depth = depth;
  return launch_swin_warp_block(p, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}

int swn_set_phase_profile(void* device_buffer) {
  g_phase_cycles = reinterpret_cast<long long*>(device_buffer);
  return 0;
}

int swn_window_attention(const void* qkv, void* out, const float* qkv_bias, const float* rpb_table, int B, int H, int W,
                         int C, int num_heads, int shift, void* stream) {
  SWN_CHECK(qkv && out && qkv_bias && rpb_table, "window_attention: null pointer");
  SWN_CHECK(B > 0 && H > 0 && W > 0 && C % 4 == 0 && num_heads > 0 && shift >= 0, "window_attention: bad sizes");
  WinAttnParams p{reinterpret_cast<const op_t*>(qkv), reinterpret_cast<op_t*>(out), qkv_bias, rpb_table,
                  B, H, W, C, num_heads, shift, nullptr};
  return launch_window_attn(p, reinterpret_cast<cudaStream_t>(stream));
}

int swn_window_attention_frags(const void* qkv, void* out, const float* qkv_bias, const float* rpb_table, const float* bias_frags,
                               int B, int H, int W, int C, int num_heads, void* stream) {
  SWN_CHECK(qkv && out && qkv_bias && rpb_table && bias_frags, "window_attention_frags: null pointer");
  SWN_CHECK(B > 0 && H > 0 && W > 0 && C % 4 == 0 && num_heads > 0, "window_attention_frags: bad sizes");
  WinAttnParams p{reinterpret_cast<const op_t*>(qkv), reinterpret_cast<op_t*>(out), qkv_bias, rpb_table,
                  B, H, W, C, num_heads, 0, bias_frags};
  return launch_window_attn(p, reinterpret_cast<cudaStream_t>(stream));
}

int swn_cross_attention(const void* q, const void* kv, void* out, int B, int Lq, int Lk, int C, int num_heads,
                        void* stream) {
  SWN_CHECK(q && kv && out, "cross_attention: null pointer");
  CrossAttnParams p{reinterpret_cast<const op_t*>(q), reinterpret_cast<const op_t*>(kv),
                    reinterpret_cast<op_t*>(out), B, Lq, Lk, C, num_heads};
  return launch_cross_attn(p, reinterpret_cast<cudaStream_t>(stream));
}

int swn_patch_embed(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, float* out,
                    int B, int Cin, int H, int W, int Ho, int Wo, int scale, void* stream) {
  SWN_CHECK(x && w && b && ln_w && ln_b && out, "patch_embed: null pointer");
  return launch_patch_embed(x, w, b, ln_w, ln_b, out, B, Cin, H, W, Ho, Wo, scale, reinterpret_cast<cudaStream_t>(stream));
}

int swn_seg_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2, float* lowres,
                 float* out, int B, int Hq, int Wq, int up, int Hout, int Wout, void* stream) {
  SWN_CHECK(tok && w1 && b1 && w2 && b2 && lowres && out, "seg_head: null pointer");
  SWN_CHECK(Hout <= Hq * up && Wout <= Wq * up, "seg_head: crop larger than upsampled map");
  return launch_seg_head(tok, w1, b1, w2, b2, lowres, out, B, Hq, Wq, up, Hout, Wout, reinterpret_cast<cudaStream_t>(stream));
}

int swn_recon_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                   int B, int Hh, int Wh, int Cout, int Hout, int Wout, void* stream) {
  SWN_CHECK(tok && w1 && b1 && w2 && b2 && out, "recon_head: null pointer");
  return launch_recon_head(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout, reinterpret_cast<cudaStream_t>(stream));
}

int swn_copy_cols(const float* src, int lds, float* dst, int ldd, long long rows, int cols, void* stream) {
  SWN_CHECK(src && dst, "copy_cols: null pointer");
  return launch_copy_cols(src, lds, dst, ldd, rows, cols, reinterpret_cast<cudaStream_t>(stream));
}

int swn_sigmoid_mask(const float* img, int Cimg, const float* seg, float* images2, float* seg_map, float* masked,
                     float* minmax, int B, int Cout, int H, int W, void* stream) {
  SWN_CHECK(img && seg && masked, "sigmoid_mask: null pointer");
  return launch_sigmoid_mask(img, Cimg, seg, images2, seg_map, masked, minmax, B, Cout, H, W,
                             reinterpret_cast<cudaStream_t>(stream));
}

int swn_dspace_histogram(const float* img, long long img_stride, const int* bin_of_pixel, int B, int n_pixels, int n_bins,
                         float* out, void* stream) {
  SWN_CHECK(img && bin_of_pixel && out, "dspace_histogram: null pointer");
  return launch_dspace_hist(img, img_stride, bin_of_pixel, B, n_pixels, n_bins, out, reinterpret_cast<cudaStream_t>(stream));
}

int swn_ensure_2ch(const float* x, float* out, int B, int HW, void* stream) {
  SWN_CHECK(x && out && B > 0 && HW > 0, "ensure_2ch: bad arguments");
  return launch_ensure_2ch(x, out, B, HW, reinterpret_cast<cudaStream_t>(stream));
}

int swn_normalize(const float* x, const float* minmax, float* out, int BC, int H, int W, float threshold, float eps,
                  int inverse, void* stream) {
  SWN_CHECK(x && minmax && out, "normalize: null pointer");
  return launch_normalize(x, minmax, out, BC, H, W, threshold, eps, inverse, reinterpret_cast<cudaStream_t>(stream));
}

int swn_adamw_multi(const swn_param_desc* table, const int32_t* chunks, int n_chunks, double lr, double beta1, double beta2,
                    double eps, double weight_decay, double grad_scale, void* stream) {
  SWN_CHECK(table && chunks && n_chunks > 0, "adamw_multi: bad arguments");
  static_assert(sizeof(swn_param_desc) == sizeof(AdamWTensor), "swn_param_desc layout");
  return launch_adamw_multi(reinterpret_cast<const AdamWTensor*>(table), reinterpret_cast<const int2*>(chunks), n_chunks, (float)lr,
                            (float)beta1, (float)beta2, (float)eps, (float)weight_decay, (float)grad_scale,
                            reinterpret_cast<cudaStream_t>(stream));
}

int swn_grad_bucket_copy(const swn_param_desc* table, const int32_t* chunks, int n_chunks, float* flat, int unpack, double scale,
                         void* stream) {
  SWN_CHECK(table && chunks && flat && n_chunks > 0, "grad_bucket_copy: bad arguments");
  return launch_bucket_copy(reinterpret_cast<const AdamWTensor*>(table), reinterpret_cast<const int2*>(chunks), n_chunks, flat,
                            unpack ? 1 : 0, (float)scale, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
