// Internal (C++) launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace swn {

// ---- rowgemm.cu -----------------------------------------------------------------------------
enum AMode { A_F32_LN = 0, A_F32 = 1, A_BF16 = 2, A_MERGE_LN = 3 };
enum EMode { E_BF16 = 0, E_F32 = 1, E_EXPAND = 2 };

struct RowGemmParams {
  // A operand / prologue
  const void* A;
  int a_mode;
  int M, K, lda;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int gH, gW, gC, gHo, gWo;  // A_MERGE_LN: source grid [B,gH,gW,gC] -> rows over [B,gHo,gWo]
  // B operand: packed [NT x 64] bf16 SWIZZLE_128B tiles, order (chunk, kblock)
  const op_t* Wp;
  int NT, nchunks, n_valid;
  // epilogue
  int e_mode;
  const float* bias;  // padded: nchunks * NT
  void* out;
  int ldo;
  const float* res;
  int ldres;
  const float* alpha;  // device scalar (cross-attention gamma) or null
  int xH, xW, xHs, xWs;  // E_EXPAND: source grid [B,xH,xW], output grid cropped to [xHs,xWs]
  const float* ln2_w;
  const float* ln2_b;
  // filled in by the launcher
  int stages, tmem_cols;
  int stg_stride, pass_rows;  // cp.async input staging: padded row stride (bytes), rows per pass
  int res_stride;             // residual staging row stride (bytes), 0 = residual read from global
};
int launch_rowgemm(RowGemmParams p, cudaStream_t stream);
// warp-level (mma.sync) PatchExpanding for K <= 48 (expand_warp.cu); returns -1 (no error) when the shape does not qualify
int launch_expand_warp(RowGemmParams p, int num_sms, cudaStream_t stream);
// persistent TMA-staged variant for narrow layers; returns -1 (no error) when the shape does not qualify
int launch_rowgemm_persist(RowGemmParams p, int num_sms, cudaStream_t stream);

// ---- mlp.cu ---------------------------------------------------------------------------------
#ifndef SWN_MLP_TWO_CTA
#define SWN_MLP_TWO_CTA 1   // C = 192: two co-resident CTAs per SM (HC = 64, TR = 96, one hidden accumulator); see mlp.cu
#endif
struct MlpParams {
  const float* x;    // [M, C] fp32 residual stream (input)
  float* out;        // [M, C] fp32 (may alias x)
  int M, C;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  const op_t* Wp;  // packed fc1/fc2 tile stream (see pack_mlp_weights in packing.py)
  const float* b1;          // [4C]
  const float* b2;          // [C16] (zero padded)
  int HC;                   // hidden chunk width
  int TR;                   // fc2 output rows per weight tile
  int stages, tmem_cols;    // filled in by the launcher
  int n_hacc;               // mlp.cu: hidden accumulators in TMEM (2 = double buffered)
  int row_stride;           // mlp_persist: staging row stride in bytes (filled in by the launcher)
  long long* phase_cycles;  // optional [16] clock64 sums (profiling aid, see swn_set_phase_profile)
};
int launch_mlp(MlpParams p, cudaStream_t stream);
int launch_mlp_persist(MlpParams p, int num_sms, cudaStream_t stream);  // C <= 96: persistent, TMA-staged

// ---- window_attn.cu -------------------------------------------------------------------------
struct WinAttnParams {
  const op_t* qkv;  // [B*H*W, 3C] token order
  op_t* out;        // [B*H*W, C]
  const float* qkv_bias;     // [3C]: q/k/v of zero-padded tokens (pad happens after norm1)
  const float* rpb_table;    // [81, nH]
  int B, H, W, C, nH, shift;
  const float* bias_frags;   // optional: [nH][2][4][32][4] bias images in accumulator order, log2 domain (packing.py::
                             // rel_pos_bias_fragments); the warp kernel then copies them instead of building them per CTA
};
int launch_window_attn(WinAttnParams p, cudaStream_t stream);

// ---- small_block.cu -------------------------------------------------------------------------
struct SmallBlockParams {   // one whole SwinTransformerBlock, C in {12, 24}, fp32 parameters in nn.Module layout
  const float* x;
  float* out;
  int B, H, W, C, nH, shift;
  float eps;
  const float *n1w, *n1b, *Wqkv, *bqkv, *table, *Wp, *bp, *n2w, *n2b, *W1, *b1, *W2, *b2;
};
int launch_swin_block_small(SmallBlockParams p, int num_sms, cudaStream_t stream);

// ---- swin_fused.cu --------------------------------------------------------------------------
struct FusedBlockParams {   // fused W-MSA (+ MLP) block, shift 0, resident packed weights (packing.py::pack_fused_block)
  const float* x;     // [B, H*W, C] fp32
  float* out;         // [B, H*W, C] fp32 (must not alias x: tiles are read by prefetch while others are written)
  int B, H, W, C, nH;
  float eps;
  const op_t* Wpk;    // [Wqkv | Wproj | W1 chunks | W2 chunks] as SWIZZLE_128B k-block images
  const float* fpk;   // bqkv' bqkv bproj b1' b2 bias-fragment images (see packing.py)
  int do_mlp;         // 1: whole block, 0: attention half only (x + proj(attn(LN1 x)))
  // filled in by the launcher
  int K16, NQ, HC, nj, n_hs, n_stage, ones_col, RS, RSB;
  int nWy, nWx, ntiles;
  long long n_windows;
  int w_bytes, nf, u_bytes;
  int off_a, off_u, off_stage, off_f, off_misc;
  int tm_y, tmem_cols;
  long long* phase_cycles;   // optional [grid][16] per-phase clock64 sums written by thread 0 of each CTA (profiling aid)
};
int launch_swin_fused(FusedBlockParams p, int num_sms, cudaStream_t stream);

// ---- swin_warp.cu ---------------------------------------------------------------------------
struct WarpBlockParams {    // whole block, shift 0, C in {12, 24}: one warp per window (packing.py::pack_warp_block)
  const float* x;     // [B, H*W, C] fp32
  float* out;         // [B, H*W, C] fp32 (must not alias x)
  int B, H, W, C, nH;
  float eps;
  const op_t* Wpk;    // per block: weight fragments q | k | v^T | proj | fc1 | fc2       (depth blocks back to back)
  const float* fpk;   // per block: b2 [K16] | relative-position bias fragments [nH][2][4][32][4]
  int depth;          // consecutive blocks (same window partition: shift 0) applied to the rows in one pass
  int nWy, nWx, n_windows;   // filled in by the launcher
};
int launch_swin_warp_block(WarpBlockParams p, int num_sms, cudaStream_t stream);

// ---- cross_attn.cu --------------------------------------------------------------------------
struct CrossAttnParams {
  const op_t* q;   // [B, Lq, C]
  const op_t* kv;  // [B, Lk, 2C]  (k | v)
  op_t* out;       // [B, Lq, C]
  int B, Lq, Lk, C, nH;
};
int launch_cross_attn(CrossAttnParams p, cudaStream_t stream);

// ---- elementwise.cu -------------------------------------------------------------------------
int launch_patch_embed(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b,
                       float* out, int B, int Cin, int H, int W, int Ho, int Wo, int scale, cudaStream_t s);
int launch_seg_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* lowres, float* out, int B, int Hq, int Wq, int up, int Hout, int Wout, cudaStream_t s);
int launch_recon_head(const float* tok, const float* w1, const float* b1, const float* w2, const float* b2,
                      float* out, int B, int Hh, int Wh, int Cout, int Hout, int Wout, cudaStream_t s);
int launch_copy_cols(const float* src, int lds, float* dst, int ldd, long long rows, int cols, cudaStream_t s);
int launch_sigmoid_mask(const float* img, int Cimg, const float* seg, float* images2, float* seg_map, float* masked,
                        float* minmax, int B, int Cout, int H, int W, cudaStream_t s);
int launch_dspace_hist(const float* img, long long img_stride, const int* bin_of_pixel, int B, int n_pix, int n_bins,
                       float* out, cudaStream_t s);
int launch_ensure_2ch(const float* x, float* out, int B, int HW, cudaStream_t s);
int launch_normalize(const float* x, const float* minmax, float* out, int BC, int H, int W, float thr, float eps,
                     int inverse, cudaStream_t s);

// ---- train_ops.cu ---------------------------------------------------------------------------
struct AdamWTensor {      // mirror of swn_param_desc (include/swinwnet_b200.h)
  float* p;               // parameter (fp32 master, updated in place)
  float* g;               // gradient or null (frozen / unused this step)
  float* m;               // exp_avg
  float* v;               // exp_avg_sq
  long long n;            // elements
  long long flat_off;     // element offset inside the flat all-reduce bucket
  float bc1, bc2_sqrt;    // 1 - beta1^t, sqrt(1 - beta2^t) with t = this TENSOR's step count (torch keeps it per parameter)
};
int launch_adamw_multi(const AdamWTensor* tab, const int2* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                       float wd, float grad_scale, cudaStream_t s);
int launch_bucket_copy(const AdamWTensor* tab, const int2* chunks, int n_chunks, float* flat, int mode, float scale, cudaStream_t s);

}  // namespace swn
