/* C ABI of libswinwnet_b200.so — the sm_100a kernels behind the SwinWNet forward hot path.
 *
 * The reference (popoff4rtem/SwinWNet-...) is pure Python/PyTorch and has NO plugin / FFI layer: its
 * operator surface for this path is the nn.Module method set SwinWNet.segment_1 / upscale / segment_2
 * (/root/reference/SwinWNet.py:886-957) plus SwinUNet.forward (:574) and SwinUNetSR.forward (:740).
 * The drop-in module in the package (model.py) keeps that surface and lowers every ATen op group the
 * reference dispatches onto the entry points below, bound with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (tensor.data_ptr()), dense row-major, 16-byte aligned;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and CUDA-graph capturable;
 *   - nothing is allocated internally; outputs/workspaces are caller-owned;
 *   - return 0 on success, non-zero on error (1 = unsupported shape/argument, 2 = CUDA error);
 *     swn_last_error() returns a thread-local message.  There is NO CPU fallback.
 */
#ifndef SWINWNET_B200_H
#define SWINWNET_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWN_ABI_VERSION 1

/* A-operand prologues of swn_rowgemm */
#define SWN_A_F32_LN 0   /* fp32 rows -> LayerNorm -> bf16          (norm1+qkv :242,:185; norm_q/kv+in_proj :779-782) */
#define SWN_A_F32 1      /* fp32 rows -> bf16                       (PatchExpanding.expand :402; decoder linears :489) */
#define SWN_A_BF16 2     /* bf16 rows                               (attention output -> proj :207 / out_proj :782)    */
#define SWN_A_MERGE_LN 3 /* 2x2 gather [x00,x10,x01,x11] + LayerNorm (PatchMerging :295-312)                           */
/* epilogues of swn_rowgemm */
#define SWN_E_BF16 0   /* out_bf16 = acc + bias                                                                     */
#define SWN_E_F32 1    /* out_f32 = alpha*(acc + bias) + residual      (proj+shortcut :277; q + gamma*attn :783)      */
#define SWN_E_EXPAND 2 /* pixel-shuffle scatter + LayerNorm(C/2) + crop (PatchExpanding :404-410, crop_to_res :414)   */

typedef struct swn_rowgemm_args {
  const void* A; int32_t a_mode; int32_t M, K, lda;
  const float* ln_w; const float* ln_b; float ln_eps;
  int32_t gH, gW, gC, gHo, gWo;
  const void* Wp;            /* packed bf16 weight tiles, see swn_rowgemm doc */
  int32_t NT, nchunks, n_valid;
  int32_t e_mode; const float* bias; void* out; int32_t ldo;
  const float* res; int32_t ldres; const float* alpha;
  int32_t xH, xW, xHs, xWs; const float* ln2_w; const float* ln2_b;
} swn_rowgemm_args;

const char* swn_last_error(void);
int swn_abi_version(void);
const char* swn_build_digest(void); /* sha256 of the sources + nvcc flags this library was built from (build.py) */
int swn_sizeof_rowgemm_args(void); /* lets FFI bindings verify their struct mirror */
int swn_operand_is_bf16(void);     /* 16-bit tensor-core operand type of this build: 0 = IEEE fp16 (default), 1 = bf16 */

/* Tile configuration of the fused MLP for channel width C (the weight packer must use the same):
 * HC = hidden-chunk width, TR = fc2 output rows per weight tile. */
int swn_mlp_config(int C, int* HC, int* TR);

/* out = epilogue(prologue(A)[M,K] * W[N,K]^T), N = nchunks*n_valid, on tcgen05 tensor cores.
 * Wp: for chunk n, k-block kb (64 K-elements): a [NT x 64] bf16 tile in the UMMA K-major SWIZZLE_128B
 * image (row r at r*128 B, 16-byte chunk c stored at chunk c ^ (r & 7)), rows >= n_valid and
 * k >= K zero; tiles ordered (n, kb).  bias: nchunks*NT floats (padded). */
int swn_rowgemm(const swn_rowgemm_args* args, void* stream);

/* out[M,C] = x + fc2(GELU(fc1(LayerNorm(x))))   (SwinTransformerBlock MLP, SwinWNet.py:226-234,278).
 * Wp = tile stream in consumption order, see packing.py::pack_mlp_weights; b2 padded to ceil16(C). */
int swn_mlp(const float* x, float* out, int M, int C, const float* ln_w, const float* ln_b, float ln_eps,
            const void* Wp, const float* b1, const float* b2, void* stream);

/* One whole SwinTransformerBlock (SwinWNet.py:236-280) for the narrow UpscalingHead layers, C in {12,24}, 3 heads,
 * as a single fp32 kernel.  w[13] = device pointers to the block's parameters in nn.Module layout, in the order
 * norm1.weight, norm1.bias, attn.qkv.weight [3C,C], attn.qkv.bias, attn.relative_position_bias_table [81,nH],
 * attn.proj.weight [C,C], attn.proj.bias, norm2.weight, norm2.bias, mlp.0.weight [4C,C], mlp.0.bias,
 * mlp.3.weight [C,4C], mlp.3.bias.  `w` itself is a HOST array.  out may alias x. */
int swn_swin_block_small(const float* x, float* out, int B, int H, int W, int C, int num_heads, int shift, float eps,
                         const float* const* w, void* stream);

/* One whole SwinTransformerBlock with shift_size 0 (SwinWNet.py:236-280), or only its attention half
 * x + proj(W-MSA(LayerNorm1(x))) when do_mlp == 0, as ONE tcgen05 kernel for C <= 64: window partition / reverse /
 * padding are index math, q/k/v, probabilities and the hidden activation never leave the SM.
 * Wpk / fpk = packed 16-bit weight images and fp32 vectors, see packing.py::pack_fused_block.  out must not alias x. */
int swn_swin_block_fused(const float* x, float* out, int B, int H, int W, int C, int num_heads, float eps,
                         const void* Wpk, const float* fpk, int do_mlp, void* stream);

/* `depth` (1..4) consecutive SwinTransformerBlocks with shift_size 0 (SwinWNet.py:236-280; a whole BasicLayer,
 * SwinWNet.py:320-345, whose blocks all share one window partition) for C = 12 / 24 with 3 heads (the UpscalingHead layers
 * at 500x960 / 250x480 tokens, SwinWNet.py:656-678): one warp per 5x5 window, every intermediate — including the rows
 * between the blocks — in mma.sync register fragments, no block-level synchronisation.  Wpk / fpk = the blocks' weight
 * fragments and fp32 vectors back to back, see packing.py::pack_warp_block.  out must not alias x. */
int swn_swin_block_warp(const float* x, float* out, int B, int H, int W, int C, int num_heads, float eps, const void* Wpk,
                        const float* fpk, int depth, void* stream);

/* Profiling aid: when set to a device buffer of [grid][16] int64 (zeroed by the caller), the fused block kernels add the
 * clock64 cycles thread 0 of each CTA spends in each barrier-delimited phase.  NULL (default) disables it. */
int swn_set_phase_profile(void* device_buffer);

/* 5x5 (shifted-)window attention core on token-ordered qkv (SwinWNet.py:86-149,183-206,246-272). */
int swn_window_attention(const void* qkv_bf16, void* out_bf16, const float* qkv_bias, const float* rpb_table,
                         int B, int H, int W, int C, int num_heads, int shift, void* stream);
/* The same for shift 0 with the relative-position bias of every head already expanded to the [nH][2][4][32][4] fp32
 * accumulator-fragment images of the kernel (log2 domain, packing.py::rel_pos_bias_fragments): the kernel copies 4 KB per
 * head instead of rebuilding the images from the [81, nH] table in every CTA (a third of the launch at 16x30 tokens). */
int swn_window_attention_frags(const void* qkv_bf16, void* out_bf16, const float* qkv_bias, const float* rpb_table,
                               const float* bias_frags, int B, int H, int W, int C, int num_heads, void* stream);

/* flash-style global cross attention core, heads of 64 or 128 channels (SwinWNet.py:782). */
int swn_cross_attention(const void* q_bf16, const void* kv_bf16, void* out_bf16, int B, int Lq, int Lk, int C,
                        int num_heads, void* stream);

/* ScaleAwarePatchEmbed: 2x2 conv, stride 2*scale, dilation scale, + LayerNorm(48) (SwinWNet.py:53-82). */
int swn_patch_embed(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b,
                    float* out_tokens, int B, int Cin, int H, int W, int Ho, int Wo, int scale, void* stream);

/* SegmentationHead: conv3x3(48->24)+GELU+conv1x1(24->1) -> lowres[B,Hq,Wq]; bilinear x`up`; crop
 * (SwinWNet.py:507-531). */
int swn_seg_head(const float* tokens, const float* w1, const float* b1, const float* w2, const float* b2,
                 float* lowres, float* out, int B, int Hq, int Wq, int up, int Hout, int Wout, void* stream);

/* UpscalingHead tail: conv3x3(12->12)+GELU+conv1x1(12->Cout), NCHW, cropped (SwinWNet.py:682-688,932). */
int swn_recon_head(const float* tokens, const float* w1, const float* b1, const float* w2, const float* b2,
                   float* out, int B, int Hh, int Wh, int Cout, int Hout, int Wout, void* stream);

/* dst[r, 0:cols] = src[r, 0:cols]  (decoder skip concat, SwinWNet.py:483). */
int swn_copy_cols(const float* src, int lds, float* dst, int ldd, long long rows, int cols, void* stream);

/* ST pipeline glue (ST_Inference_Pipline.py:32-37,90-97,127-134): images2 = ensure_2ch(img) (optional),
 * seg_map = sigmoid(seg), masked = images * seg_map, minmax[B*Cout][2] = per-image amin/amax (optional). */
int swn_sigmoid_mask(const float* img, int Cimg, const float* seg, float* images2, float* seg_map, float* masked,
                     float* minmax, int B, int Cout, int H, int W, void* stream);

/* d-space front end of the physics metrics (Qwrapper.tensor_to_d, Diffraction_metrics.py:35-70), whole batch in one
 * launch: out[b, bin] = sum over pixels p with bin_of_pixel[p] == bin of img[b * img_stride + p] (channel 0 of image b;
 * img_stride in floats).  bin_of_pixel = int32 [n_pixels], -1 = pixel dropped (d > 7.5); built once per geometry by
 * physics.py with the reference's own fp32 bucketize.  out [B, n_bins] fp32 is zeroed by the call. */
int swn_dspace_histogram(const float* img, long long img_stride, const int* bin_of_pixel, int B, int n_pixels, int n_bins,
                         float* out, void* stream);

/* ensure_2ch (ST_Inference_Pipline.py:32-37): out[b,0] = x[b,0], out[b,1] = sqrt(|x[b,0]|); x [B,1,HW], out [B,2,HW]. */
int swn_ensure_2ch(const float* x, float* out, int B, int HW, void* stream);

/* normalize_piecewise (inverse=0) / denormalize_piecewise (inverse=1) (ST_Inference_Pipline.py:39-67). */
int swn_normalize(const float* x, const float* minmax, float* out, int BC, int H, int W, float threshold, float eps,
                  int inverse, void* stream);

/* ---- training side ---------------------------------------------------------------------- */
/* One entry per parameter tensor; `g` may be null (frozen module or a branch unused in this step: the even / odd steps
 * of FullModel_supervised_trainer.py:231-288 leave different sub-modules without gradients). */
typedef struct swn_param_desc {
  float* p; float* g; float* m; float* v;
  int64_t n;          /* elements */
  int64_t flat_off;   /* element offset of this tensor inside the flat gradient bucket */
  float bc1, bc2_sqrt; /* AdamW bias corrections 1 - beta1^t, sqrt(1 - beta2^t); t = step count of THIS tensor (torch counts
                          steps per parameter: a parameter without a gradient in some steps lags behind) */
} swn_param_desc;

/* torch.optim.AdamW step for ALL parameters in one launch (the optimizer the reference trainers build:
 * FullModel_supervised_trainer.py:85-92): p *= 1 - lr*wd; m, v updates; p -= lr/bc1 * m / (sqrt(v)/bc2_sqrt + eps).
 * `table` = device array of descriptors, `chunks` = device int32 pairs (tensor index, element offset), one CTA per
 * 4096-element chunk; tensors with g == NULL are skipped; gradients are multiplied by grad_scale first. */
int swn_adamw_multi(const swn_param_desc* table, const int32_t* chunks, int n_chunks, double lr, double beta1, double beta2,
                    double eps, double weight_decay, double grad_scale, void* stream);

/* Data-parallel gradient bucket: unpack=0 gathers every gradient (zeros where g is null) into the flat fp32 bucket that
 * NCCL all-reduces; unpack=1 scatters flat*scale back into the gradients that exist. */
int swn_grad_bucket_copy(const swn_param_desc* table, const int32_t* chunks, int n_chunks, float* flat, int unpack,
                         double scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
