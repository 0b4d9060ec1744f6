/*
 * gvl.h — C ABI of libgvl_sm100a.so: the B200-native per-frame perception embedding path.
 *
 * Scope: decoded uint8 frames -> SigLIP2-so400m-patch14-384 vision tower (MAP-pooled
 * 1152-d embedding) -> ProjectorBank MLP (1152 -> 4096 -> 4096) -> timeline index + cosine top-k.
 *
 * The reference (chasemetoyer/gameplay-vision-llm) is pure Python and has NO FFI for this path: its
 * seam is a set of Python call signatures whose arithmetic runs inside HuggingFace `transformers`
 * and `torch.nn`.  Every entry point below cites the reference call it replaces (paths relative to
 * the reference root; `HF:` = transformers 5.5.0, `TV:` = torchvision 0.26, `ATen:` = torch 2.11).
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every `const void*` / `void*` data pointer is a DEVICE
 *    pointer unless the parameter name starts with `h_` (host).
 *  - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream).  All work is
 *    stream-ordered; no entry point synchronises the device.
 *  - return value 0 = ok; non-zero = error, message via gvl_last_error() (thread-local).
 *  - no CPU fallback exists: on a machine without an sm_100 GPU every compute entry point fails.
 *  - bf16 = raw uint16_t storage of IEEE bfloat16.
 */
#ifndef GVL_H_
#define GVL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVL_ABI_VERSION 11

#if defined(__GNUC__)
#define GVL_API __attribute__((visibility("default")))
#else
#define GVL_API
#endif

/* ---- status / introspection --------------------------------------------------------------- */

/* Last error message of the calling thread ("" if none). */
GVL_API const char* gvl_last_error(void);
/* Returns GVL_ABI_VERSION. Host only. */
GVL_API int gvl_abi_version(void);
/* Number of kernel launches issued by this library since load (all threads). Host only.
 * bench.py reports the delta over the timed region as `gpu_launches`. */
GVL_API unsigned long long gvl_launch_count(void);
/* 0 if device `dev` is an sm_100 GPU this library can run on. */
GVL_API int gvl_check_device(int dev);

/* Per-launch timing for bench.py's roofline: while enabled, every kernel launch of this library is bracketed
 * by CUDA events on its own stream.  gvl_prof_enable(1) starts a fresh session, gvl_prof_enable(0) stops
 * recording (the session stays readable); gvl_prof_summary sums one kernel family (it waits for the
 * recorded events).  total_work = algorithmic FLOPs (GEMM, attention) or bytes (the others). */
enum {
    GVL_K_PREPROCESS = 0, GVL_K_GEMM = 1, GVL_K_LAYERNORM = 2, GVL_K_ATTENTION = 3, GVL_K_PROBE_ATTENTION = 4,
    GVL_K_TOPK_SCORES = 5, GVL_K_TOPK_SELECT = 6, GVL_K_PATCHIFY = 7, GVL_K_COUNT = 8
};
GVL_API int gvl_prof_enable(int on);
GVL_API int gvl_prof_summary(int kernel_id, double* total_ms, unsigned long long* launches, double* total_work);

/* ---- K1: frame preprocessing --------------------------------------------------------------- */
/*
 * Replaces `self.encoder._processor(images=[image], return_tensors="pt")`
 * (src/perception/siglip_semantic_encoder.py:474-477) = HF SiglipImageProcessor, torchvision
 * backend: uint8 antialiased resize (HF:image_processing_backends.py:200-251 ->
 * TV:v2/functional/_geometry.py:271-340 -> ATen:native/cpu/UpSampleKernelAVXAntialias.h:304-437,
 * integer two-pass, horizontal then vertical, int16 weights, uint8 intermediate) followed by the fused
 * rescale+normalize `(float(u8) - sub) / div` (HF:image_processing_backends.py:292-331), and — for
 * GVL_LAYOUT_BF16_PATCH — the cast to bf16 and the im2col of the stride-14 patch-embedding Conv2d
 * (HF:models/siglip/modeling_siglip.py:177-179).
 */
enum {
    GVL_RESAMPLE_BILINEAR = 2, /* PIL.Image.BILINEAR */
    GVL_RESAMPLE_BICUBIC = 3   /* PIL.Image.BICUBIC (a = -0.5) */
};
enum {
    GVL_LAYOUT_U8_CHW = 0,    /* uint8  [B,3,out_h,out_w]  resized image (exactness checks) */
    GVL_LAYOUT_F32_CHW = 1,   /* float  [B,3,out_h,out_w]  == HF `pixel_values` */
    GVL_LAYOUT_BF16_CHW = 2,  /* bf16   [B,3,out_h,out_w]  == pixel_values.to(bfloat16) */
    GVL_LAYOUT_BF16_PATCH = 3 /* bf16   [B*gh*gw, ld]; row = b*gh*gw + py*gw + px, col = c*p*p + ky*p + kx,
                                 gh = out_h / p, gw = out_w / p; cols [3*p*p, ld) are written as zero */
};

/* Host only: integer tap tables of one axis, exactly as ATen builds them
 * (ATen:native/cpu/UpSampleKernel.cpp `_compute_index_ranges_int16_weights`).
 * h_xmin/h_xsize: [out_size]; h_weights: [out_size*max_taps] (zero padded); *h_precision: shift. */
GVL_API int gvl_resize_taps(int in_size, int out_size, int resample, int max_taps,
                    int32_t* h_xmin, int32_t* h_xsize, int16_t* h_weights, int* h_precision,
                    int* h_taps_used);

/* frames: uint8 [B,H,W,3] (HWC RGB).  sub/div: host float[3], value = (u8 - sub[c]) / div[c] in fp32.
 * patch / ld are used by GVL_LAYOUT_BF16_PATCH only (ld in elements, ld >= 3*patch*patch, ld % 8 == 0). */
GVL_API int gvl_preprocess_u8(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample,
                      const float* h_sub, const float* h_div, void* out, int layout, int patch, int ld,
                      void* stream);

/* Kernel selection of gvl_preprocess_u8 (process-wide; tests and A/B runs — every path is bit-identical):
 * AUTO = the streaming 5:1 kernel (preprocess_stream.cu) where its geometry applies, else the planar kernel, else v1. */
enum { GVL_PRE_PATH_AUTO = 0, GVL_PRE_PATH_PLANAR = 1, GVL_PRE_PATH_V1 = 2 };
GVL_API int gvl_preprocess_path(int path);

/* pixel_values: float [B,3,H,W] (what the HF processor returns) -> bf16 im2col rows [B*gh*gw, ld], i.e. the
 * `.to(bfloat16)` + patch extraction at the head of `get_image_features(pixel_values=...)`
 * (HF:models/siglip/modeling_siglip.py:175-179).  Same row/column order as GVL_LAYOUT_BF16_PATCH. */
GVL_API int gvl_patchify_f32(const float* pixel_values, int B, int H, int W, int patch, int ld, void* out,
                             void* stream);

/* Same as gvl_preprocess_u8 for the CHW layouts, but only the window [crop_y0, crop_y0+crop_h) x [crop_x0,
 * crop_x0+crop_w) of the resized out_h x out_w image is produced: out is [B,3,crop_h,crop_w].  Replaces
 * `VideoMAEImageProcessor(pil_frames, return_tensors="pt")` (scripts/extract_features.py:345-375 ->
 * HF:models/videomae/image_processing_videomae.py:35-105: shortest-edge-224 uint8 antialias resize, center crop
 * 224, fused rescale+normalize). */
GVL_API int gvl_preprocess_u8_crop(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int crop_y0,
                           int crop_x0, int crop_h, int crop_w, int resample, const float* h_sub,
                           const float* h_div, void* out, int layout, void* stream);

/* gvl_preprocess_u8_crop on frames of which only the column band [band_x0, band_x0 + band_w) is resident:
 * frames is uint8 [B, H, band_w, 3].  A center crop of a 16:9 frame reads ~57 % of every source row, so a host feed
 * (scripts/extract_features.py:345-375 hands whole PIL frames to the processor) only has to move that band.
 * band_x0 % 16 == 0; the band must hold every source column the crop window's taps read (gvl_resize_taps gives
 * them: xmin[crop_x0] .. xmin[crop_x0 + crop_w - 1] + xsize[...]).  Same arithmetic, bit-identical output. */
GVL_API int gvl_preprocess_u8_crop_band(const uint8_t* frames, int B, int H, int W, int band_x0, int band_w, int out_h,
                                int out_w, int crop_y0, int crop_x0, int crop_h, int crop_w, int resample,
                                const float* h_sub, const float* h_div, void* out, int layout, void* stream);

/* Host -> device copy of a byte band of every row: dst[r][0:band_bytes] = src_host[r][band_offset : +band_bytes]
 * (cudaMemcpy2DAsync; src_host pinned for an asynchronous copy).  The feed side of gvl_preprocess_u8_crop_band. */
GVL_API int gvl_copy_band_h2d(void* dst, const void* src_host, long long rows, long long src_row_bytes,
                      long long band_offset_bytes, long long band_bytes, void* stream);

/* pixel_values: bf16 [clips*frames, 3, H, W] (frames of a clip consecutive) -> bf16 tubelet im2col rows
 * [clips*(frames/tubelet)*(H/p)*(W/p), 3*tubelet*p*p] for the Conv3d(kernel = stride = (tubelet,p,p)) patch
 * embedding (HF:models/videomae/modeling_videomae.py:157-178: permute to (B,C,T,H,W), conv, flatten(2),
 * transpose): row = ((clip*(frames/tubelet) + t/tubelet)*gh + py)*gw + px, col = c*tubelet*p*p + (t%tubelet)*p*p +
 * ky*p + kx.  patch % 8 == 0. */
GVL_API int gvl_patchify_tubelet_bf16(const void* pixel_values, int clips, int frames, int H, int W, int patch,
                              int tubelet, void* out, void* stream);

/* ---- K2: tcgen05 GEMM with fused epilogue ---------------------------------------------------- */
/*
 * out[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + residual[(row % res_row_mod), N]
 * Replaces every nn.Linear / Conv2d-as-GEMM on the path:
 *   patch embedding + position embedding   HF:models/siglip/modeling_siglip.py:124-130,175-186
 *   q/k/v/out projections                  HF:models/siglip/modeling_siglip.py:270-273,285-287,309
 *   MLP fc1 + GELU(tanh) + fc2 + residual  HF:models/siglip/modeling_siglip.py:315-327,354-359
 *   MAP head in_proj/out_proj/MLP          HF:models/siglip/modeling_siglip.py:628-649
 *   projector Linear-GELU(erf)-Linear      src/agent_core/qwen_reasoning_core.py:1009-1013
 * A, W: bf16 row-major (K contiguous), lda/ldw in elements (multiples of 8), 16-byte aligned.
 * bias: float[N] or NULL.  residual: bf16, ldr elements, or NULL; res_row_mod = 0 -> row index as is.
 * out: bf16 (out_f32 = 0) or float (out_f32 = 1), ldo elements.  N % 8 == 0, K % 8 == 0.
 * fp32 accumulation in tensor memory; one rounding to the output type.
 */
enum { GVL_ACT_NONE = 0, GVL_ACT_GELU_TANH = 1, GVL_ACT_GELU_ERF = 2 };

GVL_API int gvl_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias,
                  const void* residual, int ldr, int res_row_mod, void* out, int ldo, int out_f32,
                  int M, int N, int K, int act, void* stream);

/* LayerNorm folded into the GEMMs on either side of it (used by gvl_siglip_forward when the weight pack says
 * fold_ln): the GEMM that WRITES the residual stream also writes, per output row and per slab of 64 columns (padded
 * to an even slab count, at most 20), the partial sums (sum x, sum x^2) of the bf16 values it stored; the GEMM that CONSUMES LayerNorm(x)
 * reads x itself as A, uses weights pre-multiplied by gamma, and applies
 *     out[m,n] = act( rstd[m] * (acc[m,n] - mean[m] * c1[n]) + c2[n] ),   c1[n] = sum_k W'[n,k],
 *     c2[n] = bias[n] + sum_k beta[k] W[n,k]   (passed through `bias`),
 * with mean / rstd from the partial sums added in a fixed order (deterministic; no atomics).  Algebraically
 * identical to LayerNorm followed by the Linear (HF:models/siglip/modeling_siglip.py:340-361); it removes one
 * read + write of the [M,D] activations per LayerNorm. */
typedef struct gvl_gemm_fusion {
    float* stats_out;      /* producer: float [M, gvl_gemm_stats_slots(N), 2], or NULL */
    const float* ln_stats; /* consumer: statistics of the rows of A written by a producer GEMM, or NULL */
    int32_t ln_slots;      /* slabs per row in ln_stats: even, <= 20; ln_stats 16-byte aligned.  0 = ln_stats is the
                            * finalised float [M, 2] array (rstd, mean * rstd) written by gvl_ln_finalize */
    int32_t ln_dim;        /* number of columns the statistics cover (D) */
    const float* ln_c1;    /* [N] */
    float ln_eps;
} gvl_gemm_fusion;
GVL_API int gvl_gemm_stats_slots(int N); /* slabs per row a producer GEMM with N output columns writes. Host only. */
/* out[m] = (rstd, mean * rstd) of row m from a producer's partial sums stats [rows, slots, 2] over `dim` columns; same
 * summation order as the consumer epilogue, so both routes give identical bits. */
GVL_API int gvl_ln_finalize(const float* stats, int rows, int slots, int dim, float eps, float* out, void* stream);
GVL_API int gvl_gemm_bf16_fused(const void* A, int lda, const void* W, int ldw, const float* bias,
                        const void* residual, int ldr, int res_row_mod, void* out, int ldo, int out_f32,
                        int M, int N, int K, int act, const gvl_gemm_fusion* fusion, void* stream);

/* ---- K3: LayerNorm -------------------------------------------------------------------------- */
/* y = (x - mean) / sqrt(var + eps) * gamma + beta over the last dim, fp32 statistics.
 * Replaces nn.LayerNorm (HF:models/siglip/modeling_siglip.py:334-336,596,636). x,y bf16; D % 4 == 0. */
GVL_API int gvl_layernorm_bf16(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                       int rows, int D, float eps, void* stream);

/* ---- K4: self-attention ---------------------------------------------------------------------- */
/* qkv: bf16 [B*T, 3*H*hd] (q | k | v, head-major inside each third); out: bf16 [B*T, H*hd].
 * softmax(q k^T * scale) v, non-causal, no mask, fp32 softmax.
 * Replaces SiglipAttention's SDPA call (HF:models/siglip/modeling_siglip.py:293-306; eager
 * definition :229-249).  hd = 64 or 72 (the two instantiations; anything else returns an error). */
GVL_API int gvl_attention_bf16(const void* qkv, void* out, int B, int T, int H, int hd, float scale,
                       void* stream);

/* Ragged batch (masked-region route): items of different lengths back to back on the token axis, ONE launch.
 * gvl_attention_varlen_tiles (host only): h_item_tokens int32 [n_items] -> h_tiles int32 [*n_tiles, 4] = {first token of
 * the item, item length, first query row of the tile, 0}, longest items first; h_tiles NULL = size query
 * (*n_tiles = sum ceil(T_i / 128)).  gvl_attention_varlen_bf16: qkv bf16 [M_total, 3*H*hd], out bf16 [M_total, H*hd],
 * tiles = that table in DEVICE memory (16-byte aligned); score_elems = sum T_i^2 (profiling only).  Each item's rows are
 * bit-identical to gvl_attention_bf16(B = 1, T = T_i) on that item. */
GVL_API int gvl_attention_varlen_tiles(int n_items, const int32_t* h_item_tokens, int32_t* h_tiles, int* n_tiles);
GVL_API int gvl_attention_varlen_bf16(const void* qkv, void* out, int M_total, const void* tiles, int n_tiles,
                              double score_elems, int H, int hd, float scale, void* stream);

/* ---- K5: MAP-head probe attention -------------------------------------------------------------- */
/* One query (the learned probe, already projected and pre-scaled: q float[H*hd]) attends over the T
 * keys/values of each image.  kv: bf16 [B*T, 2*H*hd] (k | v).  out: bf16 [B, H*hd].
 * Replaces nn.MultiheadAttention(probe, h, h) inside SiglipMultiheadAttentionPoolingHead
 * (HF:models/siglip/modeling_siglip.py:639-643); the q/k/v/out projections run through gvl_gemm_bf16. */
GVL_API int gvl_probe_attention_bf16(const float* q, const void* kv, void* out, int B, int T, int H, int hd,
                             void* stream);

/* ---- K6: whole-tower forward -------------------------------------------------------------------- */
/* Device-pointer weight pack, repacked by the host from the HF state_dict
 * (names: SURVEY.md §8a row M1; src/perception/siglip_semantic_encoder.py:195-204 loads them). */
typedef struct gvl_vit_layer {
    const float *ln1_g, *ln1_b;
    const void* w_qkv;  /* bf16 [3D, D]  = cat(q_proj, k_proj, v_proj).weight */
    const float* b_qkv; /* [3D] */
    const void* w_o;    /* bf16 [D, D] */
    const float* b_o;
    const float *ln2_g, *ln2_b;
    const void* w_fc1; /* bf16 [I, D] */
    const float* b_fc1;
    const void* w_fc2; /* bf16 [D, I] */
    const float* b_fc2;
    /* fold_ln packs only: w_qkv / w_fc1 hold W * gamma (ln1 / ln2), b_qkv / b_fc1 hold c2, and these hold c1 */
    const float* c1_qkv; /* [3D] */
    const float* c1_fc1; /* [I] */
} gvl_vit_layer;

typedef struct gvl_vit_weights {
    int32_t D, I, H, hd, L, T, patch_k, patch_ld; /* hidden, intermediate, heads, head dim, layers,
                                                    tokens/image, 3*p*p, padded patch row length */
    float eps;
    int32_t act; /* GVL_ACT_* of the encoder MLPs */
    const void* w_patch;  /* bf16 [D, patch_ld] (zero padded beyond patch_k) */
    const float* b_patch; /* [D] */
    const void* pos;      /* bf16 [T, D] */
    const gvl_vit_layer* layers; /* HOST array of L entries */
    const float *post_g, *post_b;
    /* MAP head */
    const float* probe_q;  /* [D] = (probe . Wq^T + bq) * hd^-0.5, precomputed on the host in fp32 */
    const void* w_kv;      /* bf16 [2D, D] = in_proj_weight[D:3D] */
    const float* b_kv;     /* [2D] */
    const void* w_ho;      /* bf16 [D, D]  = attention.out_proj.weight */
    const float* b_ho;
    const float *hln_g, *hln_b;
    const void* w_hfc1; /* bf16 [I, D] */
    const float* b_hfc1;
    const void* w_hfc2; /* bf16 [D, I] */
    const float* b_hfc2;
    /* LayerNorm folding (gvl_gemm_fusion): when fold_ln != 0 the per-layer LayerNorms and post_layernorm are not run
     * as kernels; w_kv / b_kv hold the folded MAP-head K/V projection and c1_kv its column sums.  ln*_g / ln*_b and
     * post_g / post_b must still be valid (last_hidden output, VideoMAE final norm). */
    int32_t fold_ln;
    const float* c1_kv; /* [2D] */
} gvl_vit_weights;

/* Bytes of scratch gvl_siglip_forward needs for a batch of B images. Host only. */
GVL_API size_t gvl_siglip_workspace_bytes(const gvl_vit_weights* w, int B);

/* patches: bf16 [B*T, patch_ld] (GVL_LAYOUT_BF16_PATCH output).  pooled: bf16 [B, D].
 * last_hidden: optional bf16 [B*T, D] (post-layernorm tokens) or NULL.
 * Replaces `self.encoder._model.get_image_features(**inputs)`
 * (src/perception/siglip_semantic_encoder.py:479-481 -> HF:models/siglip/modeling_siglip.py:787-816,
 * 604-625). */
GVL_API int gvl_siglip_forward(const gvl_vit_weights* w, const void* patches, int B, void* workspace,
                       size_t workspace_bytes, void* pooled, void* last_hidden, void* stream);

/* Ragged batch (the masked-region route, K9): n_groups groups, group g = B items of T tokens each with its own position
 * table pos (bf16 [T, D], device), token rows concatenated in group order.  Row-wise work (GEMMs, LayerNorms) runs once
 * over all rows; the patch GEMM's position add, the attention and the MAP-head probe attention run per group.  Each
 * item's rows are bit-identical to a gvl_siglip_forward call on a pack with that T / pos (every kernel's per-row
 * arithmetic is independent of the rows around it).  patches: bf16 [sum B*T, patch_ld]; pooled: bf16 [sum B, D];
 * last_hidden: optional bf16 [sum B*T, D].  w->T / w->pos are ignored. */
typedef struct gvl_ragged_group {
    int32_t B, T;
    const void* pos;
} gvl_ragged_group;
GVL_API size_t gvl_siglip_ragged_workspace_bytes(const gvl_vit_weights* w, long long M_total, int B_total);
GVL_API int gvl_siglip_forward_ragged(const gvl_vit_weights* w, const void* patches, int n_groups,
                              const gvl_ragged_group* groups, void* workspace, size_t workspace_bytes, void* pooled,
                              void* last_hidden, void* stream);

/* ---- K6b: VideoMAE clip encoder ------------------------------------------------------------------- */
/* out[b, :] = mean over the T tokens of x[b] (bf16 [B,T,D]); out bf16 or float [B,D].
 * Replaces `outputs.last_hidden_state.mean(dim=1)` (scripts/extract_features.py:381). */
GVL_API int gvl_mean_tokens_bf16(const void* x, int B, int T, int D, void* out, int out_f32, void* stream);

/* The VideoMAE encoder takes the same weight pack as the SigLIP tower with: T = (frames/tubelet)*(H/p)*(W/p),
 * patch_k = patch_ld = 3*tubelet*p*p, pos = the fixed sinusoid table (bf16 [T,D]), act = GVL_ACT_GELU_ERF,
 * eps = 1e-12, b_qkv = cat(q_bias, 0, v_bias) (the key projection has no bias), post_g/post_b = the final
 * LayerNorm or NULL when the checkpoint was trained with use_mean_pooling (no final norm); the MAP-head fields
 * are unused.  patches: bf16 [B*T, patch_ld] (gvl_patchify_tubelet_bf16).  pooled: bf16 or float [B, D].
 * last_hidden: optional bf16 [B*T, D].  Replaces `VideoMAEModel(**inputs).last_hidden_state.mean(dim=1)`
 * (scripts/extract_features.py:377-381 -> HF:models/videomae/modeling_videomae.py:407-475). */
GVL_API size_t gvl_videomae_workspace_bytes(const gvl_vit_weights* w, int B);
GVL_API int gvl_videomae_forward(const gvl_vit_weights* w, const void* patches, int B, void* workspace,
                         size_t workspace_bytes, void* pooled, int pooled_f32, void* last_hidden, void* stream);

/* ---- K7: projector ------------------------------------------------------------------------------ */
/* out = W2 . gelu_erf(W1 . x + b1) + b2.  Replaces MultiModalProjector.forward
 * (src/agent_core/qwen_reasoning_core.py:1007-1027) as used by ProjectorBank.project_region (:1076).
 * x bf16 [M, enc]; w1 bf16 [llm, enc]; w2 bf16 [llm, llm]; hidden: scratch bf16 [M, llm];
 * out: bf16 or float [M, llm]. */
GVL_API int gvl_project(const void* x, int M, int enc_dim, int llm_dim, const void* w1, const float* b1,
                const void* w2, const float* b2, void* hidden, void* out, int out_f32, void* stream);

/* ---- K8: cosine top-k over the timeline index ------------------------------------------------------ */
/* For each of Q queries: score_n = <q, e_n> / (max(|q|,eps) * max(|e_n|,eps)), fp32; returns the k
 * best in (score descending, index ascending) order.  Replaces TimelineRetriever.retrieve_by_semantic
 * (src/agent_core/qwen_reasoning_core.py:1492-1528: cos_sim + argsort(descending)[:k]) and
 * SigLIPSemanticEncoder.find_similar_regions (src/perception/siglip_semantic_encoder.py:616-638:
 * stable sort => ties keep the lower index first).
 * index: bf16 [N, D] (ld = D); queries: bf16 [Q, D]; scratch: float [gvl_topk_scratch_floats(N, Q)], 16-byte aligned;
 * out_scores: float [Q,k]; out_idx: int32 [Q,k] (-1 / -inf when fewer than k rows are eligible); k <= 64; D % 8 == 0. */
GVL_API int gvl_topk_cosine(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                    float* scratch, float* out_scores, int32_t* out_idx, void* stream);
GVL_API size_t gvl_topk_scratch_floats(int N, int Q); /* Host only. */

/* Scoring path: SCAN = CUDA-core fp32 scan, the index is read once per 8 queries (any Q, lowest latency for one
 * query); TENSOR = one skinny tcgen05 GEMM scores = Q . E^T (fp32 out) that reads the index once for the whole query
 * batch, a scale pass with 1/|q| and 1/|e_n|, and an exact second stage: every row within 6e-5 of the provisional
 * k-th score is re-scored with SCAN's fp32 arithmetic and the final k are chosen among those, so both paths return
 * identical indices and scores (more than 512 such candidates for one query: its provisional result is kept and the
 * int at float offset Q*ld + ld + roundup4(Q) of scratch, ld = roundup4(N), counts the queries this happened to).  AUTO = TENSOR for Q > 8 (more than one
 * scan pass) over >= 4096 rows. */
enum { GVL_TOPK_AUTO = 0, GVL_TOPK_SCAN = 1, GVL_TOPK_TENSOR = 2 };

/* gvl_topk_cosine restricted, per query, to the rows [row_lo[q], row_hi[q]) of the index.  The timeline index is in
 * timestamp order, so this is the reference's "events within +-window seconds of t" filter
 * (scripts/realtime_inference.py:988-998 `abs(e["timestamp"] - timestamp) < window`;
 * src/agent_core/qwen_reasoning_core.py:1462-1490 retrieve_by_timestamp, start <= ts <= end) fused with the ranking
 * of retrieve_by_semantic: rows outside a query's range are neither scored nor ranked.
 * row_lo / row_hi: device int32 [Q], both NULL = whole index; [span_lo, span_hi) = union of the ranges (host-known:
 * only these rows are scored).  inv_norm: optional device float [N] holding 1/max(|e_n|, eps) (gvl_row_inv_norm), so
 * repeated searches over an unchanged index skip the norm pass of the TENSOR path; NULL = computed into scratch. */
GVL_API int gvl_topk_cosine_ex(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                       const int32_t* row_lo, const int32_t* row_hi, int span_lo, int span_hi, int mode,
                       const float* inv_norm, float* scratch, float* out_scores, int32_t* out_idx, void* stream);
/* gvl_topk_cosine for fp32 rows and queries (index: float [N, D], queries: float [Q, D], D % 4 == 0, scan path only):
 * SigLIPSemanticEncoder.compute_similarity / find_similar_regions upcast both sides with .float() before
 * F.cosine_similarity (src/perception/siglip_semantic_encoder.py:604-638), so fp32 embeddings (VideoMAE clip vectors,
 * fp32 projections) must not be ranked through a bf16 cast. */
GVL_API int gvl_topk_cosine_f32(const float* index, int N, int D, const float* queries, int Q, int k, float eps,
                        float* scratch, float* out_scores, int32_t* out_idx, void* stream);
/* inv_norm[n] = 1 / max(|rows[n]|, eps), rows bf16 [N, D]. */
GVL_API int gvl_row_inv_norm(const void* rows, int N, int D, float eps, float* inv_norm, void* stream);

/* ---- K9: masked-region variant ---------------------------------------------------- */
/* The region route of `SigLIPSemanticEncoder.encode_masked_regions` (src/perception/siglip_semantic_encoder.py:485-562,
 * called from scripts/extract_features.py:551-585 when SAM detections exist): bounding-box crops of one frame are
 * resized with `PIL.Image.resize(..., BICUBIC)` (:155), scaled by 1/255 and normalised with the ImageNet constants in
 * fp32 (:357-365), zero-padded to the largest region of the batch (:527-537), cast to the model dtype (:539) and
 * encoded; `mean` / `max` / `cls` pooling (:426-443).
 *
 * Host only: Pillow's 8-bit coefficient tables of one axis (Pillow src/libImaging/Resample.c `precompute_coeffs` with
 * the bicubic filter, a = -0.5, support 2 scaled by max(in/out, 1), + `normalize_coeffs_8bpc`, 22-bit fixed point).
 * h_xmin / h_count: [out_size]; h_coeffs: [out_size * max_taps] (zero padded); *h_ksize = taps per output (the
 * stride Pillow allocates).  With all three table pointers NULL only *h_ksize is written. */
GVL_API int gvl_pil_bicubic_taps(int in_size, int out_size, int max_taps, int32_t* h_xmin, int32_t* h_count,
                         int32_t* h_coeffs, int* h_ksize);

/* One region = GVL_REGION_DESC_INTS int32 (HOST array, validated, uploaded by the call):
 *   [0] x1  [1] y1  [2] crop_w  [3] crop_h   source box inside the frame (`frame[y1:y2, x1:x2]`, :342)
 *   [4] out_w  [5] out_h                      `AspectPreservingResizer.compute_optimal_size` (:97-135)
 *   [6] kh  [7] kv                            tap strides of the two tables (>= gvl_pil_bicubic_taps' ksize)
 *   [8] h_off  [9] v_off                      int32 offsets into `tabs` of the horizontal / vertical table,
 *                                             each laid out as xmin[out] | count[out] | coeffs[out * k] */
#define GVL_REGION_DESC_INTS 10
GVL_API size_t gvl_region_scratch_bytes(int R, const int32_t* h_desc); /* Host only. */

/* frame: device uint8 [H, W, 3]; tabs: device int32 [tabs_ints]; lut: device bf16 [3, 256] = the model-dtype value of
 * ((v / 255) - mean[c]) / std[c] evaluated in fp32 (built by the host with the reference's own three fp32 operations).
 * patches: bf16 [R * (canvas_h/patch) * (canvas_w/patch), ld], GVL_LAYOUT_BF16_PATCH order per region; pixels outside a
 * region's out_h x out_w rectangle and the columns [3*patch*patch, ld) are zero (`F.pad` of the normalised tensor, :536).
 * canvas_h = canvas_w = 0 is the RAGGED form: every region on its own out_h x out_w canvas (multiples of the patch size),
 * patch rows of the regions back to back in `patches` (bf16 [sum (out_h/patch)*(out_w/patch), ld]) — the input layout of
 * gvl_siglip_forward_ragged when regions of equal size are adjacent; resized_u8 must be NULL then.
 * resized_u8: optional uint8 [R, canvas_h, canvas_w, 3] = the resized regions themselves (zero outside), bit-identical
 * to Pillow's bytes (horizontal pass into a uint8 intermediate, then vertical); either output may be NULL. */
GVL_API int gvl_region_patches_pil_u8(const uint8_t* frame, int H, int W, int R, const int32_t* h_desc,
                              const int32_t* tabs, long long tabs_ints, const uint16_t* lut, int canvas_h, int canvas_w,
                              int patch, int ld, void* patches, uint8_t* resized_u8, void* scratch, size_t scratch_bytes,
                              void* stream);

/* The same two integer passes with ANY caller-supplied tap tables and shifts:
 *   out = clip8((2^(p-1) + sum px * k) >> p), horizontal pass (precision_h) into a uint8 intermediate, then vertical.
 * With ATen's tables (gvl_resize_taps, weights widened to int32) this is torchvision's uint8 antialias resize for ANY
 * geometry — the general route of gvl_preprocess_u8's callers when an axis is up-scaled (crops and small images
 * through `encode_image`, e.g. scripts/realtime_inference.py:281-295), which the tuned K1 kernels do not cover.
 * Outputs (any subset): patches (bf16 via lut_bf16 [3,256]), resized_u8 [R, canvas_h, canvas_w, 3], f32_chw float
 * [R, 3, canvas_h, canvas_w] via lut_f32 [3,256] (HF `pixel_values`); zero outside a region's rectangle.  Windows that
 * would leave the crop are clipped in the kernel (the tables are device memory the host cannot validate). */
GVL_API int gvl_resize_two_pass_u8(const uint8_t* frame, int H, int W, int R, const int32_t* h_desc, const int32_t* tabs,
                           long long tabs_ints, int precision_h, int precision_v, const uint16_t* lut_bf16,
                           const float* lut_f32, int canvas_h, int canvas_w, int patch, int ld, void* patches,
                           uint8_t* resized_u8, float* f32_chw, void* scratch, size_t scratch_bytes, void* stream);

/* Position table for a gh x gw patch grid: pos bf16 [g*g, D] -> out bf16 [gh*gw, D], bicubic, align_corners = false,
 * A = -0.75, border-clamped taps, fp32 arithmetic.  Replaces `SiglipVisionEmbeddings.interpolate_pos_encoding`
 * (HF:models/siglip/modeling_siglip.py:137-174 -> ATen native/UpSampleBicubic2d, UpSample.h) — the
 * "interpolate_pos_encoding fixed" of SURVEY.md §8 f.4: without it the reference's region route raises for every
 * non-square region of a 729-position checkpoint.  gvl_siglip_forward then runs on a copy of the weight pack whose
 * T = gh*gw and pos = out. */
GVL_API int gvl_pos_interp_bicubic_bf16(const void* pos, int g, int D, int gh, int gw, void* out, void* stream);

/* out[b, :] = max over the T tokens of x[b] (bf16 [B,T,D]); out bf16 or float [B,D].
 * Replaces `sequence.max(dim=1)[0]` (src/perception/siglip_semantic_encoder.py:438-440). */
GVL_API int gvl_max_tokens_bf16(const void* x, int B, int T, int D, void* out, int out_f32, void* stream);

/* ---- K10: hardware video decode for the frame ingest --------------------------------- */
/* Replaces the decode half of `extract_frames` (scripts/extract_features.py:230-264: decord decodes EVERY frame of the
 * file to a host RGB array, the sampling rule then keeps every int(video_fps / fps)-th) with the GPU's NVDEC engines
 * (libnvcuvid.so.1, bound with dlopen — part of the GPU driver, not of this image's toolkit): the caller feeds
 * elementary-stream bytes (the host demuxes the container), pictures are decoded on the engine, and only the frames the
 * sampling rule keeps are converted NV12 -> packed RGB, directly into the caller's device batch buffer.
 * Codec ids are cuvid's: 4 = H.264, 8 = HEVC, 9 = VP8, 10 = VP9, 11 = AV1, 2 = MPEG-4, 5 = JPEG.  8-bit 4:2:0 only. */
typedef struct gvl_nvdec gvl_nvdec;
enum { GVL_CODEC_MPEG4 = 2, GVL_CODEC_H264 = 4, GVL_CODEC_JPEG = 5, GVL_CODEC_HEVC = 8, GVL_CODEC_VP9 = 10, GVL_CODEC_AV1 = 11 };
/* 1 if libnvcuvid loaded and exports every entry point used.  Host only, no device call. */
GVL_API int gvl_nvdec_available(void);
/* What this GPU's NVDEC decodes (cuvidGetDecoderCaps, 8-bit 4:2:0). */
GVL_API int gvl_nvdec_caps(int codec, int* supported, int* max_w, int* max_h, int* n_engines);
/* max_display_delay: pictures the parser may hold back to pipeline decode with display (0 = none; 2-4 recommended). */
GVL_API int gvl_nvdec_open(int codec, int max_display_delay, gvl_nvdec** out);
/* Frames are numbered from 0 in display order; frame i is kept iff i >= first_frame and (i - first_frame) % interval
 * == 0 (the reference's `range(0, total_frames, frame_interval)`, scripts/extract_features.py:247-253).
 * matrix_override: -1 = the stream's VUI matrix_coefficients (1 = BT.709, anything else BT.601), 1 / 6 force one;
 * full_range_override: -1 = the stream's video_full_range_flag, 0 / 1 force. */
GVL_API int gvl_nvdec_sampling(gvl_nvdec* d, long long first_frame, long long interval, int matrix_override,
                       int full_range_override);
/* data: HOST bytes of the elementary stream (Annex-B for H.264 / HEVC), any chunking.  out_rgb: device uint8
 * [cap, H, W, 3]: kept frames displayed during this call are written to slots 0..*frames_written-1 (conversion kernels on
 * `stream`, finished when the call returns); more than `cap` kept frames in one call is an error.  H x W must be the
 * video's display size.  *frames_displayed: running count of displayed frames.  end_of_stream flushes the parser. */
GVL_API int gvl_nvdec_feed(gvl_nvdec* d, const uint8_t* data, size_t bytes, int end_of_stream, uint8_t* out_rgb, int cap,
                   int H, int W, int* frames_written, long long* frames_displayed, void* stream);
/* info8: coded_w, coded_h, display_w, display_h, fps numerator, fps denominator, matrix_coefficients, full_range. */
GVL_API int gvl_nvdec_info(gvl_nvdec* d, int32_t* info8);
GVL_API int gvl_nvdec_close(gvl_nvdec* d);

#ifdef __cplusplus
}
#endif
#endif /* GVL_H_ */
