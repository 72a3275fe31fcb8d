/* libmudg_sm100.so -- C ABI of the B200-native MuDG sampler hot path.
 *
 * The reference (heiheishuang/MuDG) is pure Python/PyTorch and has no FFI of its own; its extension point is
 * utils/utils.py:27-42 (instantiate_from_config).  This ABI is what the drop-in Python classes
 * (lvdm.modules.networks.openaimodel3d.UNetModel, lvdm.models.autoencoder.AutoencoderKL,
 * lvdm.models.samplers.ddim.DDIMSampler) bind with ctypes; each entry point names the reference code it replaces.
 *
 * Conventions: plain C symbols, int return (0 = ok, negative = error, text via mudg_last_error(), thread local);
 * no exceptions / torch types cross the ABI; every tensor is a caller-owned, contiguous DEVICE pointer; the library
 * owns only packed fp16 weights + a workspace arena (freed by mudg_destroy); every call enqueues on the caller's
 * cudaStream_t (passed as void*) and never synchronises the host; one context per (process, device).
 */
#ifndef MUDG_H_
#define MUDG_H_
#include <stddef.h>
#include <stdint.h>
#if defined(__GNUC__)
#define MUDG_EXPORT __attribute__((visibility("default")))
#else
#define MUDG_EXPORT
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct MudgCtx MudgCtx;

enum { MUDG_F32 = 0, MUDG_F16 = 1, MUDG_U8 = 2 /* mudg_postdecode input only */ };
enum { MUDG_UNET = 0, MUDG_VAE = 1, MUDG_RESAMPLER = 2, MUDG_CLIP_IMAGE = 3, MUDG_CLIP_TEXT = 4 };

/* unet_config.params of configs/stage{1,2}-*_infer.yaml:26-56 (only the keys that shape the graph) */
typedef struct {
  int in_channels, out_channels, model_channels, num_res_blocks;
  int channel_mult[8], n_channel_mult;
  int attention_resolutions[8], n_attention_resolutions;
  int num_head_channels;   /* must be 64 */
  int context_dim;
  int init_attn_heads;     /* openaimodel3d.py:408 hard-codes 8 */
  int text_context_len;    /* attention.py:45: 77 */
} MudgUNetConfig;

/* first_stage_config.params.ddconfig (infer yaml :63-77) */
typedef struct {
  int ch;
  int ch_mult[8], n_ch_mult;
  int num_res_blocks, z_channels, out_ch, embed_dim;
} MudgVaeConfig;

MUDG_EXPORT const char* mudg_last_error(void);
MUDG_EXPORT int mudg_create(int device, const MudgUNetConfig* unet, const MudgVaeConfig* vae, MudgCtx** out);
MUDG_EXPORT void mudg_destroy(MudgCtx* ctx);

/* One state_dict entry (reference key names, e.g. "input_blocks.1.0.temopral_conv.conv1.2.weight").  Replaces
 * nn.Module.load_state_dict for UNetModel (openaimodel3d.py:281-565) / AutoencoderKL (autoencoder.py:27-32).
 * The tensor is converted to the kernels' fp16 layout immediately; the source may be freed after the call's
 * stream work completes.  `which` = MUDG_UNET | MUDG_VAE. */
MUDG_EXPORT int mudg_load_weight(MudgCtx* ctx, int which, const char* key, const void* dev_ptr, int dtype,
                                 const int64_t* shape, int ndim, void* stream);
/* Validates that every key the graph needs is present and builds the fused layouts (QKV, K|V, GEGLU interleave). */
MUDG_EXPORT int mudg_finalize_weights(MudgCtx* ctx, int which, void* stream);

/* Cross-attention context, constant over the DDIM steps: context [N, L, context_dim] (the `context` argument of
 * UNetModel.forward, openaimodel3d.py:567,580-587).  Precomputes to_k/to_v/to_k_ip/to_v_ip of all 16
 * SpatialTransformers (attention.py:89-94) once per clip.  T = frames; L == 77 + 16*T selects per-frame image tokens. */
MUDG_EXPORT int mudg_set_context(MudgCtx* ctx, const void* context, int dtype, int N, int L, int T, void* stream);

/* UNetModel.forward (openaimodel3d.py:567-628).  x [N, in_channels, T, h, w] fp32, t/c_label/fs [N] int64 (device),
 * out [N, out_channels, T, h, w] fp16 (the reference returns fp16 under autocast). */
MUDG_EXPORT int mudg_unet_forward(MudgCtx* ctx, const void* x, const int64_t* t, const int64_t* c_label,
                                  const int64_t* fs, int N, int T, int h, int w, void* out, void* stream);

/* The classifier-free-guidance form of the same forward: the two (ddim.py:221-222) or three (ddim_multiplecond.py:213-
 * 235) reference calls of one DDIM step differ only in `context`.  x / t / c_label / fs hold N / dup distinct samples tiled
 * dup times (sample n == sample n % (N / dup)) -- a promise of the caller -- and the context set by mudg_set_context has N
 * rows.  Everything before the first cross-attention (conv_in, init_attn, the first ResBlock, the first
 * SpatialTransformer's self-attention) is evaluated once per distinct sample; results equal mudg_unet_forward's. */
MUDG_EXPORT int mudg_unet_forward_shared(MudgCtx* ctx, const void* x, const int64_t* t, const int64_t* c_label,
                                         const int64_t* fs, int N, int dup, int T, int h, int w, void* out, void* stream);

/* DDIMSampler.p_sample_ddim after the UNet calls (ddim.py:226-277) + rescale_noise_cfg (utils_diffusion.py:147-158)
 * + predict_{eps,start}_from_z_and_v (ddpm3d.py:239-251).  x/noise/x_prev/pred_x0 fp32 [B, n]; v_* fp16 [B, n];
 * v_uncond may be NULL (no guidance).  Scalars are the per-step table entries. */
MUDG_EXPORT int mudg_ddim_step(const void* x, const void* v_cond, const void* v_uncond, const void* noise,
                               void* x_prev, void* pred_x0, int B, int64_t n, float cfg_scale, float guidance_rescale,
                               float sqrt_alphas_cumprod_t, float sqrt_one_minus_alphas_cumprod_t, float rescale,
                               float a_prev, float sigma_t, void* stream);

/* AutoencoderKL.decode (autoencoder.py:104-107) + Decoder.forward (ae_modules.py:539-578), frame by frame like
 * decode_core with perframe_ae (ddpm3d.py:646-667).  z [F, z_channels, h, w] fp32 ALREADY divided by scale_factor;
 * out [F, out_ch, 8h, 8w] fp16. */
MUDG_EXPORT int mudg_vae_decode(MudgCtx* ctx, const void* z, int F, int h, int w, void* out, void* stream);

/* "Next" row: AutoencoderKL.encode (autoencoder.py:97-102) + Encoder.forward (ae_modules.py:432-463), the
 * get_latent_z step before the sampler (virtual_pose_render.py:54-59).  x [F, 3, H, W] fp32 in [-1,1]; moments
 * [F, 2*z_channels, H/8, W/8] fp32 (mean | logvar) -- sampling the posterior stays on the host side (CPU RNG parity). */
MUDG_EXPORT int mudg_vae_encode(MudgCtx* ctx, const void* x, int F, int H, int W, void* moments, void* stream);

/* "Next" row: the per-frame CPU code the driver runs on the decoded clip before writing files
 * (virtual_pose_render.py:243 clamp; eval_tools.py:22-27 uint8 conversion; :70-74 + colormap :205-236 depth mean and
 * Spectral colouring; visualize_semantic :297-347 nearest-of-19-palette class + colour).  frames [B, 3, T, H, W]
 * (decode_first_stage layout, dtype MUDG_F16 | MUDG_F32, or MUDG_U8 for frames that are already uint8); modes: HOST array [B] of MUDG_POST_{COLOR,DEPTH,SEMANTIC}
 * (class labels 0 / 500 / 1 of the driver).  rgb_u8 [B, T, 3, H, W] uint8 (the frame, or its depth / semantic
 * visualisation); depth_f32 [B, T, H, W] fp32 in [0,1] and class_u8 [B, T, H, W] uint8 are written for the samples of
 * that mode only and may be NULL.  Bit-exact against the reference's CPU arithmetic. */
enum { MUDG_POST_COLOR = 0, MUDG_POST_DEPTH = 1, MUDG_POST_SEMANTIC = 2 };
MUDG_EXPORT int mudg_postdecode(const void* frames, int dtype, int B, int T, int H, int W, const int* modes,
                                void* rgb_u8, void* depth_f32, void* class_u8, void* stream);

/* "Next" row: Resampler.forward (lvdm/modules/encoders/resampler.py:131-144, with
 * PerceiverAttention :48-101 and FeedForward :31-37) -- the once-per-clip projection of the image-encoder tokens to the
 * UNet's image context.  Weights: state_dict keys of the reference module loaded with which = MUDG_RESAMPLER
 * ("latents" as [num_queries*video_length, dim]); dimensions are derived from the shapes at mudg_finalize_weights.
 * x [B, L, embedding_dim] (MUDG_F32 | MUDG_F16) -> out [B, num_queries*video_length, output_dim] fp32. */
MUDG_EXPORT int mudg_resampler_forward(MudgCtx* ctx, const void* x, int dtype, int B, int L, void* out, void* stream);

/* "Next" row: the OpenCLIP ViT-H/14 towers, once per clip.  The reference calls open_clip for
 * them (not vendored); weights are loaded under open_clip's state-dict names relative to `model.visual.` (which =
 * MUDG_CLIP_IMAGE: conv1.weight, class_embedding, positional_embedding, ln_pre.*, transformer.resblocks.N.{ln_1,ln_2}.*,
 * .attn.in_proj_{weight,bias}, .attn.out_proj.*, .mlp.{c_fc,c_proj}.*) and to `model.` (which = MUDG_CLIP_TEXT:
 * token_embedding.weight, positional_embedding, transformer.resblocks.N.*, ln_final.*); widths, depth, patch and grid
 * are derived from the shapes at mudg_finalize_weights, the head count is an argument (16 for ViT-H/14).
 *
 * mudg_clip_image_forward = FrozenOpenCLIPImageEmbedderV2.forward (lvdm/modules/encoders/condition.py:334-372):
 *   img [B, 3, H, W] (MUDG_F32 | MUDG_F16) -> out [B, 1 + grid^2, width] fp32 token-level transformer output (no
 *   ln_post / proj).  resize != 0: img is in [-1, 1] at any H x W and first goes through `preprocess` (:318-326: kornia
 *   bicubic resize to the tower size with align_corners and the anti-alias gaussian, (x + 1) / 2, CLIP mean / std);
 *   resize == 0: img is the already normalised tower input (H = W = grid * patch).
 * mudg_clip_text_forward = FrozenOpenCLIPEmbedder.encode_with_transformer (:214-232): tokens [B, L <= 77] int64 (device)
 *   -> out [B, L, width] fp32 = ln_final(blocks[0 .. layers - skip_last) (token_embedding + positional_embedding)) under
 *   the causal mask; skip_last is the reference's layer_idx (0 "last", 1 "penultimate").  Tokenisation stays on the host. */
MUDG_EXPORT int mudg_clip_image_forward(MudgCtx* ctx, const void* img, int dtype, int B, int H, int W, int resize, int heads,
                                        void* out, void* stream);
MUDG_EXPORT int mudg_clip_text_forward(MudgCtx* ctx, const int64_t* tokens, int B, int L, int heads, int skip_last, void* out,
                                       void* stream);

/* colormap(image, cmap="Spectral", bytes=True) of the reference (eval_tools.py:137-250, method_custom): map fp32 [n] in
 * [0,1] (clamped) -> out_u8 [n, 3] (HWC).  Used for the ground-truth depth visualisation (eval_tools.py:82). */
MUDG_EXPORT int mudg_colormap_spectral(const void* map_f32, int64_t n, void* out_u8, void* stream);

/* Workspace the library needs (and will allocate on first use) for a forward of this shape. */
MUDG_EXPORT size_t mudg_workspace_bytes(MudgCtx* ctx, int N, int T, int h, int w);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
MUDG_EXPORT int64_t mudg_launch_count(MudgCtx* ctx);

/* Per-launch profiler (bench.py's roofline leg): while enabled, every kernel family of the path is bracketed by a
 * CUDA-event pair on its launching stream (forwards run eagerly instead of replaying their CUDA graph).
 * mudg_profile_report(NULL, 0) synchronises, builds the report and returns its size; a second call with a buffer copies
 * it out.  CSV rows: family,shape,launches,ms,flops,bytes (algorithmic FLOPs / HBM bytes of the launches, summed). */
MUDG_EXPORT int mudg_profile(int enable);
MUDG_EXPORT size_t mudg_profile_report(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
