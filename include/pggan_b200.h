/* pggan_b200 — C ABI of the B200-native PGGAN training-step kernels (libpggan_b200.so).
 *
 * Drop-in boundary: the reference has no FFI; its hot path is reached through
 * Python classes (lib/layers.py, lib/blocks.py, pggan/nets.py, pggan/loss.py).  The Python
 * package `pggan_b200` mirrors those classes and calls ONLY the entry points below, on raw
 * device pointers owned by the caller (torch tensors).  Every entry point
 *   - launches asynchronously on the given CUDA stream (cudaStream_t passed as void*),
 *   - never allocates, frees or synchronises (workspaces are passed in),
 *   - returns PG_OK or a negative error code; pggan_last_error() gives the message.
 *
 * Layouts.  Feature maps are NHWC ("channels_last"): [B][H][W][C], dtype PG_BF16, PG_F16 or PG_F32
 * (PG_F32 operands run on kind::tf32 tensor-core MMAs; the Python layer feeds them as 3xTF32 splits).
 * RGB images, logits, latents, parameters and parameter gradients are fp32 in the reference's
 * own layouts (NCHW images, OIHW conv weights, [out,in] linear weights).
 *
 * Each entry cites the reference code it replaces (file:line in yukyeongleee/pggan).
 */
#ifndef PGGAN_B200_H
#define PGGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_ARG (-1)
#define PG_ERR_CUDA (-2)
#define PG_ERR_UNSUPPORTED (-3)

#define PG_BF16 0
#define PG_F32 1
#define PG_F32_RAW 2 /* output-only: fp32 result from any operand dtype (final logits / images) */
#define PG_F16 3     /* IEEE half operands / feature maps: same kernels and speed as PG_BF16, 11-bit significand
                      * (meets the 1e-2 forward bar that bf16's 8 bits miss beyond 16x16), 5-bit exponent */

/* conv backend selection (bring-up / debugging): 0 = tcgen05 wherever the shape allows (default),
 * 1 = force the SIMT direct kernels. */
#define PG_BACKEND_AUTO 0
#define PG_BACKEND_SIMT 1

int pggan_version(void);
const char* pggan_last_error(void);
int pggan_set_conv_backend(int backend);
/* number of tcgen05 / SIMT GEMM-class launches since load (evidence for "which path ran") */
int pggan_launch_counts(int64_t* tcgen05_launches, int64_t* simt_launches, int64_t* other_launches);
/* sizeof() of the parameter structs as this library was compiled — a binding checks its own struct layout against
 * these before the first call (a short struct would make the library read past the caller's buffer). */
int pggan_sizeof_conv_params(void);
int pggan_sizeof_wgrad_params(void);

/* ------------------------------------------------------------------------------------------
 * Weight packing.  The tensor-core kernels want K-major operands [n][t][c]; parameters stay in the
 * reference's layouts (OIHW conv, [out,in] linear).  A pg_pack_map describes where logical element
 * (n,t,c) lives in the fp32 parameter:
 *     off = base + n1*s_n1 + n2*s_n2 + t*s_t + c1*s_c1 + c2*s_c2,   n = n1*n2_len + n2,  c = c1*c2_len + c2
 * (strides in elements; s_t < 0 gives the 180-degree tap flip of dgrad; the two-level n / c indices
 * express the NCHW<->NHWC flatten permutations of the two big linear layers; rows >= n_valid and
 * columns >= c_valid are zero padding, e.g. the 513 -> 576 channel pad of the minibatch-stddev conv).
 * Replaces: the implicit weight use of nn.Conv2d / nn.Linear in ConstrainedLayer.forward
 * (lib/layers.py:58-63).
 */
typedef struct {
    int32_t n_total, n_valid, n2, taps, c_total, c_valid, c2;
    int64_t base, s_n1, s_n2, s_t, s_c1, s_c2;
} pg_pack_map;
/* out[n][t][c] (dtype) = in[off(n,t,c)] or 0 in the padding */
int pggan_pack_weight(const float* in, void* out, int dtype, const pg_pack_map* map, void* stream);
/* gradients back: out_fp32[off(n,t,c)] (+)= scale * sum_s in_fp32[s][n][t][c] for the valid (n,c); `in` holds
 * the `slices` dense [n_total][taps][c_total] fp32 partial-sum slices written by pggan_conv_wgrad.
 * accumulate: 0 = overwrite, 1 = add to what `out` holds (gradient accumulation straight into .grad). */
int pggan_unpack_wgrad(const float* in, float* out, const pg_pack_map* map, int slices, float scale, int accumulate,
                       void* stream);

/* out (to_dtype) = in (from_dtype), elementwise, between PG_BF16 / PG_F16 / PG_F32 (conversions to IEEE half saturate).
 * The mixed-precision mode keeps activations in fp16 and gradient tensors in bf16; tcgen05.mma kind::f16 needs BOTH
 * operands in one format (a bf16 x fp16 instruction descriptor raises an illegal-instruction fault — measured,
 * tools/probe_mixed.py), so the activation operand of a weight-gradient GEMM is re-rounded to bf16 by this kernel. */
int pggan_convert(const void* in, void* out, int64_t n, int from_dtype, int to_dtype, void* stream);

/* fp32 mode = 3xTF32.  The tensor core truncates fp32 operands to tf32 (hi); this writes three copies
 * of in [m][c], one of them (slot lo_slot) holding lo = x - trunc_tf32(x), concatenated along channels
 * ([m][3c]) or along rows ([3m][c]).  A GEMM over (x,x,x_lo) x (w,w_lo,w) then yields
 * hi*hi + hi*lo + lo*hi, i.e. fp32-class products, with the unchanged tensor-core kernels. */
int pggan_split3(const float* in, float* out, int64_t m, int c, int along_channels, int lo_slot, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear, forward-style:
 *   y[b,h,w,n] = post( scale * ( sum_{t,c} x[b, h+dh(t), w+dw(t), c] * wp[n][t][c] + bias[n] ) )
 *   taps = 9: 3x3, pad 1, (dh,dw) = (t/3-1, t%3-1);  taps = 1: 1x1 / linear (H=W=1).
 *   post(v) = lrelu(v, slope) then, if mask_src != NULL, v * (mask_src[b,h,w,n] > 0 ? 1 : mask_slope).
 * Serves fprop, dgrad (flipped/transposed wp) and the grad-of-dgrad fprop.
 * Replaces: EqualizedConv2d / EqualizedLinear forward = nn.Conv2d|nn.Linear + `x *= self.weight`
 * (lib/layers.py:58-63) + the following nn.LeakyReLU (lib/blocks.py:127,137,185,192,254) and, for
 * dgrad, autograd's convolution_backward + leaky_relu_backward.
 * Tensor-core path (tcgen05 + TMA) needs: H,W powers of two, C*elsize % 128 == 0, N % 16 == 0;
 * anything else runs the SIMT direct kernel.
 */
typedef struct {
    const void* x;        /* [B,H,W,C]  dtype */
    const void* wp;       /* [N][taps][C] dtype (pggan_pack_weight) */
    const float* bias;    /* [N] fp32 or NULL */
    const void* mask_src; /* [B,H,W,N] dtype or NULL */
    void* y;              /* [B,H,W,N] out_dtype */
    int32_t B, H, W, C, N, taps;
    int32_t dtype, out_dtype;
    float scale, slope, mask_slope;
    /* PG_F32 only — 3xTF32 workspaces the tensor-core path fills with pggan_split3:
     * ws_x: 3*B*H*W*C floats, ws_w: 3*N*taps*C floats (ignored by the SIMT path / PG_BF16). */
    void* ws_x;
    void* ws_w;
    /* Optional fused pixel norm (PixelwiseVectorNorm, lib/layers.py:13-14, after the LeakyReLU): if pn_rnorm != NULL,
     * y = post(..) * rsqrt(mean_n(post(..)^2) + pn_eps) and pn_rnorm[b,h,w] (fp32) receives the factor.  Available where
     * one epilogue thread owns all N channels of a pixel (pggan_conv_pn_fusable(p) != 0); PG_ERR_UNSUPPORTED elsewhere. */
    float* pn_rnorm;
    float pn_eps;
    /* Fused nearest-2x up-sampling (upscale2d + conv of ProgressiveGeneratorBlock, lib/blocks.py:145-151 with
     * lib/utils.py:106-118): if up2 != 0, x is the LOW-resolution tensor [B, H/2, W/2, C] and the convolution reads
     * up2(x) — H, W stay the OUTPUT dims.  Output tiles are formed per pixel parity (h%2, w%2): for a fixed parity the
     * tap (dh,dw) of every pixel of the tile reads low-res pixel (h/2 + floor((h%2+dh)/2), w/2 + floor((w%2+dw)/2)), a
     * plain shifted TMA box of x, so the up-sampled tensor is never written to HBM.  Available where
     * pggan_conv_up2_fusable(p) != 0 (tensor-core plain kernel); PG_ERR_UNSUPPORTED elsewhere. */
    int32_t up2;
    /* Optional split-K workspace (any dtype; NULL: never split).  A layer with few output tiles and a long K — the
     * 4x4 / 8x8 trunk layers — is cut into several K ranges per output tile so that idle SMs share the operand traffic;
     * the partial accumulators are added (fp32 red.global) into per-tile pair buffers in ws_k — at most two addends per
     * buffer, so a buffer's content does not depend on arrival order — and the CTA that arrives last at a tile adds the
     * buffers in a fixed order and runs the epilogue: results are reproducible bit for bit.  ws_k: 16-byte aligned,
     * ws_k_bytes >= 4096, ALL ZERO before the first call; every call leaves it all zero again; calls that share a
     * workspace must be stream-ordered.  The library decides per shape whether (and how far) to split within
     * ws_k_bytes. */
    void* ws_k;
    int64_t ws_k_bytes;
} pg_conv_params;
int pggan_conv_fprop(const pg_conv_params* p, void* stream);
/* which kernel pggan_conv_fprop would run this shape on (pointers may be NULL): lets the host decide where fusing a
 * LeakyReLU backward into the dgrad epilogue (mask_src) is free (PLAIN / ROWS: one output row per thread) and where
 * it is not (SWAPPED: transposing, instruction-bound epilogue). */
#define PG_CONV_KIND_SIMT 0
#define PG_CONV_KIND_PLAIN 1
#define PG_CONV_KIND_SWAPPED 2
#define PG_CONV_KIND_ROWS 3
int pggan_conv_kernel_kind(const pg_conv_params* p);
int pggan_conv_pn_fusable(const pg_conv_params* p); /* 1 if the call may carry pn_rnorm */
int pggan_conv_up2_fusable(const pg_conv_params* p); /* 1 if the call may carry up2 (dims as for the fused call) */

/* Weight-gradient GEMM:  sum_s acc[s][n][t][c] = sum_{b,h,w} g[b,h,w,n] * x[b, h+dh(t), w+dw(t), c]
 * acc holds pggan_conv_wgrad_slices(p) dense fp32 [N][taps][C] slices (split-K over pixels writes one slice per
 * split with plain stores); the call defines every element, nothing needs pre-zeroing.
 * Also the grad-of-dgrad w.r.t. W with operands (g_pre, ggx).
 * Replaces: autograd convolution_backward (weight) / addmm backward of lib/layers.py:60.
 */
typedef struct {
    const void* g;   /* [B,H,W,N] dtype */
    const void* x;   /* [B,H,W,C] dtype */
    float* acc;      /* [N][taps][C] fp32 */
    int32_t B, H, W, C, N, taps;
    int32_t dtype;
    /* PG_F32 only — 3xTF32 workspaces: ws_g: 3*B*H*W*N floats, ws_x: 3*B*H*W*C floats */
    void* ws_g;
    void* ws_x;
    /* Optional direct accumulation (NULL: off).  If `direct` != NULL and pggan_conv_wgrad_direct_ok(p) != 0, every
     * unit of the launch ADDS direct_scale * (its partial sums) straight into direct[N][taps][C] (fp32, red.global) —
     * autograd's accumulation into a .grad that has the operand's own layout (a channels_last conv weight, a plain
     * [out,in] linear weight) — and `acc` is not touched: no slice buffer, no pggan_unpack_wgrad pass. */
    float* direct;
    float direct_scale;
    int32_t reserved_;
} pg_wgrad_params;
int pggan_conv_wgrad(const pg_wgrad_params* p, void* stream);
int pggan_conv_wgrad_slices(const pg_wgrad_params* p);
/* 1 if pggan_conv_wgrad honours p->direct for this shape (the tensor-core kernels), 0 if the caller must use acc +
 * pggan_unpack_wgrad (SIMT fallback shapes) */
int pggan_conv_wgrad_direct_ok(const pg_wgrad_params* p); /* how many slices the call above will write (>= 1) */

/* ------------------------------------------------------------------------------------------
 * Elementwise / reduction kernels on NHWC feature maps (n = element count, m = pixels).
 */
/* A coefficient that may depend on the fade-in alpha (Generator.alpha / Discriminator.alpha, pggan/model.py:146-147,
 * 179-182): value = c0 + c1 * (*alpha), or c0 when alpha == NULL.  `alpha` points to ONE fp32 in device memory that the
 * kernel reads when it runs, so a captured CUDA graph follows change_alpha() without being captured again. */
typedef struct {
    float c0, c1;
    const float* alpha;
} pg_coef;
/* out = g * (y > 0 ? 1 : slope)      — leaky_relu_backward, lib/blocks.py:127,137,185,192 */
int pggan_lrelu_mask_mul(const void* g, const void* y, void* out, int64_t n, float slope, int dtype, void* stream);
/* out[b,2h+i,2w+j,c] = scale * x[b,h,w,c]   — upscale2d, lib/utils.py:106-118 (exact) and avg-pool backward */
int pggan_upsample2x(const void* x, void* out, int B, int H, int W, int C, pg_coef scale, int dtype, void* stream);
/* the same fused with a LeakyReLU backward: out = up2(scale*x) * (mask_src > 0 ? 1 : slope), mask_src [B,2H,2W,C]
 * — backward of "conv -> LeakyReLU -> AvgPool2d" (lib/blocks.py:192-193) in one pass instead of
 * avg_pool2d_backward + leaky_relu_backward */
int pggan_upsample2x_mask(const void* x, const void* mask_src, void* out, int B, int H, int W, int C, pg_coef scale,
                          float slope, int dtype, void* stream);
/* ... and with the bias gradient of that conv in the same pass: colsum[c] += cs_scale * sum over output pixels of
 * out[.,c] (autograd's bias gradient of lib/layers.py:62 for the second conv of a D block); C/VN must divide 256 */
int pggan_upsample2x_mask_colsum(const void* x, const void* mask_src, void* out, float* colsum, int B, int H, int W,
                                 int C, pg_coef scale, float slope, float cs_scale, int dtype, void* stream);
/* out[b,h,w,c] = scale * sum_{i,j<2} x[b,2h+i,2w+j,c]; if blend_with != NULL:
 * out = (1-alpha)*blend_with + alpha*that   — nn.AvgPool2d(2) lib/blocks.py:193 (+ fade pggan/nets.py:265) */
int pggan_pool2x2(const void* x, const void* blend_with, void* out, int B, int H, int W, int C, pg_coef scale,
                  pg_coef alpha, int dtype, void* stream);
/* out = sa*a + sb*b (b may be NULL)  — fade-in blends pggan/nets.py:156,265 and their backward */
int pggan_axpby(const void* a, const void* b, void* out, int64_t n, pg_coef sa, pg_coef sb, int dtype, void* stream);
/* y = x * rsqrt(mean_c(x^2) + eps), rnorm[m] saved   — PixelwiseVectorNorm, lib/layers.py:13-14 */
int pggan_pixelnorm_fwd(const void* x, void* y, float* rnorm, int64_t m, int C, float eps, int dtype, void* stream);
/* gx = rnorm * (gy - y * mean_c(gy*y)); if act_slope != 1 the LeakyReLU mask (sign of y) is applied too;
 * if colsum != NULL also colsum[c] += cs_scale * sum_m gx[m,c] (the bias gradient, same pass).
 * dtype: format of gy / gx; y_dtype: format of the saved activation y (mixed precision: bf16 gradients, fp16 y) */
int pggan_pixelnorm_bwd(const void* gy, const void* y, const float* rnorm, void* gx, int64_t m, int C,
                        float act_slope, float* colsum, float cs_scale, int dtype, int y_dtype, void* stream);
/* fused LeakyReLU backward + bias gradient: out = g * (y > 0 ? 1 : slope) (y NULL: out = g; out NULL: no
 * store) and colsum[c] += scale * sum_m out[m,c] (colsum NULL: none).  C/vec must divide 256. */
int pggan_mask_colsum(const void* g, const void* y, void* out, float* colsum, int64_t m, int C, float slope,
                      float scale, int dtype, void* stream);
/* out[c] (+)= scale * sum_m g[m,c]   — bias gradient of lib/layers.py:60 (bias is He-scaled too) */
int pggan_colsum(const void* g, float* out, int64_t m, int C, float scale, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * 1x1 RGB convolutions (bandwidth kernels; N=3 / K=3 is not tensor-core shaped).
 */
/* out[b,h,w,c] = lrelu(scale*(sum_j w[c][j]*p_j + bias[c]), slope), p = img[b,:,h,w] or, if pool,
 * the 2x2 mean of img[b,:,2h..,2w..].  img fp32 NCHW [B,cimg,Hi,Wi] (cimg <= 4), w fp32 [C][cimg],
 * out NHWC [B,H,W,C].
 * — fromRGBBlock lib/blocks.py:285-292 (+ downscale2d lib/utils.py:120-124); also toRGB dgrad. */
int pggan_rgb_expand(const float* img, const float* w, const float* bias, void* out, int B, int H, int W, int C,
                     int cimg, int pool, float scale, float slope, int dtype, void* stream);
/* out[b,j,h,w] = scale*(sum_c w[j][c]*x[b,h,w,c] + bias[j]) fp32 NCHW, w fp32 [cimg][C]; if up, every
 * value (times up_scale) is written to the 2x2 block of a [B,cimg,2H,2W] output
 * — toRGBBlock lib/blocks.py:164-170; also fromRGB dgrad (+ avg-pool backward). */
int pggan_rgb_reduce(const void* x, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                     int cimg, int up, float scale, float up_scale, int dtype, void* stream);
/* acc[c][j] += scale * sum_{b,h,w} f[b,h,w,c] * p_j(b,h,w)   (p as in rgb_expand) — 1x1 weight grads;
 * if bsum != NULL also bsum[c] += scale * sum f[b,h,w,c] (the bias gradient, same pass). */
int pggan_rgb_wgrad(const void* f, const float* img, float* acc, float* bsum, int B, int H, int W, int C, int cimg,
                    int pool, float scale, int dtype, void* stream);
/* fp32 NCHW image fade.  low == NULL: out = (1-a)*up2(avgpool2(x)) + a*x  (pggan/model.py:217-221)
 *                        low != NULL: out = (1-a)*up2(low) + a*x          (pggan/nets.py:155-156)   */
int pggan_image_fade(const float* x, const float* low, float* out, int B, int Cimg, int H, int W, pg_coef alpha,
                     void* stream);
/* out[b,c,h,w] = scale * sum_{2x2} x   (fp32 NCHW) — backward of the RGB up-sampling */
int pggan_image_pool_sum(const float* x, float* out, int B, int Cimg, int H, int W, pg_coef scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * GPU input pipeline: the transform chain of UnsupervisedDataset (lib/dataset.py:100-111) and the
 * real-image fade of train_step (pggan/model.py:217-221) on decoded image batches.  Arithmetic = torchvision's
 * float-tensor backend (same formulas and operation order, fp32; the PIL backend the reference uses quantises to uint8
 * after every operation).
 */
/* anti-aliased bilinear resize (transforms.Resize((Ho,Wo)) — the triangle filter of half-width max(scale,1) that PIL /
 * torchvision apply): src [planes][Hs][Ws] uint8 (src_is_u8) or fp32, times in_mul (1/255 for uint8 -> [0,1]);
 * tmp: planes*Hs*Wo floats; dst [planes][Ho][Wo] fp32 */
int pggan_resize_aa(const void* src, int src_is_u8, float* tmp, float* dst, int planes, int Hs, int Ws, int Ho, int Wo,
                    float in_mul, void* stream);
/* out [B,3,S,S] = fade( Normalize(0.5,0.5)( hflip?( ColorJitter(img) ) ) ), img [B,3,S,S] fp32 in [0,1].
 * iparams [B][5] = {flip, op0..op3} (op: 0 brightness, 1 contrast, 2 saturation, 3 hue — the sample's random order,
 * transforms.ColorJitter.forward), fparams [B][4] = the four factors; mean_ws: B floats of workspace (the grey mean
 * contrast blends with).  fade != 0: out = (1-alpha)*up2(avgpool2(.)) + alpha*(.)  (pggan/model.py:217-221). */
int pggan_jitter_normalize_fade(const float* img, const int32_t* iparams, const float* fparams, float* mean_ws, float* out,
                                int B, int S, int fade, pg_coef alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * Minibatch standard deviation (lib/blocks.py:204-233): groups of `group` consecutive samples,
 * unbiased variance, eps inside sqrt, mean over (C,H,W).  x is [B][K] (K = H*W*C, any order).
 */
int pggan_mbstd_fwd(const void* x, float* stat, int B, int64_t K, int group, int dtype, void* stream);
/* gx[i,k] = gs[g] * (x-mu)/(K (n-1) sigma).  dtype: format of x; g_dtype: format of the gradient tensors (gx here;
 * u and gx2 below) — they differ in the mixed-precision mode */
int pggan_mbstd_bwd(const void* x, const float* gs, void* gx, int B, int64_t K, int group, int dtype, int g_dtype,
                    void* stream);
/* double backward of the above w.r.t. cotangent u (= ggx): ggs[g] and gx2[i,k] */
int pggan_mbstd_bwd2(const void* x, const void* u, const float* gs, float* ggs, void* gx2, int B, int64_t K,
                     int group, int dtype, int g_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * WGAN-GP / R1 reductions over fp32 image gradients g [B][n]  (pggan/loss.py:85-87, lib/loss.py:132-134)
 */
/* sumsq[b] = sum_k g[b,k]^2   (caller zeroes sumsq) */
int pggan_rows_sumsq(const float* g, float* sumsq, int B, int64_t n, void* stream);
/* out[b,k] = coef[b] * g[b,k] */
int pggan_rows_scale(const float* g, const float* coef, float* out, int B, int64_t n, void* stream);
/* fused epilogue on [B] scalars: res[0] = w_gp*sum_b (sqrt(sumsq_b)-1)^2, res[1] = w_drift*sum_b pred_b^2,
 * coef[b] = w_gp*2*(norm_b-1)/norm_b  (the backward row scale). pred may be NULL. */
int pggan_gp_finalize(const float* sumsq, const float* pred, float* res, float* coef, int B, float w_gp,
                      float w_drift, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam (torch.optim.Adam semantics, lib/model.py:95-97, lib/utils.py:72-75).
 * table: device array of n_tensors entries {param*, grad*, m*, v*, numel}; step: device int64 counter
 * (incremented by the kernel: CUDA-graph friendly).  grad_scale folds the 1/world_size of the DP mean.
 */
typedef struct {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t numel;
    /* optional 16-bit copy of the updated parameter, same element order (NULL: none), written in the same pass: the
     * tensor cores' packed fprop operand of a conv weight stored [O][kh][kw][I] (channels_last) IS that copy, so the
     * fp32 -> fp16/bf16 re-pack of lib/model.py:95-97's update costs no extra pass */
    void* shadow;
    int32_t shadow_dtype; /* PG_BF16 / PG_F16 */
    int32_t pad_;
} pg_adam_entry;
int pggan_adam_chunk_size(void); /* elements handled per chunk-table entry */
int pggan_adam_step(const pg_adam_entry* table, int n_tensors, int64_t total_chunks, const int32_t* chunk_tensor,
                    const int64_t* chunk_offset, int64_t* step, float lr, float beta1, float beta2, float eps,
                    float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGGAN_B200_H */
