/* b2r.h — C-ABI of libb2r.so, the sm_100a (B200) kernels behind the degrade -> restore -> classify path.
 *
 * The reference (LordTARN1SHED/Image-Restoration-for-Road-Sign-Recognition-in-Autonomous-Driving) has no FFI of
 * its own: its boundary for this path is the PyTorch module contract (`nn.Module.forward`, `load_state_dict`)
 * and a handful of NumPy/OpenCV functions.  Each entry point below names the reference call site it replaces
 * (file:line relative to the reference root).  The Python host in
 * `image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/` binds these with ctypes and keeps
 * the reference's module surface on top (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its comment says "host".
 *   - the library never allocates or frees device memory and never keeps a pointer after a call returns.
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*).
 *   - return value: 0 = ok, negative = error (B2R_E*); `b2r_last_error()` gives the message (thread-local).
 *   - activations between layers: NHWC bf16.  There is no CPU fallback anywhere in this library.
 */
#ifndef B2R_H_
#define B2R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_VERSION 200 /* bumped on every struct / signature change; the ctypes binding refuses any other value */

#define B2R_OK 0
#define B2R_EINVAL (-22)   /* bad argument (shape, alignment, null pointer) */
#define B2R_ECUDA (-5)     /* a CUDA runtime / driver call failed */
#define B2R_ENODEV (-19)   /* no sm_100 device */

/* activation codes */
#define B2R_ACT_NONE 0
#define B2R_ACT_RELU 1
#define B2R_ACT_PRELU 2 /* single shared slope, nn.PReLU() default (14_train_unified_advanced.py:101) */

int b2r_version(void);
/* sizeof of an ABI struct as THIS build sees it (which = 0: b2r_conv_gemm_desc, 1: b2r_tensor), so a binding can check its
 * own layout against the library it loaded instead of trusting the version number alone; -1 for an unknown `which`. */
int b2r_abi_sizeof(int which);
const char* b2r_last_error(void);
/* Diagnostic: name of the kernel the last b2r_conv_gemm call of this thread launched ("conv_gemm_pair_kernel<256>",
 * "conv_w3_kernel<head>", ...); bench.py uses it to split its per-launch timings by kernel. */
const char* b2r_last_conv_kernel(void);

/* ---------------------------------------------------------------------------------------------------------------
 * (1) Fused compound degradation: motion blur (+) fog (+) AWGN, u8 NHWC in -> u8 NHWC out, one launch.
 *
 * Replaces apply_compound_distortion (16_gen_compound_data.py:14-37, order Blur->Fog->Noise),
 * apply_random_distortions (14_train_unified_advanced.py:31-64, order Fog->Noise->Blur),
 * make_compound_distortion (15_test_unified.py:93-120) and the single degradations
 * add_gaussian_noise (02_gen_noise.py:12-27), apply_motion_blur (03_gen_blur.py:11-30, without the min-max
 * renormalisation) and add_fog (04_gen_fog.py:12-31).
 *
 * Per-image parameters (device arrays of length N):
 *   ksize[i]   blur kernel side d (0 or 1 = no blur, else 2..B2R_MAX_BLUR = 15; a larger value makes the kernel trap: the
 *              launch fails with a CUDA error instead of overrunning its tap tables); taps[i*225 + ky*d + kx] row-major d x d f32,
 *              correlation with anchor d/2 and BORDER_REFLECT_101, f32 accumulation over the non-zero taps in
 *              row-major order, round-half-even + saturate to u8 (cv2.filter2D on u8).
 *   fog_on[i]  0 skips the fog stage; fog_t[i] = transmission t; fog_add[i] = float32(A*(1-t)), the airlight term
 *              evaluated by the host in double precision exactly as Python evaluates `A * (1 - t)`:
 *              I = fl32(fl32(J*t) + fog_add).
 *   sigma[i]   AWGN standard deviation on the [0,1] scale; 0 skips the noise stage.
 * order: B2R_ORDER_BLUR_FOG_NOISE = chain(blur(in)) (script 16); B2R_ORDER_FOG_NOISE_BLUR = blur(chain(in))
 *        (scripts 14/15), with chain(v) = quant(noise(fog(v/255))).
 * flags: B2R_DEG_CLIP_AFTER_NOISE clips to [0,1] right after the noise is added (15:108).  Every float -> u8 step
 *        is clip(x*255, 0, 255) followed by truncation, as in the reference.
 * in == out is allowed only when nothing blurs (ksize == NULL): the blur reads a halo that neighbouring CTAs overwrite;
 *        overlapping in / out ranges with ksize != NULL return B2R_EINVAL.
 * Noise: if `noise` is non-null it is an f64 NHWC tensor of the noise values themselves (what
 *        np.random.normal(0, sigma, shape) returned), added in float64 like NumPy does: the parity path ("AWGN is
 *        compared by injecting the same noise tensor").  Otherwise Philox4x32-10 keyed by `seed`, noise stream 2 (ABI 200):
 *        the 12 values of 4 consecutive pixels of a row take the four Box-Muller normals of each of three calls with
 *        counter = (y * ceil(W / 4) + x / 4, image_index0 + i (64 bit), 4 + k), scaled by sigma[i]; a pure function of
 *        (seed, global image index, y, x, channel, W): independent of batch split, launch geometry and world size
 *        (oracle/degrade_oracle.py::philox_normals_v2).
 * ------------------------------------------------------------------------------------------------------------- */
#define B2R_ORDER_BLUR_FOG_NOISE 0
#define B2R_ORDER_FOG_NOISE_BLUR 1
#define B2R_DEG_CLIP_AFTER_NOISE 1
#define B2R_MAX_BLUR 15

int b2r_degrade(const uint8_t* in_nhwc, uint8_t* out_nhwc, int N, int H, int W,
                const float* taps /* [N,225] or NULL when no image blurs */, const int32_t* ksize,
                const float* fog_t, const float* fog_add, const int32_t* fog_on, const float* sigma,
                const double* noise /* [N,H,W,3] or NULL */, uint64_t seed, uint64_t image_index0, int order,
                int flags, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (2) First conv layer (C_in = 3), bias + activation fused, output NHWC bf16 [N,H,W,64].
 *
 * Replaces SimpleUNet.enc1[0..1] (07_train_restoration.py:80), ResUNet.enc1 (14_train_unified_advanced.py:122)
 * and VGG16 features[0..1] (torchvision vgg16, used at 18_test_unified_benchmark.py:58) together with the tensor
 * hand-off that precedes each of them: ToTensor (17_run_unified_inference.py:66) or ToTensor + Normalize
 * (18_test_unified_benchmark.py:28-32).
 *   in_fmt B2R_IN_F32_NCHW: `in` is f32 [N,3,H,W] (the nn.Module.forward argument)
 *   in_fmt B2R_IN_U8_NHWC : `in` is u8 [N,H,W,3]; x = u8/255, then (x - mean[c]) / std[c] when `mean`/`std` are
 *                           non-null (host pointers to 3 floats).
 * weights_packed: bf16 [64][64], row co, column 2*t + j (j = 0, 1) = w[co][ci][kh][kw] with t = (kh*3 + kw)*3 + ci,
 *                 columns >= 54 zero: every weight appears twice because the kernel feeds each fp32 input as a
 *                 (hi, lo) bf16 pair (csrc/conv_c3.cu); the host packs it from the state_dict's f32 OIHW tensor.
 * bias f32 [64].  Runs on the tensor cores (tcgen05) with an im2col producer; H, W arbitrary.
 * ------------------------------------------------------------------------------------------------------------- */
#define B2R_IN_F32_NCHW 0
#define B2R_IN_U8_NHWC 1

int b2r_conv3x3_c3(const void* in, int in_fmt, const float* mean_host, const float* std_host,
                   const void* weights_packed, const float* bias, int act, float slope, void* out_nhwc_bf16, int N,
                   int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (3) Implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged NHWC bf16 tiles, TMEM f32 accumulators).
 *
 * One call = one fused layer:  out = act( sum_kblocks  A_kb * W_kb  + bias ), optionally followed by a fused
 * 2x2 max-pool (second output), with the M dimension = output pixels (128 per tile) and N = output channels.
 * The K loop is a list of "k-blocks": 64 input channels of one source tensor at one spatial offset.  That single
 * mechanism gives
 *   - conv3x3, padding 1 (nine offsets per 64-channel chunk; zero padding = TMA out-of-bounds fill);
 *   - concat-free decoder convs: torch.cat((up, skip), 1) (07:112, 14:171) is two sources in one K loop;
 *   - the ResidualBlock shortcut (14:107-115): the 1x1 conv (or identity) on the block input is one more centre
 *     k-block group, so BN(conv(.)) + shortcut(x) and the ReLU happen in one accumulator / one epilogue;
 *   - ConvTranspose2d(k=2, s=2) (07:91, 14:140): a 1x1 GEMM with 4*C_out rows of weights, each quadrant stored
 *     through its own strided output view (out_mode B2R_OUT_CONVT2X2);
 *   - nn.Linear (VGG16 classifier): H = 1, W = batch rows, one source.
 * Eval-mode BatchNorm (14:101,104) is folded into `weights`/`bias` by the host at load time.
 *
 * Replaces the nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d / nn.ReLU / nn.PReLU / nn.MaxPool2d / torch.cat /
 * nn.Linear calls inside SimpleUNet.forward (07_train_restoration.py:99-120), ResidualBlock.forward and
 * ResUNet.forward (14_train_unified_advanced.py:114-115, 151-186) and torchvision VGG16.forward
 * (18_test_unified_benchmark.py:46).
 * ------------------------------------------------------------------------------------------------------------- */
#define B2R_DBG_TILES 64
#define B2R_MAX_SRC 3
#define B2R_MAX_KBLOCKS 96
#define B2R_OUT_NHWC 0
#define B2R_OUT_CONVT2X2 1

/* flags: B2R_CONV_GENERIC_ONLY keeps the layer on the generic kernel even when the C_out = 64 specialisation
 * (resident weights + column-shifted halo boxes, csrc/conv_n64.cu) applies; used by the A/B parity test. */
#define B2R_CONV_GENERIC_ONLY 1
#define B2R_CONV_NO_W3 2 /* skip the tap-folded kernel (csrc/conv_w3.cu) even if weights_w3 is given */
#define B2R_CONV_NO_PAIR 8 /* C_out % 256 == 0 layers: one CTA per tile instead of cta_group::2 CTA pairs (A/B parity test) */
#define B2R_CONV_NO_HALO 4 /* generic kernel: one box per tap instead of the (TH + 2)-row halo boxes (A/B parity test) */

/* k-block encoding: bits [0,2) source index, [2,4) dh+1, [4,6) dw+1, [8,24) first channel / 64 */
#define B2R_KBLOCK(src, dh, dw, c64) \
    ((uint32_t)(src) | ((uint32_t)((dh) + 1) << 2) | ((uint32_t)((dw) + 1) << 4) | ((uint32_t)(c64) << 8))

typedef struct b2r_conv_gemm_desc {
    int32_t N, H, W;                 /* pixel grid of the sources (and of `out` for B2R_OUT_NHWC) */
    int32_t num_src;
    const void* src[B2R_MAX_SRC];    /* bf16 NHWC [N,H,W,src_C[i]] */
    int32_t src_C[B2R_MAX_SRC];      /* multiples of 64 */
    const void* weights;             /* bf16 [cout_total][64*num_kblocks], K order = k-block order */
    const void* weights_w3;          /* optional, C_out = 64 only: the same weights as bf16 [192][64*num_ksteps] for the
                                        tap-folded kernel (csrc/conv_w3.cu): one k-step per (3x3 group, kernel row kh)
                                        with rows (kw, co), and one per centre block with rows kw = 0, 2 zero; NULL =
                                        use the [64][K] layout only */
    const float* bias;               /* f32 [cout_total] */
    int32_t cout_total;              /* multiple of 64; for CONVT2X2 = 4*C_out, quadrant-major (q = 2*i + j) */
    int32_t num_kblocks;
    const uint32_t* kblocks_host;    /* HOST pointer, num_kblocks entries; NULL = source 0, centre tap, c = 64*kb */
    int32_t act;
    float slope;
    int32_t out_mode;
    void* out;                       /* bf16 NHWC; NHWC mode: [N,H,W,out_C]; CONVT: [N,2H,2W,out_C]; may be NULL */
    void* out_pool;                  /* bf16 NHWC [N,H/2,W/2,out_C] (2x2 max-pool of the activated output) or NULL */
    int32_t out_C;                   /* channel pitch of out / out_pool */
    int32_t tile_w, tile_h, tile_n;  /* pixels per tile along W, H, N (product 128); 0,0,0 = choose */
    int32_t block_n;                 /* 64 / 128 / 256; 0 = choose */
    int32_t max_ctas;                /* 0 = one per SM */
    int32_t flags;                   /* B2R_CONV_* */
    /* Optional fused output head (C_out = 64 layers on the tap-folded kernel only): the restorers' final 1x1 conv
     * 64 -> 3 (07_train_restoration.py:97,119; 14_train_unified_advanced.py:149,186) applied to this layer's activated
     * fp32 output before it is rounded to bf16, plus the reference's post-processing, so the 64-channel tensor never
     * goes to HBM.  head_w f32 [3][64], head_b f32 [3]; head_out_f32 = f32 NCHW [N,3,H,W] (module output, unclamped)
     * and/or head_out_u8 = u8 NHWC [N,H,W,3] (clamp(0,1)*255 truncated, 17_run_unified_inference.py:86-92).
     * With a head, `out` / `out_pool` may both be NULL. */
    const float* head_w;
    const float* head_b;
    float* head_out_f32;
    uint8_t* head_out_u8;
    int64_t* debug_timeline;         /* optional device buffer int64[B2R_DBG_TILES][8]: CTA 0 records clock64() stamps of
                                        its first tiles per warp role (tools/role_timeline.py); NULL = off */
} b2r_conv_gemm_desc;

int b2r_conv_gemm(const b2r_conv_gemm_desc* desc /* host */, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (4) Output head of the restorers: final 1x1 conv 64 -> 3 (07:97, 14:149) + the reference's post-processing.
 *   out_f32_nchw (optional): raw f32 [N,3,H,W] — the nn.Module.forward result, unclamped.
 *   out_u8_nhwc  (optional): clamp(0,1) -> x*255 -> truncation to u8, [N,H,W,3]
 *                            (17_run_unified_inference.py:86-92, 08_run_inference.py:96-98, 15_test_unified.py:186-188).
 * weights f32 [3][64], bias f32 [3].
 * ------------------------------------------------------------------------------------------------------------- */
int b2r_final_conv1x1(const void* in_nhwc_bf16, const float* weights, const float* bias, float* out_f32_nchw,
                      uint8_t* out_u8_nhwc, int N, int H, int W, void* stream);

/* 2x2/2 max-pool on NHWC bf16 (nn.MaxPool2d(2,2): 07:81, 14:124) — standalone form of the fused epilogue; odd H / W drop the
 * last row / column like PyTorch (output [N, H/2, W/2, C] with floor division). */
int b2r_maxpool2x2(const void* in_nhwc_bf16, void* out_nhwc_bf16, int N, int H, int W, int C, void* stream);

/* torch.nn.functional.interpolate(x, size=(out_h, out_w)) in its default mode 'nearest' on NHWC bf16: the re-alignment of the
 * up-sampled tensor to its skip connection in ResUNet.forward when H or W is not a multiple of 8
 * (14_train_unified_advanced.py:169-170, 175-176, 181-182).  src = min(floor(dst * (float)in / out), in - 1).  C % 8 == 0. */
int b2r_resize_nearest_bf16(const void* in_nhwc_bf16, void* out_nhwc_bf16, int N, int H, int W, int out_h, int out_w, int C,
                            void* stream);

/* AdaptiveAvgPool2d((7,7)) of torchvision VGG16 on NHWC bf16 [N,H,W,C] -> [N,7,7,C] (identity at H=W=7). */
int b2r_adaptive_avgpool7(const void* in_nhwc_bf16, void* out_nhwc_bf16, int N, int H, int W, int C, void* stream);

/* Small-N linear layer with f32 output: out[b][o] = sum_k in[b][k] * w[o][k] + bias[o]
 * (VGG16 classifier[6] = nn.Linear(4096, 43): 18_test_unified_benchmark.py:59).  in bf16 [B,K], w bf16 [O,K]. */
int b2r_linear_f32out(const void* in_bf16, const void* w_bf16, const float* bias, float* out, int B, int K, int O,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (5) Top-1 + correct count: torch.max(outputs, 1) and (predicted == labels).sum()
 * (18_test_unified_benchmark.py:47-49, 06_test_baseline.py:53-55).  Lowest index wins ties, like torch.max.
 * pred (optional) int64 [N]; conf (optional) f32 [N] = softmax probability of the arg-max (15_test_unified.py:125-129);
 * labels (optional) int64 [N]; counts (optional) int64[2] += {correct, N} (accumulated with one atomic per block).
 * ------------------------------------------------------------------------------------------------------------- */
int b2r_argmax_count(const float* logits, const int64_t* labels, int64_t* pred, float* conf, int64_t* counts, int N,
                     int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (6) Callers either side of the hot path: the single-degradation dataset generators with the
 * reference's exact arithmetic, and the image-quality reduction.  Images are u8, `elems_per_image` contiguous bytes
 * each (H*W*3 for the reference's HWC arrays).  All streaming, HBM-bound.
 * ------------------------------------------------------------------------------------------------------------- */

/* out[n][i] = lut[n][in[n][i]]; lut u8 [N][256] (4-byte aligned).  Carries the reference's float64 point operations on
 * u8 images exactly, the host evaluating the 256 possible results with the reference's own arithmetic: fog on
 * `image / 255.0` (04_gen_fog.py:12-31 with its clipped random t; 13_pipeline_stress_test.py:50-56). */
int b2r_lut_u8(const uint8_t* in, const uint8_t* lut, uint8_t* out, int N, int64_t elems_per_image, void* stream);

/* Per-image extrema over all channels, minmax int32 [N][2] = {min, max} (the cv2.minMaxIdx step of
 * cv2.normalize(blurred, blurred, 0, 255, cv2.NORM_MINMAX), 03_gen_blur.py:29). */
int b2r_minmax_u8(const uint8_t* in, int32_t* minmax, int N, int64_t elems_per_image, void* stream);

/* The convertTo step of the same cv2.normalize call given the extrema: scale = 255 / (max - min) (0 for a constant
 * image), shift = -min * scale in double, both cast to float, out = saturate_u8(rint(fma(in, scale, shift))).
 * in == out is allowed (the reference normalises in place). */
int b2r_normalize_minmax_u8(const uint8_t* in, const int32_t* minmax, uint8_t* out, int N, int64_t elems_per_image,
                            void* stream);

/* The float64 noise generators: x = image / 255 (float64) + noise, then
 *   B2R_NOISE_CLIP_SCRIPT02  02_gen_noise.py:12-27 add_gaussian_noise: low = -1 if any x of the image < 0 else 0;
 *                            out = np.uint8(clip(x, low, 1) * 255), i.e. truncation toward zero followed by wrap-around
 *                            modulo 256 for negative values (-127.5 -> 129).  Two passes (the decision needs the image).
 *   B2R_NOISE_CLIP_UNIT      13_pipeline_stress_test.py:33-38 add_noise: out = (clip(x, 0, 1) * 255).astype(uint8).
 * noise: optional float64 [N][elems] (the reference's own draw, for parity tests); NULL draws N(0, sigma[n]^2) from
 * Philox4x32-10 keyed by (seed, image_index0 + n, pixel) as b2r_degrade does.
 * neg_flags: int32 [N] workspace (the per-image "any negative" decision of script 02; left filled for inspection). */
#define B2R_NOISE_CLIP_SCRIPT02 0
#define B2R_NOISE_CLIP_UNIT 1
int b2r_noise02(const uint8_t* in, uint8_t* out, int N, int64_t elems_per_image, const float* sigma, const double* noise,
                uint64_t seed, uint64_t image_index0, int32_t* neg_flags, int clip_rule, void* stream);

/* transforms.Resize((out_h, out_w)) of a ragged batch of PIL images (17_run_unified_inference.py:66,79-82;
 * 18_test_unified_benchmark.py:28-30), i.e. Pillow's BILINEAR resampling, bit for bit.
 * src: the images packed back to back, u8 HWC RGB; offsets int64 [N] = byte offset of image n; hw int32 [N][2] = its
 * height, width.  tabs int32 [T][S][2 + K]: Pillow's coefficient tables, one per distinct (input size -> output size)
 * pair, entry = {first input index, tap count, taps in 22-bit fixed point}; xtab_index / ytab_index int32 [N] select the
 * table of image n for the horizontal / vertical pass.  out u8 [N][out_h][out_w][3].
 * tile_rows output rows share one CTA, max_rows = the largest number of input rows such a tile needs (both from the
 * host, which owns the tables: imageio.resize_batch). */
int b2r_resize_bilinear_u8(const uint8_t* src, const int64_t* offsets, const int32_t* hw, const int32_t* xtab_index,
                           const int32_t* ytab_index, const int32_t* tabs, int K, int S, uint8_t* out, int N, int out_h,
                           int out_w, int tile_rows, int max_rows, void* stream);

/* cv2.resize(img, (out_w, out_h)) (default INTER_LINEAR) of a ragged batch of u8 HWC images, bit for bit: what
 * 08_run_inference.py:119 applies to the clean image before PSNR / SSIM.  Packing as for b2r_resize_bilinear_u8;
 * tabs int32 [T][S][3] = {first source index, cvRound((1 - f) * 2048), cvRound(f * 2048)} per output coordinate, from the
 * host (imageio.cv_linear_table restates OpenCV's coordinate arithmetic); out u8 [N][out_h][out_w][3]. */
int b2r_resize_cv_linear_u8(const uint8_t* src, const int64_t* offsets, const int32_t* hw, const int32_t* xtab_index,
                            const int32_t* ytab_index, const int32_t* tabs, int S, uint8_t* out, int N, int out_h,
                            int out_w, void* stream);

/* VGG feature taps: mean over the middle axis of a bf16 tensor viewed as [outer][reduce][inner], f32 [outer][inner] out.
 * inner = 1: channel mean of an NHWC feature map (11_visualize_hidden_states.py:50, torch.mean(features, dim=1));
 * inner > 1: global average pooling of [N][H*W][C] (12_generate_umap_pt.py:52, torch.mean(feature, dim=[2, 3])).
 * The contiguous extent (reduce if inner = 1, else inner) must be a multiple of 8; fp32 accumulation. */
int b2r_mean_bf16(const void* in_bf16, float* out, int64_t outer, int reduce, int inner, void* stream);

/* sse[n] = sum_i (a[n][i] - b[n][i])^2 as uint64: PSNR = 10 log10(255^2 * elems / sse) (08_run_inference.py:118-129
 * reports skimage's peak_signal_noise_ratio on the u8 images). */
int b2r_sse_u8(const uint8_t* a, const uint8_t* b, uint64_t* sse, int N, int64_t elems_per_image, void* stream);

/* ssim[n] = skimage.metrics.structural_similarity(a[n], b[n], data_range=data_range, channel_axis=2) for u8 HWC images
 * [N][H][W][C] (08_run_inference.py:123): 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, float64, the
 * 3-pixel border cropped, mean over pixels and channels.  Window sums are exact integers; deterministic.
 * H, W >= 7 (skimage raises ValueError below that). */
int b2r_ssim_u8(const uint8_t* a, const uint8_t* b, double* ssim, int N, int H, int W, int C, double data_range,
                void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (7) Whole-network entry points: the layer graphs of the three modules behind ONE call each, for hosts that are not
 * Python.  Replaces, for such a host, `model = ResUNet(); model.load_state_dict(torch.load(path)); model.eval()`
 * (17_run_unified_inference.py:59-64) + `model(input_tensor)` (17:85) [+ clamp / x255 / astype(uint8), 17:86-92], the same
 * for SimpleUNet (08_run_inference.py:68-70,93-98), and `models.vgg16` + `classifier[6] = nn.Linear(4096, 43)` +
 * `model(inputs)` (18_test_unified_benchmark.py:58-59,46) with its Resize/ToTensor/Normalize hand-off (18:28-32).
 *
 * b2r_net_create packs a reference state_dict — HOST float32 tensors under the reference's own key names, e.g. what
 * torch.load('restoration_unified_resnet.pth') holds — into the device layouts of sections (2)-(3): eval-mode BatchNorm
 * folded in fp64, k-block ordered bf16 weight matrices, ConvTranspose as four stacked 1x1 matrices, classifier[0]
 * columns permuted to NHWC.  Strict like load_state_dict: a missing key or a wrong shape returns B2R_EINVAL and
 * b2r_last_error() names the key; int64 `num_batches_tracked` entries and unknown keys are ignored.
 * Ownership: the caller owns `dev_weights` (>= b2r_net_weight_bytes, 256-byte aligned) and every workspace; the plan holds
 * pointers into dev_weights until b2r_net_destroy.  The forwards enqueue on `stream` and never allocate.
 * in_fmt as in section (2): B2R_IN_F32_NCHW (the nn.Module.forward argument) or B2R_IN_U8_NHWC (ToTensor fused).
 * H, W: multiples of 4 (SimpleUNet) / 8 (ResUNet) / 32 (VGG16), else B2R_EINVAL.  (The reference's nearest-neighbour
 * re-alignment for other ResUNet sizes, 14:169-183, is available to a C host through b2r_maxpool2x2 + b2r_resize_nearest_bf16
 * + b2r_conv_gemm and is what the Python module does; the one-call forward keeps the fully fused graph.)
 * ------------------------------------------------------------------------------------------------------------- */
#define B2R_NET_SIMPLE_UNET 0
#define B2R_NET_RESUNET 1
#define B2R_NET_VGG16 2
#define B2R_DT_F32 0
#define B2R_DT_I64 1

typedef struct b2r_tensor {      /* one state_dict entry */
    const char* name;            /* reference key, e.g. "res1.conv_block.1.running_var" */
    const void* data;            /* HOST pointer, contiguous */
    int32_t dtype;               /* B2R_DT_* */
    int32_t ndim;                /* 0..4 */
    int64_t shape[4];
} b2r_tensor;

typedef struct b2r_net b2r_net;  /* opaque */

/* upper bound of the device bytes b2r_net_create needs for this state_dict */
int b2r_net_weight_bytes(const b2r_tensor* state, int num_tensors, size_t* bytes);
int b2r_net_create(int arch, int num_classes /* VGG16 head; ignored otherwise */, const b2r_tensor* state, int num_tensors,
                   void* dev_weights, size_t dev_weight_bytes, void* stream, b2r_net** net);
void b2r_net_destroy(b2r_net* net);
/* activation workspace (device, 1024-byte aligned) one forward of N x H x W needs */
int b2r_net_workspace_bytes(const b2r_net* net, int N, int H, int W, size_t* bytes);

/* SimpleUNet.forward (07_train_restoration.py:99-120) / ResUNet.forward (14_train_unified_advanced.py:151-186).
 * out_f32_nchw (optional): f32 [N,3,H,W], the module output, unclamped; out_u8_nhwc (optional): clamp(0,1) * 255 truncated,
 * u8 [N,H,W,3] (17:86-92).  At least one output. */
int b2r_unet_forward(const b2r_net* net, const void* in, int in_fmt, float* out_f32_nchw, uint8_t* out_u8_nhwc, int N, int H,
                     int W, void* workspace, size_t workspace_bytes, void* stream);
int b2r_resunet_forward(const b2r_net* net, const void* in, int in_fmt, float* out_f32_nchw, uint8_t* out_u8_nhwc, int N, int H,
                        int W, void* workspace, size_t workspace_bytes, void* stream);
/* VGG16-43 forward: logits f32 [N, num_classes].  normalize = 1 (u8 input only): ToTensor + Normalize(ImageNet mean / std,
 * 18:28-32) fused into the first conv; normalize = 0: the input is the already normalised tensor (f32 NCHW) or plain u8/255. */
int b2r_vgg16_forward(const b2r_net* net, const void* in, int in_fmt, int normalize, float* logits, int N, int H, int W,
                      void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H_ */
