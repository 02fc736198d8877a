/* pda_b200.h -- C ABI of libpda_b200.so: hand-written sm_100a kernels for the Probabilistic U-Net
 * training / Monte-Carlo inference path of Probabilistic-Domain-Adaptation.
 *
 * The reference has no FFI for this path: its boundary is the Python nn.Module API of
 * prob_utils.my_models.  Each entry point below therefore names the reference
 * Python code whose ATen/cuDNN launches it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; buffers are owned by the caller
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *  - activations are NHWC bf16 ([B][H][W][C], C % 64 == 0); images / labels / outputs with C == 1 are fp32
 *  - return value: PDA_OK (0) or a negative PDA_ERR_* code; pda_error_string() names it
 */
#ifndef PDA_B200_H_
#define PDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDA_OK 0
#define PDA_ERR_SHAPE (-1)     /* unsupported / inconsistent sizes */
#define PDA_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define PDA_ERR_DRIVER (-3)    /* cuTensorMapEncodeTiled entry point not found */
#define PDA_ERR_TENSORMAP (-4) /* TMA descriptor encoding failed */
#define PDA_ERR_ARG (-5)       /* null pointer / bad flag */

#define PDA_ABI_VERSION 1

int pda_abi_version(void);
const char* pda_error_string(int code);

/* Number of kernels launched through this library since load / the last reset (bench.py "gpu_launches"). */
long long pda_launch_count(void);
void pda_reset_launch_count(void);

/* ACTIVATION FORMAT.  Feature maps are NHWC 16-bit tensors in one of two storage formats, selected per call by
 * `act_f16`: 0 = bf16 (the training path: activation gradients need the fp32 exponent range), 1 = fp16 (the no-grad /
 * Monte-Carlo inference path: 11 mantissa bits instead of 8 cut the rounding error of the 21-layer trunk by 8x, which is
 * what brings sampled logits within the 1e-2 tolerance at full image sizes; fp16 is also the dtype the reference's own
 * student forward runs in under torch_em's autocast).  Tensor-core operands match the activations (bf16 x bf16 or
 * fp16 x fp16), accumulation is always fp32.  fp16 stores saturate at +-65504 (never inf); kernels that can produce
 * large values take `range_flag` (device int, may be NULL) and set it to 1 when a value exceeded the range. */

/* Weights of nn.Conv2d(cin, cout, 3): OIHW fp32 -> [cout][tap = ky*3+kx][cin] 16-bit (K-major GEMM B operand), bf16 or
 * fp16 (f16 != 0).  With rot180 != 0 the taps are reversed and cin/cout swapped ([cin][8-tap][cout]): the dgrad operand. */
int pda_pack_conv3x3_weights(const float* w_oihw, void* w_packed, int cout, int cin, int rot180, int f16,
                             void* stream);

/* Same for many convs in one launch (after an optimizer or EMA step).  table: device int64 [n_chunks][8] =
 * (w_oihw ptr, bf16 packed ptr or 0, bf16 rot180-packed ptr or 0, fp16 packed ptr or 0, cout, cin, co0, ci0): one row =
 * one 32 (co) x 32 (ci) x 9 tile starting at (co0, ci0); cout and cin must be multiples of 32. */
int pda_pack_conv3x3_weights_multi(const int64_t* table, int n_chunks, void* stream);

/* First layer of every net (cin = 1, or 2 for the posterior whose input is cat(patch, segm),
 * probabilistic_unet.py:118): x0/x1 are fp32 [B][H][W] planes (x1 may be NULL), w is OIHW fp32.
 * Replaces unet_blocks.py:19-20 / probabilistic_unet.py:56-57 for block 0.  out: NHWC bf16 / fp16 (act_f16). */
int pda_conv3x3_first(const float* x0, const float* x1, const float* w_oihw, const float* bias, void* out, int B,
                      int H, int W, int cout, int relu, int act_f16, void* stream);

/* conv3x3(pad 1) + bias (+ReLU) (+ 2x2 average pool) on tcgen05 tensor cores.
 * Input = channel concat of src0 (c0 ch) and src1 (c1 ch, may be NULL/0) -- the torch.cat of
 * unet_blocks.py:56 is never materialised.  out and/or out_pool may be NULL.
 * Replaces unet_blocks.py:17-24 (DownConvBlock) and probabilistic_unet.py:53-61 (Encoder).
 * relu_mask (NHWC bf16 [B][H][W][cout], may be NULL): outputs are zeroed where relu_mask <= 0 -- the ReLU backward of
 * the layer that produced this conv's input, fused into the epilogue when the kernel runs as dgrad.
 * bn_tile: 0 = auto, else 64/128 output channels per CTA.  act_f16 / range_flag: see ACTIVATION FORMAT above (inputs,
 * packed weights and outputs share the format). */
int pda_conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* w_packed, const float* bias,
                   void* out, void* out_pool, const void* relu_mask, int B, int H, int W, int cout, int relu,
                   int bn_tile, int act_f16, int* range_flag, void* stream);

/* The first conv of an UpConvBlock (unet_blocks.py:51-57) with the bilinear x2 up-sampling FUSED: input = channel concat of
 * F.interpolate(up_src, scale_factor=2, mode='bilinear', align_corners=True) (up_src NHWC [B][H/2][W/2][c0], never
 * materialised at full resolution: the conv kernel's epilogue warps interpolate the operand slabs straight into shared
 * memory) and src1 [B][H][W][c1].  Same outputs, bit for bit, as pda_upsample2x_bilinear followed by pda_conv3x3_tc.
 * H, W even; CTA-pair kernel only (PDA_ERR_SHAPE for images of a single pixel tile). */
int pda_conv3x3_up_tc(const void* up_src, int c0, const void* src1, int c1, const void* w_packed, const float* bias,
                      void* out, void* out_pool, int B, int H, int W, int cout, int relu, int act_f16, int* range_flag,
                      void* stream);

/* Kernel selection for pda_conv3x3_tc: 0 = one CTA per 128-pixel tile (csrc/conv3x3_tc.cu), 1 = CTA pairs issuing
 * M = 256 tcgen05.mma.cta_group::2 with the weight tile split across the pair (csrc/conv3x3_tc2.cu).  mode < 0 only
 * queries.  Returns the previous mode.  Results are bit-identical between the two. */
int pda_set_conv_pair(int mode);

/* Kernel selection for pda_conv3x3_first with cout = 64: 1 (default) = tensor-core kernel (kind::tf32 with split operands:
 * fp32-equivalent accuracy, csrc/conv_first_tc.cu), 0 = CUDA-core kernel.  mode < 0 only queries.  Returns the previous
 * mode. */
int pda_set_first_conv_tc(int mode);

/* SMs the persistent tensor-core kernels (conv, weight gradient, Fcomb backward) may occupy (default 148).  Data-parallel
 * training leaves a few SMs to NCCL so that the gradient all-reduce runs next to the backward kernels instead of
 * between them.  sms <= 0 only queries.  Returns the previous value. */
int pda_set_sm_budget(int sms);

/* Same contract (bf16 only) on plain CUDA cores (one thread per output element).  Cross-check kernel for the
 * parity tests; the product path never selects it implicitly. */
int pda_conv3x3_bf16_simt(const void* src0, int c0, const void* src1, int c1, const void* w_packed,
                          const float* bias, void* out, void* out_pool, int B, int H, int W, int cout, int relu,
                          void* stream);

/* nn.AvgPool2d(2, 2, 0, ceil_mode=True) for even H, W (unet_blocks.py:17).  NHWC bf16 / fp16. */
int pda_avgpool2(const void* in, void* out, int B, int H, int W, int C, int act_f16, void* stream);

/* F.interpolate(mode='bilinear', scale_factor=2, align_corners=True) (unet_blocks.py:51). NHWC (h,w)->(2h,2w) */
int pda_upsample2x_bilinear(const void* in, void* out, int B, int h, int w, int C, int act_f16, void* stream);

/* AxisAlignedConvGaussian head (probabilistic_unet.py:126-137): mean over H then W of the encoder output
 * enc [B][P][C] bf16 / fp16, then the 1x1 conv C -> 2*latent (w_head [2L][C] fp32, b_head [2L]).
 * scratch: fp32, at least B * pda_gauss_head_scratch_rows(P) * C elements.  out: [B][2L] fp32 = (mu | log_sigma). */
int pda_gauss_head_scratch_rows(int P);
int pda_gauss_head(const void* enc, const float* w_head, const float* b_head, float* scratch, float* mu_logsigma,
                   int B, int P, int C, int latent, int act_f16, void* stream);

/* z[s][b][:] = mu[b] + exp(log_sigma[b]) * eps[s][b]  (Normal.rsample, probabilistic_unet.py:302/349). */
int pda_latent_samples(const float* mu_logsigma, const float* eps, float* z, int S, int B, int latent, void* stream);

/* Analytic KL(q || p) of two diagonal Gaussians (probabilistic_unet.py:332), per batch element -> kl[B]. */
int pda_kl_diag_gauss(const float* mu_logsigma_q, const float* mu_logsigma_p, float* kl, int B, int latent,
                      void* stream);

/* Fused Fcomb + sigmoid + cross-sample mean + consensus for S latent samples.
 * Replaces S x Fcomb.forward (probabilistic_unet.py:200-214: tile, cat, 3 x conv1x1) and the consensus
 * arithmetic of mean_teacher_trainer.py:74-86 (+ 3 copies) / punet_predictions.py:31-32,117-124.
 *   feat [B][P][64] bf16 / fp16 (feat_f16), z [S][B][L] fp32,
 *   w1 [64][64+L] (first 64 input channels = features, last L = z), b1[64], w2[64][64], b2[64], w3[64], b3[1]: fp32.
 * Outputs (any may be NULL): mean_prob [B][P] fp32 = sum_s sigmoid(logit_s) / S;
 *   cons_weight [B][P] fp32 = #{s: p_s >= upper or p_s <= lower} / S;  cons_mask [B][P] int64 = (count == S);
 *   logits / probs [S][B][P] fp32 (per-sample, for sample()/reconstruct() and the parity tests).
 * scratch: caller-allocated fp32 [pda_fcomb_scratch_floats(S, B)], private to this call (no state is shared between
 *   launches, streams or CUDA graphs).  Word 0 is the fp16 RANGE FLAG (int): the hidden layer runs in packed fp16;
 *   when |F.W1f| or |b1 + W1z.z| reaches 32000 anywhere in the batch the flag is raised on the device and the same
 *   call re-computes every output with the exact fp32 kernel below (no host synchronisation).  Pass the same scratch
 *   to pda_fcomb_bwd so that the backward follows the path the forward took. */
long long pda_fcomb_scratch_floats(int S, int B);
int pda_fcomb_mc_consensus(const void* feat, const float* z, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int B, int P, int S, int latent,
                           float upper, float lower, float* mean_prob, float* cons_weight, int64_t* cons_mask,
                           float* logits, float* probs, float* scratch, int feat_f16, void* stream);

/* Same contract in exact-order fp32 on CUDA cores (no bf16 rounding of weights / hidden activations): the
 * numerics baseline of the tensor-core kernel above.  ~20x slower; selected explicitly by the caller only. */
int pda_fcomb_mc_consensus_fp32(const void* feat, const float* z, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, int B, int P, int S, int latent,
                                float upper, float lower, float* mean_prob, float* cons_weight, int64_t* cons_mask,
                                float* logits, float* probs, int feat_f16, void* stream);

/* Same contract for any depth: no_convs_fcomb = n_mid + 2 with n_mid >= 0 hidden 64 -> 64 layers (wmid [n_mid][64][64],
 * bmid [n_mid][64]).  The reference's DEFAULT constructor builds no_convs_fcomb = 4 (probabilistic_unet.py:231-236); every
 * script uses 3, which the two kernels above serve.  Plain fp32, one pixel per thread: a correctness path for non-script
 * architectures (forward only). */
int pda_fcomb_mc_consensus_deep(const void* feat, const float* z, const float* w1, const float* b1, const float* wmid,
                                const float* bmid, int n_mid, const float* w3, const float* b3, int B, int P, int S,
                                int latent, float upper, float lower, float* mean_prob, float* cons_weight,
                                int64_t* cons_mask, float* logits, float* probs, int feat_f16, void* stream);

/* Mean-teacher EMA over many tensors in one launch: t = t*m + p*(1-m)  (mean_teacher_trainer.py:52-55,
 * adamt_trainer.py:40-43).  table: device int64 [n_chunks][3] = (teacher_ptr, student_ptr, numel<=65536). */
int pda_multi_tensor_ema(const int64_t* table, int n_chunks, double momentum, void* stream);
/* The AdaMT form (adamt_trainer.py:40-43): momentum_t = min(1 - 1 / (iteration + 1), momentum) with the iteration count
 * read from DEVICE memory (int64; incremented by the call), so that a CUDA graph that captured the step can be replayed. */
int pda_multi_tensor_ema_warmup(const int64_t* table, int n_chunks, double momentum, int64_t* iteration_dev,
                                void* stream);

/* Tiled prediction driver (the reference calls torch_em.util.prediction.predict_with_halo,
 * punet_predictions.py:41-49, one block at a time from numpy).  rois / inner: device int32 [T][4] = (y0, x0, h, w) in
 * image coordinates; all T outer blocks have the same size th x tw.
 * gather: out[t] = (image[roi_t] - mean_t) / (std_t + 1e-7)  (per-block standardisation, population std);
 *         stats: double [2*T] scratch.   scatter: out_image[inner_t] = pred[t][inner_t - roi_t origin]. */
int pda_tile_gather_standardize(const float* image, int H, int W, const int32_t* rois, int T, int th, int tw,
                                double* stats, float* out, void* stream);
int pda_tile_scatter(const float* pred, int T, int th, int tw, const int32_t* rois, const int32_t* inner, float* out,
                     int H, int W, void* stream);

/* On-device weak / strong view augmentation.  Replaces the per-sample CPU transforms that the
 * DataLoader workers run: get_raw_transform(normalizer=my_standardize_torch, augmentation1=Compose([my_standardize_torch,
 * RandomApply(GaussianBlur), RandomApply(AdditiveGaussianNoise), RandomApply(RandomContrast)]))
 * (MitoEM/common.py:50-68, LIVECell/livecell_fm.py:43-67, livecell_adamatch.py:16-38, applied at
 * prob_utils/my_datasets/my_image_collection_dataset.py:349-357; my_standardize_torch = prob_utils/my_utils/util.py:9-14).
 * pda_image_stats: stats[b] = (sum, sum of squares) of image b (fp64), n = pixels per image; shared by all views.
 * pda_augment_view: out[b] = contrast(noise(blur(standardize^k(img[b])))) in one kernel.  params: device float [B][8] =
 *   (blur kernel size (odd, <= 31; <= 1: no blur), sigma, noise scale (0: none), contrast alpha (1: none), contrast mean,
 *    k = number of standardisations (0..2), unused, unused) -- the random decisions, drawn on the host in the reference's
 *   order.  noise: unit-normal field [B][H][W] or NULL.  Blur = torchvision GaussianBlur (reflect padding, separable).
 *   max_ksize: the largest kernel size in `params` (validated against H, W on the host). */
int pda_image_stats(const float* img, int B, long long n, double* stats, void* stream);
int pda_augment_view(const float* img, const float* noise, float* out, int B, int H, int W, const double* stats,
                     const float* params, float eps, int max_ksize, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training (backward) entry points.  They replace what torch.autograd runs behind loss.backward() in the step
 * bodies punet_trainer.py:24-36 / mean_teacher_trainer.py:111-119 (cuDNN dgrad/wgrad, ATen elementwise backward).
 * dgrad of conv3x3 is pda_conv3x3_tc itself, called with the rot180-packed weights and relu = 0.
 * --------------------------------------------------------------------------------------------------------- */

/* Weight (+bias) gradient of conv3x3: dW[co][ci][ky][kx] = sum_p dZ[p][co] * X[p + tap][ci] on tcgen05 tensor cores
 * (K = pixels; both operands read as MN-major straight from NHWC; persistent stream-K schedule).
 * X = concat(src0, src1) as in the forward.
 * dz: NHWC bf16 [B][H][W][cout] (already masked by ReLU).  dw_oihw fp32 OIHW, dbias fp32 [cout] (may be NULL).
 * accumulate != 0 adds to dw/dbias instead of overwriting.
 * scratch: fp32 [pda_conv3x3_wgrad_scratch_floats(c0 + c1, cout)] (tap-major partial sums, bias sums).  It must be all
 *   zero when the kernel starts and the call leaves it all zero again (the re-layout kernel that follows clears what it
 *   reads): a caller that keeps one scratch per layer passes scratch_is_zero != 0 after having zeroed it once; with 0
 *   the call clears it first (one memset).  (Measured dead end: re-laying out inside the main kernel by the last CTA
 *   of each 64 x 64 item removed the second launch but cost +35 % kernel time -- one SM per item, latency-bound.) */
long long pda_conv3x3_wgrad_scratch_floats(int ctot, int cout);
int pda_conv3x3_wgrad_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz, float* scratch,
                           float* dw_oihw, float* dbias, int B, int H, int W, int cout, int accumulate,
                           int scratch_is_zero, void* stream);

/* Deterministic variant (bit-identical results from run to run for a fixed shape and SM budget): every CTA of the stream-K
 * schedule stores its partial accumulators to its own slot of `scratch` (plain stores; no memset, no atomics) and a
 * second kernel sums the slots of each 64 x 64 x 9 item in CTA order while re-laying out to OIHW.
 * scratch: fp32 [pda_conv3x3_wgrad_det_scratch_floats(c0 + c1, cout, B, H, W)], contents irrelevant on entry.
 * (torch.use_deterministic_algorithms-style opt-in: the host side selects it with PDA_WGRAD_DETERMINISTIC=1.) */
long long pda_conv3x3_wgrad_det_scratch_floats(int ctot, int cout, int B, int H, int W);
int pda_conv3x3_wgrad_bf16_det(const void* src0, int c0, const void* src1, int c1, const void* dz, float* scratch,
                               float* dw_oihw, float* dbias, int B, int H, int W, int cout, int accumulate,
                               void* stream);

/* dZ = (dFull + 0.25 * dPool[y/2][x/2]) * (Y > 0): ReLU backward fused with the backward of the 2x2 average pool that
 * consumes Y (unet_blocks.py:17,20).  NHWC bf16; dfull or dpool may be NULL; y == NULL skips the ReLU mask
 * (plain AvgPool2d backward).  dbias (fp32 [C], may be NULL) receives the conv bias gradient sum_pixels dZ. */
int pda_relu_pool_bwd_bf16(const void* dfull, const void* dpool, const void* y, void* dz, float* dbias, int B, int H,
                           int W, int C, void* stream);

/* Backward of the bilinear x2 upsample (unet_blocks.py:51): dout (2h,2w) -> din (h,w), NHWC bf16. */
int pda_upsample2x_bilinear_bwd_bf16(const void* dout, void* din, int B, int h, int w, int C, void* stream);

/* First layer (cin 1 or 2) weight/bias gradient; out = forward output (for the ReLU mask), dout its gradient.
 * out == NULL: dout is already masked by this layer's ReLU (it comes out of the next layer's dgrad conv, whose
 * epilogue applies that mask): the kernel then reads half the bytes.
 * scratch: caller-allocated fp32 [pda_conv3x3_first_bwd_scratch_floats(...)] for the per-block partial sums. */
long long pda_conv3x3_first_bwd_scratch_floats(int B, int H, int W, int cout, int cin);
int pda_conv3x3_first_bwd(const float* x0, const float* x1, const void* out, const void* dout, float* dw, float* db,
                          int B, int H, int W, int cout, float* scratch, void* stream);

/* Gaussian head: spatial mean [B][C] from the forward's stage-1 scratch, and the backward to the encoder output. */
int pda_gauss_head_mean(const float* scratch, float* mean, int B, int P, int C, void* stream);
int pda_gauss_head_bwd(const float* dmls, const float* w_head, const float* mean, const void* enc, float* dw,
                       float* db, float* dmean_scratch, void* denc, int B, int P, int C, int latent, void* stream);

/* d KL(q||p) / d (mu|log_sigma) of both distributions, given dkl [B]. */
int pda_kl_diag_gauss_bwd(const float* q, const float* p, const float* dkl, float* dq, float* dp, int B, int latent,
                          void* stream);

/* Reconstruction loss of ProbabilisticUnet.elbo (probabilistic_unet.py:347-369): BCE-with-logits (dice = 0) or
 * Dice-with-logits (dice = 1) of (logits * consm, segm * consm); consm fp32 or int64, both NULL = no mask.
 * out2 = (sum, mean); stats3 is kept for the backward.  partial: double [3 * pda_recon_loss_blocks(n)]. */
int pda_recon_loss_blocks(long long n);
int pda_recon_loss_fwd(const float* logits, const float* segm, const float* consm_f32, const int64_t* consm_i64,
                       long long n, int dice, double* partial, float* out2, float* stats3, void* stream);
int pda_recon_loss_bwd(const float* logits, const float* segm, const float* consm_f32, const int64_t* consm_i64,
                       long long n, int dice, const float* stats3, const float* gout2, float* dlogits, void* stream);

/* Validation metric dice_score (my_utils/util.py:17-44; punet_trainer.py:78-81 calls it on the Monte-Carlo mean and the
 * ground truth after a device->host copy): 2 sum(gt*seg) / (sum(gt) + sum(seg) + 1e-7).  thr_seg / thr_gt: NaN = no
 * threshold, else x -> (x > thr).  partial: double [3 * pda_recon_loss_blocks(n)].  out: fp32 [1]. */
int pda_dice_score(const float* seg, const float* gt, long long n, float thr_seg, float thr_gt, double* partial,
                   float* out, void* stream);

/* l2_regularisation (utils.py:32-40): out = sum_t ||W_t||_2 over many tensors in two launches.
 * table int64 [n_chunks][4] = (ptr, numel<=65536, tensor_index, 0);
 * backward table = (w_ptr, byte offset of this chunk's gradient inside grad_base, numel, tensor_index): the gradients
 * gout / ||W_t|| * W_t of all tensors are written into one flat fp32 buffer. */
int pda_multi_tensor_l2norm_fwd(const int64_t* table, int n_chunks, int n_tensors, double* partial, float* norms,
                                float* out, void* stream);
int pda_multi_tensor_l2norm_bwd(const int64_t* grad_table, int n_chunks, const float* norms, const float* gout,
                                void* grad_base, void* stream);

/* Fcomb backward for one latent sample z [B][L]: dlogit [B][P] -> dfeat [B][P][64] bf16, parameter grads, dz [B][L].
 * Five chained tcgen05 GEMMs per 128-pixel tile (bf16 operands, fp32 accumulate).  scratch: fp32 [64*64 + 2*B*64].
 * fwd_range_flag: word 0 of the scratch the forward call (pda_fcomb_mc_consensus, S = 1) used, or NULL.  When the
 *   forward raised its fp16 range flag, the tensor-core kernel steps aside and the exact fp32 kernel computes the
 *   gradients of the function that was actually evaluated (decided on the device). */
int pda_fcomb_bwd(const void* feat, const float* z, const float* w1, const float* b1, const float* w2, const float* b2,
                  const float* w3, const float* dlogit, int B, int P, int latent, void* dfeat, float* dw1, float* db1,
                  float* dw2, float* db2, float* dw3, float* db3, float* dz, float* scratch, const int* fwd_range_flag,
                  void* stream);

/* Same contract in exact-order fp32 on CUDA cores: the numerics baseline of the tensor-core kernel above (~10x
 * slower; explicit opt-in only).  scratch: fp32 [64*64 + B*64]. */
int pda_fcomb_bwd_fp32(const void* feat, const float* z, const float* w1, const float* b1, const float* w2,
                       const float* b2, const float* w3, const float* dlogit, int B, int P, int latent, void* dfeat,
                       float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3, float* dz,
                       float* scratch, void* stream);

/* torch.optim.Adam step (the optimizer of every reference script, e.g. LIVECell/livecell_punet.py:58) over many
 * tensors in one launch.  table int64 [n_chunks][5] = (param, grad, exp_avg, exp_avg_sq, numel<=65536), all fp32.
 * step >= 1 is the step count AFTER this update (bias correction).  inv_scale (device float, may be NULL) multiplies
 * the gradients first (GradScaler unscale); found_inf (device float, may be NULL) != 0 turns the launch into a no-op. */
int pda_multi_tensor_adam(const int64_t* table, int n_chunks, double lr, double beta1, double beta2, double eps,
                          double weight_decay, long long step, const float* inv_scale, const float* found_inf,
                          void* stream);

/* FixMatch distribution alignment (fixmatch_trainer.py:77-84): y_bin = y >= 0.5; target = class frequencies of y_bin
 * (a single-class batch gives target = [1.0] for both entries, as torch.unique's single count does); ratio = source /
 * target; out = clip(where(y < 0.5, y * ratio[0], y * ratio[1]), 0, 1).  source_dist: device float[2] = (background,
 * foreground) frequency of the source domain; scratch: device uint64[1]; ratio: device float[2] (output).  No host
 * synchronisation (the reference's torch.unique sorts and syncs). */
int pda_distribution_alignment(const float* y, long long n, const float* source_dist, unsigned long long* scratch,
                               float* out, float* ratio, void* stream);

/* The same step with the step count (int64, count BEFORE this update; incremented by the call) and the learning rate
 * (float) read from DEVICE memory: nothing step-dependent is baked into the launch, so a CUDA graph that captured it can
 * be replayed (torch.optim.Adam(capturable=True) semantics). */
int pda_multi_tensor_adam_capturable(const int64_t* table, int n_chunks, const float* lr_dev, double beta1, double beta2,
                                     double eps, double weight_decay, int64_t* step_dev, const float* inv_scale,
                                     const float* found_inf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDA_B200_H_ */
