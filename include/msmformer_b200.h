/*
 * msmformer_b200.h - C ABI of libmsmformer_b200.so, the sm_100a implementation of the MSMFormer
 * segmentation hot path (reference: YoungSean/UnseenObjectsWithMeanShift @ d1c8487).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless its name ends in _host; tensors are
 *     dense row-major with the innermost (channel) axis contiguous; sizes/strides are in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued on it, nothing synchronises, so the calls can be captured in a CUDA graph;
 *   - return value 0 = success; <0 = bad argument (MSM_E_*); >0 = cudaError_t of a failed launch.
 *     msm_last_error() returns a static, thread-local description of the last non-zero return;
 *   - `workspace` buffers are caller-owned scratch; query the size with the matching *_workspace_bytes.
 *
 * Reference paths below are relative to MSMFormer/meanshiftformer/modeling/ in the reference.
 */
#ifndef MSMFORMER_B200_H_
#define MSMFORMER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSM_E_BADARG (-1)
#define MSM_E_UNSUPPORTED (-2)
#define MSM_E_WORKSPACE (-3)

#define MSM_ABI_VERSION 1

int msm_abi_version(void);
const char* msm_last_error(void);
/* compute capability major*10+minor of the current device (100 on B200), <0 on error */
int msm_device_arch(void);

/* ------------------------------------------------------------------------------------------------
 * vMF ("hypersphere") attention core.
 * Replaces hypersphere_attention(q, k, v, attn_mask, dropout_p, kappa),
 *   transformer_decoder/attention_util.py:30-82:
 *     out = unit( softmax_s( kappa * unit(q).unit(k_s) + mask ) . v )
 * for G = batch*heads independent problems of Nq queries x Ns keys x hd channels.
 *
 * Element (b, h, i, d) of q lives at  q[b*q_sb + h*q_sh + i*q_sl + d]  (same scheme for k, v, out),
 * so both the reference's seq-first [L, N, E] projections and batch-first [N, L, E] buffers are
 * addressed without a transpose.
 *
 * Mask, exactly one of (all may be NULL = no mask):
 *   blocked_bits  packed bits [batch][Nq][words_per_row], bit (s & 31) of word (s >> 5) set = key s
 *                 may NOT be attended (the bool attn_mask of meanshiftformer_transformer_decoder.py:675-680,
 *                 stored once for all heads). row_open [batch][Nq] (int32, may be NULL): 0 = this row
 *                 blocks every key and is therefore treated as un-masked (decoder.py:618).
 *   add_mask      fp32 additive mask [G][Nq][Ns] (0 / -inf) as hypersphere_attention receives it.
 * flags: bit0 = L2-normalise q rows (eps 1e-12), bit1 = L2-normalise k rows. The reference attention
 *        uses both; the mean-shift update (mean_shift.py:90-109) uses neither.
 * den (may be NULL): [G][Nq] softmax denominators sum_s exp(kappa*(cos-1))*open, for msm_vmf_attention_weights.
 * ---------------------------------------------------------------------------------------------- */
#define MSM_VMF_NORMALIZE_Q 1
#define MSM_VMF_NORMALIZE_K 2
/* training: `den` is [2][G][Nq]; the second plane receives |softmax . v| (the norm removed by the final L2
 * normalisation), which msm_vmf_attention_bwd needs */
#define MSM_VMF_SAVE_NORM 4

size_t msm_vmf_attention_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd);

int msm_vmf_attention_fwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                          const float* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                          const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                          float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                          float* den,
                          const uint32_t* blocked_bits, int words_per_row, const int32_t* row_open,
                          const float* add_mask,
                          int batch, int heads, int Nq, int Ns, int hd, float kappa, int flags,
                          void* workspace, size_t workspace_bytes, void* stream);

/* attention weights [G][Nq][Ns] (second return value of hypersphere_attention, attention_util.py:75).
 * Needs `den` from msm_vmf_attention_fwd. Not on the hot path: the decoder layers discard it. */
int msm_vmf_attention_weights(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                              const float* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                              const float* den,
                              const uint32_t* blocked_bits, int words_per_row, const int32_t* row_open,
                              const float* add_mask, float* attn,
                              int batch, int heads, int Nq, int Ns, int hd, float kappa, int flags, void* stream);

/* Backward of the attention core (training). The reference differentiates
 * hypersphere_attention (attention_util.py:64-82) with torch.autograd; this entry point recomputes the weights
 * tile by tile instead of saving [G][Nq][Ns] tensors. `out` is the forward result, `den` the [2][G][Nq] buffer the
 * forward filled under MSM_VMF_SAVE_NORM, masks / flags / kappa exactly as passed to the forward. grad_q / grad_k /
 * grad_v are addressed like q / k / v, each with its own strides. Nq <= 128, hd <= 64. Deterministic. */
size_t msm_vmf_attention_bwd_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd);

int msm_vmf_attention_bwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                          const float* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                          const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                          const float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                          const float* grad_out, int64_t go_sb, int64_t go_sh, int64_t go_sl,
                          const float* den,
                          float* grad_q, int64_t gq_sb, int64_t gq_sh, int64_t gq_sl,
                          float* grad_k, int64_t gk_sb, int64_t gk_sh, int64_t gk_sl,
                          float* grad_v, int64_t gv_sb, int64_t gv_sh, int64_t gv_sl,
                          const uint32_t* blocked_bits, int words_per_row, const int32_t* row_open,
                          const float* add_mask,
                          int batch, int heads, int Nq, int Ns, int hd, float kappa, int flags,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Mask head.
 * msm_mask_logits replaces  torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
 *   (meanshiftformer_transformer_decoder.py:668 / :1020):
 *   masks[b][q][p] = sum_c embed[b][q][c] * feat[b][c][p],   p in [0, HW).
 * msm_mask_to_attn_bits replaces  F.interpolate(masks, size, "bilinear", align_corners=False)
 *   .sigmoid() ... < 0.5  (decoder.py:675-680) and the un-mask rule of decoder.py:618:
 *   writes bits [B][Q][ceil(Ht*Wt/32)] (1 = blocked) and row_open [B][Q] (1 = some key open).
 * ---------------------------------------------------------------------------------------------- */
int msm_mask_logits(const float* embed, const float* feat, float* masks,
                    int B, int Q, int C, int64_t HW, void* stream);

int msm_mask_to_attn_bits(const float* masks, uint32_t* bits, int32_t* row_open,
                          int B, int Q, int H, int W, int Ht, int Wt, void* stream);

/* msm_resample_bilinear_fwd replaces  F.interpolate(x, size=(Ht, Wt), mode="bilinear", align_corners=...)
 *   (align_corners = 0: the resampling of decoder.py:675, applied to the mask FEATURES by the inference path that
 *   skips the auxiliary full-resolution masks; align_corners = 1: upsample_bilinear of the embedding network,
 *   lib/networks/resnet_dilated.py:325): x [planes][H][W] -> y [planes][Ht][Wt], PyTorch's source-index rules. */
int msm_resample_bilinear_fwd(const float* x, float* y, int64_t planes, int H, int W, int Ht, int Wt, int align_corners,
                              void* stream);

/* msm_upsample_add_fwd replaces  cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear",
 *   align_corners=False)  (the FPN top-down step, pixel_decoder/msdeformattn.py:349-352):
 *   x [planes][H][W], add and y [planes][Ht][Wt]. */
int msm_upsample_add_fwd(const float* x, const float* add, float* y, int64_t planes, int H, int W, int Ht, int Wt,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layer  Y[M][N] = act( X[M][K] . W[N][K]^T + bias[N] ),  act: 0 = identity, 1 = ReLU.
 * Replaces the torch.nn.functional.linear calls of the hot path: the packed q/k/v in-projection
 *   ms_in_projection_packed (transformer_decoder/attention_util.py:84-140), FFNLayer / MLP
 *   (meanshiftformer_transformer_decoder.py:300-304, :329-341) and the projections + FFN of the deformable
 *   encoder (pixel_decoder/ops/modules/ms_deform_attn.py:96-124, pixel_decoder/msdeformattn.py:76-84).
 * The weight is converted ONCE (msm_linear_prepare_weight) into the layout the tensor-core kernel streams:
 *   bf16 [hi|lo][K/8][N][8]; `prepared` must hold msm_linear_weight_bytes(N, K) bytes, 128-byte aligned.
 * X rows start every ldx floats, Y rows every ldy floats (both multiples of 4, 16-byte aligned bases), so
 * column slices of wider buffers are addressed in place. N and K must be multiples of 32.
 * ---------------------------------------------------------------------------------------------- */
size_t msm_linear_weight_bytes(int N, int K);

int msm_linear_prepare_weight(const float* W, int64_t ldw, void* prepared, int N, int K, void* stream);

/* The prepared form of the TRANSPOSE: W fp32 [R][C] -> msm_linear_weight_bytes(C, R) bytes usable as the weight of
 * Y[M][C] = X[M][R] . (W^T)^T, i.e. the input gradient of a dense layer dX = dY . W (what torch.autograd computes for
 * F.linear in training) without a transposed copy. R and C multiples of 32. */
int msm_linear_prepare_weight_t(const float* W, int64_t ldw, void* prepared, int R, int C, void* stream);

int msm_linear_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias, float* Y, int64_t ldy,
                   int M, int N, int K, int act, void* stream);

/* Y = LayerNorm_N( residual + X . W^T + bias ) * gamma + beta  (biased variance, eps inside the square root, as
 * torch.nn.LayerNorm): the post-norm residual blocks of the deformable encoder, norm1(src + output_proj(...)) and
 * norm2(src + linear2(...)) (pixel_decoder/msdeformattn.py:64-84), with the add and the normalisation done in the
 * GEMM epilogue on the accumulator row. N must be 32 or 64 (one thread holds a whole output row). */
int msm_linear_ln_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias, const float* residual,
                      int64_t ldr, const float* gamma, const float* beta, float eps, float* Y, int64_t ldy, int M,
                      int N, int K, void* stream);

/* General fused form (every stage optional, applied in this order on each output row):
 *   v  = act( X . W^T + bias[n] + rowbias[row % rowbias_period][n] ) + residual[row][n]
 *   y  = LayerNorm(v; ln_gamma, ln_beta, ln_eps)
 *   z  = y / max(|y|_2, 1e-12)                     (l2_normalize; F.normalize)
 *   Y  = z ;   Y2 = LayerNorm(z; ln2_gamma, ln2_beta, ln2_eps)     (second output, optional)
 * One launch for the post-norm residual blocks of the decoder layers,  norm(tgt + out_proj(attn)),
 * norm(tgt + linear2(relu(linear1(tgt)))) followed by the block's F.normalize and the prediction heads'
 * decoder_norm (meanshiftformer_transformer_decoder.py:171-181, :245-260, :300-304, :637-638, :663), and - through
 * rowbias - for  in_proj(tgt + query_pos) = in_proj(tgt) + in_proj_nobias(query_pos)  with the second term a per-layer
 * [num_queries][N] table. Row stages (residual / LayerNorm / normalise / Y2) need N <= 256. */
int msm_linear_fused_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias, const float* rowbias,
                         int rowbias_period, int act, const float* residual, int64_t ldr, const float* ln_gamma,
                         const float* ln_beta, float ln_eps, int l2_normalize, const float* ln2_gamma,
                         const float* ln2_beta, float ln2_eps, float* Y2, int64_t ldy2, float* Y, int64_t ldy, int M,
                         int N, int K, void* stream);

/* Post-norm feed-forward block in one kernel:  Y = LayerNorm_D( X + relu(X . W1^T + b1) . W2^T + b2 ),
 * X, Y [M][D] (rows of ldx / ldy floats), W1 [F][D], W2 [D][F] both given in the prepared layout of
 * msm_linear_prepare_weight. Replaces MSDeformAttnTransformerEncoderLayer.forward_ffn
 * (pixel_decoder/msdeformattn.py:76-84); the [M][F] hidden activation stays in tensor memory.
 * D must be 32 or 64, F a multiple of 128 (F <= 1792 at D = 64: the bias vector lives in shared memory). */
int msm_ffn_ln_fwd(const float* X, int64_t ldx, const void* w1_prepared, const float* b1, const void* w2_prepared,
                   const float* b2, const float* gamma, const float* beta, float eps, float* Y, int64_t ldy, int M,
                   int D, int F, void* stream);

/* Row-wise tail of the decoder's post-norm residual blocks (meanshiftformer_transformer_decoder.py:181, :260, :304,
 * :637-638, :663) in one launch, rows of C contiguous floats (C <= 1024):
 *   v = x + y (y may be NULL);  o = LayerNorm(v; gamma, beta, eps);  o = o / max(|o|_2, 1e-12) if l2_normalize;
 *   out = o;  out2 = LayerNorm(o; gamma2, beta2, eps2) if out2 != NULL. */
int msm_add_layernorm_fwd(const float* x, const float* y, const float* gamma, const float* beta, float eps,
                          int l2_normalize, const float* gamma2, const float* beta2, float eps2, float* out,
                          float* out2, int rows, int C, void* stream);

/* 3x3 / stride 2 / padding 1 max-pool of a channels-last map x [B][H][W][C] -> y [B][(H-1)/2+1][(W-1)/2+1][C]
 * (C % 4 == 0): nn.MaxPool2d(3, 2, 1) of the ResNet stem (torchvision resnet.py; detectron2 BasicStem) on the
 * channels_last tensors the cuDNN backbone produces. Outside the head's path; kept here because ATen's kernel for
 * it costs 3 % of the whole-model step. */
int msm_maxpool3x3s2_nhwc_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream);

/* 1x1 convolution on NCHW input with the same kernel: X [B][K][HW] (pixels contiguous), weight prepared as above
 * from the conv weight viewed as [N][K]. y_nchw != 0: Y [B][N][HW] (what nn.Conv2d returns); y_nchw == 0:
 * Y [B][HW][N] (token-major, what the decoders consume after flatten(2).transpose(1,2)).
 * Replaces the kernel_size=1 Conv2d layers of the head: pixel-decoder input_proj / lateral / mask_features
 *   (pixel_decoder/msdeformattn.py:206-216, :258-262, :242) and the decoder's input_proj
 *   (meanshiftformer_transformer_decoder.py:498-499, :575). HW must be a multiple of 4. */
int msm_conv1x1_fwd(const float* X, const void* prepared, const float* bias, float* Y, int y_nchw, int B, int HW,
                    int N, int K, int act, void* stream);

/* The same for a channels-last map: X [B][HW][K] (the memory of a channels_last NCHW tensor, e.g. what a cuDNN
 * channels_last backbone hands to pixel_decoder/msdeformattn.py:258-262) -> Y [B][N][HW]. HW % 128 == 0. */
int msm_conv1x1_nhwc_fwd(const float* X, const void* prepared, const float* bias, float* Y, int B, int HW, int N, int K,
                         int act, void* stream);

/* 3x3 convolution, padding 1, stride 1, as an implicit GEMM over K' = 9*C on the same kernel (each tap is the
 * input tile's TMA box shifted by (kx, ky)). X is the input ALREADY zero-padded: [B][C][H+2][Wp], one zero row
 * above and below, one zero column on the left and Wp - W - 1 >= 1 on the right, Wp a multiple of 4;
 * Y is [B][N][H][W]. `prepared` comes from msm_linear_prepare_weight on the conv weight permuted to
 * [N][ky][kx][C] and viewed as [N][9*C]. Replaces SimpleBasePixelDecoder.mask_features (pixel_decoder/fpn.py:238-246,
 * 90.6 GFLOP per 640x480 image) and the FPN output conv layer_1 (pixel_decoder/msdeformattn.py:258-262). */
int msm_conv3x3_fwd(const float* X, const void* prepared, const float* bias, float* Y, int B, int C, int H, int W,
                    int Wp, int N, int act, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, forward.
 * Replaces MultiScaleDeformableAttention.ms_deform_attn_forward(value, spatial_shapes,
 *   level_start_index, sampling_loc, attn_weight, im2col_step)
 *   (pixel_decoder/ops/src/vision.cpp:18-21, ms_deform_attn.h:25-45, cuda/ms_deform_attn_cuda.cu:25-85,
 *    kernel cuda/ms_deform_im2col_cuda.cuh:242-304).
 * value [N][S][M][D]; spatial_shapes int64 [L][2] (H,W) and level_start_index int64 [L] on the DEVICE,
 * as the reference passes them; sampling_loc [N][Lq][M][L][P][2] (x,y in [0,1]); attn_weight
 * [N][Lq][M][L][P]; out [N][Lq][M*D]. im2col_step only chunks the batch in the reference and has no
 * numerical effect; it is accepted and ignored.
 * ---------------------------------------------------------------------------------------------- */
int msm_ms_deform_attn_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                           const float* sampling_loc, const float* attn_weight, float* out,
                           int N, int S, int M, int D, int L, int Lq, int P, int im2col_step, void* stream);

/* Fused sampling stage of MSDeformAttn.forward (pixel_decoder/ops/modules/ms_deform_attn.py:96-121): the
 * softmax over the L*P attention logits of each head, sampling_locations = reference_points[l] +
 * offsets / (W_l, H_l)  (reference_points with last dim 2, :103-106) and the op above, in one kernel.
 * offsets_logits: one row of ld_ol floats per (n, query): [M*L*P*2 raw sampling offsets (m,l,p,xy) |
 * M*L*P raw attention logits (m,l,p)], i.e. the outputs of the two projections side by side;
 * reference_points [N][Lq][L][2] (x, y in [0,1]). */
int msm_ms_deform_attn_fused_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                 const float* offsets_logits, int64_t ld_ol, const float* reference_points,
                                 float* out, int N, int S, int M, int D, int L, int Lq, int P, void* stream);

/* backward of the above (ms_deform_attn.h:47-67, cuda/ms_deform_attn_cuda.cu:88-158): grads are
 * ACCUMULATED into grad_value (caller zero-fills), grad_sampling_loc and grad_attn_weight are written. */
int msm_ms_deform_attn_bwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                           const float* sampling_loc, const float* attn_weight, const float* grad_out,
                           float* grad_value, float* grad_sampling_loc, float* grad_attn_weight,
                           int N, int S, int M, int D, int L, int Lq, int P, int im2col_step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * vMF mean-shift hill climbing.
 * Replaces seed_hill_climbing_ball(X, Z, kappa, max_iters, metric='cosine'),
 *   transformer_decoder/mean_shift.py:79-109 (= lib/utils/mean_shift.py:79-109), batched over images:
 *   repeat max_iters times:  Z <- unit( exp(kappa * Z X^T) X ).
 * X [B][n][d] unit rows, Z0 [B][m][d], Z_out [B][m][d] (may alias Z0).
 * ---------------------------------------------------------------------------------------------- */
size_t msm_mean_shift_workspace_bytes(int B, int n, int m, int d);

int msm_mean_shift_hill_climb(const float* X, const float* Z0, float* Z_out,
                              int B, int n, int m, int d, float kappa, int max_iters,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The rest of the classical clusterer (cosine metric), batched over images, no host round-trips.
 *
 * msm_select_smart_seeds replaces select_smart_seeds(X, num_seeds, return_selected_indices=True),
 *   transformer_decoder/mean_shift.py:128-189: seed 0 = X[first_index[b]] (the reference draws it with
 *   np.random.randint, :155; the caller supplies the draw), seed i+1 = the point with the largest distance
 *   0.5 (1 - x.s) to its nearest chosen seed, ties -> smallest index (torch.argmax).
 *   X [B][n][d] unit rows, first_index [B] int64 (device), seeds [B][num_seeds][d], selected [B][num_seeds] int64.
 *   d in {16, 32, 64, 128}. One cooperative launch per <= (co-resident CTAs) images.
 * msm_seed_connected_components replaces connected_components(Z, epsilon), mean_shift.py:41-76:
 *   seed_labels [B][m] int64, num_labels [B] int32 = number of distinct labels (len(torch.unique(.)), :218).
 * msm_assign_clusters replaces mean_shift.py:207-227: every point takes the label of its closest seed
 *   (first minimum, torch.argmin), then label 0 and the most populous label among 0..num_labels-1 swap.
 *   labels [B][n] int64.
 * ---------------------------------------------------------------------------------------------- */
size_t msm_smart_seeds_workspace_bytes(int B, int n, int num_seeds);

int msm_select_smart_seeds(const float* X, const int64_t* first_index, float* seeds, int64_t* selected,
                           int B, int n, int d, int num_seeds,
                           void* workspace, size_t workspace_bytes, void* stream);

int msm_seed_connected_components(const float* Z, int64_t* seed_labels, int32_t* num_labels,
                                  int B, int m, int d, float epsilon, void* stream);

size_t msm_assign_clusters_workspace_bytes(int B, int m);

int msm_assign_clusters(const float* X, const float* Z, const int64_t* seed_labels, const int32_t* num_labels,
                        int64_t* labels, int B, int n, int m, int d,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Eval-mode tail: mask upsample + instance_inference, for the kept queries only.
 * Replaces F.interpolate(pred_masks, image size, "bilinear", align_corners=False) followed per image by
 *   instance_inference(mask_cls, mask_pred): MSMFormer/meanshiftformer/pretrained_meanshiftformer_model.py:337-343,
 *   461-497 (= meanshiftformer_model.py:289-295, 414-450), with panoptic_on = False.
 * msm_instance_topk: scores = softmax(logits)[:, :-1]; the T best of the Q*(K1-1) scores per image.
 *   logits [B][Q][K1]; topk_query / topk_class int64 [B][T], topk_score [B][T]; order: descending score
 *   (the reference's topk(sorted=False) promises none).
 * msm_instance_masks: for kept query t of image b, the low-resolution logits mask_logits [B][Q][h][w] resampled to
 *   H x W:  pred_masks [B][T][H][W] = (m > 0) as 0/1 floats (NULL = not wanted); boxes [B][T][4] = (x0, y0, x1+1, y1+1) of the
 *   foreground, zeros when empty (detectron2 BitMasks.get_bounding_boxes); scores [B][T] =
 *   topk_score * sum(sigmoid(m) * mask) / (sum(mask) + 1e-6).
 * ---------------------------------------------------------------------------------------------- */
int msm_instance_topk(const float* logits, int64_t* topk_query, int64_t* topk_class, float* topk_score,
                      int B, int Q, int K1, int T, void* stream);

size_t msm_instance_masks_workspace_bytes(int B, int T, int H);

int msm_instance_masks(const float* mask_logits, const int64_t* topk_query, const float* topk_score,
                       float* pred_masks, float* boxes, float* scores,
                       int B, int Q, int h, int w, int T, int H, int W,
                       void* workspace, size_t workspace_bytes, void* stream);

/* get_confident_instances + combine_masks (lib/fcn/test_utils.py:35-52, 93-112) without materialising the masks:
 * instance_label [B][T] int32 = 2 + (confident instances before t) or -1 when t is dropped (topk_mode: class == 1 and
 * score > low_threshold when num_class >= 2, everything otherwise; else score > score_threshold); label_map [B][H][W]
 * = the label of the LAST kept instance whose upsampled logit is > 0 at the pixel, else 0. scores / classes are the
 * outputs of msm_instance_masks (pred_masks may be NULL there) / msm_instance_topk. T <= 64. */
int msm_instance_label_map(const float* mask_logits, const int64_t* topk_query, const float* scores,
                           const int64_t* classes, int32_t* instance_label, float* label_map,
                           int B, int Q, int h, int w, int T, int H, int W,
                           int topk_mode, int num_class, float score_threshold, float low_threshold, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Two-stage ("zoom-in") glue, lib/fcn/test_dataset.py:62-198. Label maps are fp32 images holding small
 * non-negative integer ids, as the reference passes them; ids outside [0, L) are ignored by the statistics.
 * msm_label_stats: per (image, id) int32 [N][L][6] = {pixels, pixels with depth_z > 0, W - xmin, H - ymin,
 *   xmax + 1, ymax + 1} (all 0 when the id is absent): the torch.unique / mask_to_tight_box / depth-fraction
 *   reductions of crop_rois :69,82-83 and filter_labels_depth :186-199 in one pass. depth_z = the z plane of
 *   image 0 (may be NULL), depth_stride = elements between consecutive images.
 * msm_relabel_lut: out[n][p] = lut[n][in[n][p] - lo] for ids in [lo, lo + L), else unchanged (:120-121, :197).
 * msm_crop_resize: every ROI at once (:97-113): rgb / depth crops [num][3][S][S] by bilinear align_corners=True,
 *   mask crops [num][S][S] = nearest resize of (labels == ids[c]); rois int32 [num][4] = x_min, y_min, x_max, y_max.
 * msm_crop_label_stats: per (crop, local id) int32 [num][L][4] = {pixels, pixels where init_crop != 0,
 *   pixels with depth z > 0, 0} and double [num][L] depth sums (:118-135).
 * msm_paste_crops: refined [H][W]: crops resized back (nearest) into their ROIs in `order`, later ones overwrite
 *   earlier ones where their relabelled value new_label[c][id] is non-zero (:151-180).
 * ---------------------------------------------------------------------------------------------- */
int msm_label_stats(const float* labels, const float* depth_z, int64_t depth_stride, int32_t* stats,
                    int N, int H, int W, int L, void* stream);

int msm_relabel_lut(const float* in, const float* lut, float* out, int N, int64_t per_image, int L, int lo,
                    void* stream);

int msm_crop_resize(const float* rgb, const float* depth, const float* labels, const int32_t* rois, const float* ids,
                    float* rgb_crops, float* depth_crops, float* mask_crops, int num, int H, int W, int S,
                    void* stream);

int msm_crop_label_stats(const float* labels_crop, const float* init_crop, const float* depth_crop, int32_t* stats,
                         double* depth_sum, int num, int S, int L, void* stream);

int msm_paste_crops(const float* labels_crop, const float* new_label, const int32_t* order, const int32_t* rois,
                    float* refined, int num, int H, int W, int S, int L, void* stream);

/* ----------------------------------------------------------------------------------------------
 * One decoder layer between its cross-attention and its mask head, as ONE kernel (csrc/decoder_block.cu): a cluster
 * of 8 CTAs per image, activations exchanged through distributed shared memory. Replaces, per layer,
 *   tgt = norm(tgt + out_proj(cross_attn))                       meanshiftformer_transformer_decoder.py:253-257
 *   MeanShiftSelfAttentionLayer.forward_post                     :171-181 (hypersphere attention 100 x 100 per head,
 *                                                                attention_util.py:64-82, in/out projections :121-140, :425)
 *   FFNLayer.forward_post                                        :300-304
 *   F.normalize(output, dim=-1)  (DECODER_BLOCK_NORM)            :637-638
 *   decoder_norm, class_embed, mask_embed MLP                    :661-664
 *   the NEXT layer's in_proj_q(tgt + query_pos)                  :250, attention_util.py:135-137
 * Fixed to the configuration every UOIS YAML selects: hidden 256, 8 heads, FFN 2048, post-norm, Q <= 128 queries.
 * wblob: msm_decoder_block_weight_bytes(with_qn) bytes, 128-byte aligned: for cluster rank r = 0..7 the fp16 hi | lo
 *   pieces ([hi|lo][4 k-groups of 8][N][8], one per 32 input channels, as msm_linear_prepare_weight) of, in order:
 *   cross out_proj rows [32r,32r+32); self in_proj rows q|k|v of head r (96); self out_proj rows; linear1 rows
 *   [256r,256r+128) and [256r+128,256r+256); linear2[0:128, 256r:256r+256] and linear2[128:256, 256r:256r+256];
 *   (with_qn) next layer's cross in_proj q rows; [mask_embed.layers.0 rows | class_embed padded to 32 rows] (64);
 *   mask_embed.layers.1 rows; mask_embed.layers.2 rows.   (host mirror: ops.decoder_block_pack)
 * t_qk [Q][768] = [query_pos . Wq^T | query_pos . Wk^T | 0] of the self-attention; t_qn [Q][256] = query_pos . Wq^T of
 *   the next layer's cross-attention (q_next, b_qn, t_qn all NULL for the last layer); b_c: 32 floats.
 * Outputs: state_out, embed, q_next [B][Q][256]; logits [B][Q][32] (columns >= num_classes + 1 are padding).
 * ---------------------------------------------------------------------------------------------- */
size_t msm_decoder_block_weight_bytes(int with_qn);

int msm_decoder_block_fwd(const float* o_cross, const float* state, const void* wblob, const float* b_o1,
                          const float* g1, const float* be1, float eps1, const float* b_qkv, const float* t_qk,
                          const float* b_o2, const float* g2, const float* be2, float eps2, const float* b_f1,
                          const float* b_f2, const float* g3, const float* be3, float eps3, int block_norm,
                          const float* gd, const float* bed, float epsd, const float* b_qn, const float* t_qn,
                          const float* b_m1, const float* b_c, const float* b_m2, const float* b_m3,
                          float* state_out, float* logits, float* embed, float* q_next, int B, int Q, float kappa,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSMFORMER_B200_H_ */
