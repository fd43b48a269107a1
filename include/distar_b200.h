/* distar_b200 — C-ABI of the B200-native AlphaStar policy hot path (libdistar_b200.so).
 *
 * The reference (opendilab/DI-star @ 12b1c69) is pure Python/PyTorch: it has no FFI of its own, so the
 * "binding" a maintainer adds is the ctypes stub in INTEGRATION.md (distar_b200/lib.py is that stub).
 * Every entry point below replaces a block of PyTorch-eager calls in the reference; the file:line it
 * replaces is cited per function (paths relative to distar/agent/default/).
 *
 * Conventions
 *  - plain device pointers + sizes; the caller (PyTorch) owns every buffer including workspaces;
 *    the library never allocates, frees or synchronises, and keeps no state beyond cached
 *    cudaFuncSetAttribute calls and TMA descriptors built per call on the host stack.
 *  - all work is enqueued on `stream` (CUDA-graph capturable unless noted).
 *  - return 0 on success, negative dsb_status on failure; dsb_last_error() gives a thread-local message.
 *  - tensors are dense row-major unless a stride argument says otherwise.
 */
#ifndef DISTAR_B200_H_
#define DISTAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dsb_stream_t; /* cudaStream_t */

enum dsb_status { DSB_OK = 0, DSB_ERR_ARG = -1, DSB_ERR_CUDA = -2, DSB_ERR_UNSUPPORTED = -3 };

const char* dsb_last_error(void);
int dsb_version(void);
/* number of kernels launched by this library in the calling process since load (bench gpu_launches). */
int64_t dsb_launch_count(void);

/* ---- scatter_connection  (model/module_utils.py:11-34 'add' + masking of model/encoder.py:37-38) ----
 * out[n,c,y,x] = sum_{e < entity_num[n], clamp(ey)=y, clamp(ex)=x} project[n,e,c]   (entity order, deterministic)
 * project [N,E,C=32] f32, ex/ey [N,E] u8, entity_num [N] i64 (NULL = all E valid), out [N,32,H,W] f32 (H=W=128).
 * Every output element is written exactly once (no memset pass). */
int dsb_scatter_connection_fwd(const float* project, const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num,
                               float* out, int N, int E, int H, int W, dsb_stream_t stream);
/* grad_project[n,e,c] = grad_out[n,c,clamp(ey),clamp(ex)] for e < entity_num[n], else 0. */
int dsb_scatter_connection_bwd(const float* grad_out, const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num,
                               float* grad_project, int N, int E, int H, int W, dsb_stream_t stream);

/* ---- entity feature expansion  (EntityEncoder.forward one-hot / binary / unsqueeze + cat, entity_encoder.py:59-78) ----
 * fields: host array of 36 device pointers (one per entity field, [tokens] each) with host tables kind (0 one-hot,
 * 1 11-bit binary MSB first, 2 scalar), offset (first column), vocab, dtype (0 u8, 1 i16, 2 i8, 3 f16).
 * Writes the 1024-wide (997 + zero padding) feature rows as the bf16 (hi, lo) pair of the embedding GEMM.
 * lo_col_base >= 0 selects the exact-operand layout: one-hot / binary columns are exact in bf16, and the bf16 residual of
 * the j-th scalar field is written to the spare hi column lo_col_base + j (the caller repeats that field's weight column
 * there), so the row needs no lo tensor (lo may be NULL) and the GEMM one product less (dsb_gemm_args.a_exact).
 * error_flag (device int, zeroed by the caller) is set to 1 if a one-hot id is negative (reference raises). */
int dsb_entity_features(const void* const* fields, const int* kind, const int* offset, const int* vocab, const int* dtype,
                        int num_fields, void* hi, void* lo, int lo_col_base, int64_t tokens, int* error_flag,
                        dsb_stream_t stream);

/* ---- fused spatial-encoder stem  (scatter_connection + plane expansion spatial_encoder.py:51-71 + project conv :72
 *      + first max_pool2d :75-79) ----
 * planes: host array of 7 device pointers (height_map, visibility_map, creep, player_relative, alerts, pathable,
 * buildable; each u8 [N,H,W]); effects: host array of 6 device pointers (int16 [N,100] flat pixel lists, zero padded);
 * project [N,E,32] f32 (masked scatter_project output), ex/ey u8 [N,E], entity_num i64 [N] (NULL = E),
 * weight [32,56], bias [32] of the 1x1 project conv; lut_workspace: 320 * 32 floats of scratch (the kernel rebuilds the combined
 * lookup table of the six categorical planes from the current weights there at every call).
 * out [N, H/2, W/2, out_c] f32 channels-last, out_c = 32 or 64 (channels >= 32 written as 0); out_hi/out_lo optional bf16
 * split of out.  Backward: grad_weight [32,56], grad_bias [32], grad_project [N,E,32] must be ZEROED by the caller; the
 * kernel accumulates into the first two and writes each valid entity row of the third. */
int dsb_spatial_stem_fwd(const void* const* planes, const void* const* effects, const float* project, const uint8_t* ex,
                         const uint8_t* ey, const int64_t* entity_num, const float* weight, const float* bias,
                         float* lut_workspace, float* out, void* out_hi, void* out_lo, int out_c, int N, int E, int H, int W,
                         dsb_stream_t stream);
int dsb_spatial_stem_bwd(const void* const* planes, const void* const* effects, const float* project, const uint8_t* ex,
                         const uint8_t* ey, const int64_t* entity_num, const float* weight, const float* bias,
                         float* lut_workspace, const float* grad_out, int out_c, float* grad_weight, float* grad_bias,
                         float* grad_project, int N, int E, int H, int W, dsb_stream_t stream);

/* ---- V-trace / UPGO / TD(lambda) return scans  (rl_training/as_rl_utils.py:157-218,265-312; call sites :15,:40,:238) ----
 * reward [F,T,B], value [F,T+1,B] (bootstrap row already zeroed where terminal, rl_loss.py:47-49),
 * rho [R,T,B] clipped importance ratios (R = 6 heads), gamma_td [F] (device), lambda_td scalar.
 * vtrace_adv [F,R,T,B]  (gamma=1, lambda=1, c=rho),  upgo_adv [R,T,B] = rho*(G - V[:T]) on field 0 (winloss),
 * td_return [F,T,B].  One thread per (column b, item); the T loop lives in registers. */
int dsb_return_scan(const float* reward, const float* value, const float* rho, const float* gamma_td, float lambda_td,
                    float* vtrace_adv, float* upgo_adv, float* td_return, int F, int R, int T, int B,
                    dsb_stream_t stream);

/* ---- per-row categorical statistics  (rl_loss.py:63-90; as_rl_utils.py:52-103; sl_loss.py CE) ----
 * For each of `rows` rows of `logits` [rows,C] (and optional `teacher` [rows,C]) with label action[rows] (i64):
 *   lse[r] = (row max, log sum exp(z - max)) as two floats, logp[r] = (logits[r,a]-max)-logsum,
 *   entropy[r] = -sum p log p, kl[r] = sum pt (log pt - log p), lse_t[r] likewise two floats for the teacher.
 * Masked classes carry -1e9 as in the reference.  teacher/kl/lse_t may be NULL.  mean_logp (optional) receives
 * mean_j log p_j, the smoothing term of LabelSmoothingCrossEntropy (sl_training/sl_loss.py:16-34).  A label outside [0, C)
 * (torch raises there) is clamped and bit 2 is OR-ed into error_flag (optional device int, zeroed by the caller). */
int dsb_categorical_stats_fwd(const float* logits, const float* teacher, const int64_t* action, float* lse,
                              float* logp, float* entropy, float* kl, float* lse_t, float* mean_logp, int* error_flag,
                              int64_t rows, int C, dsb_stream_t stream);
/* grad_logits[r,j] = g_logp[r]*(1[j==a]-p_j) - g_ent[r]*p_j*(log p_j + H_r) + g_kl[r]*(p_j - pt_j) + g_mean[r]*(1/C - p_j)
 * (g_* are per-row upstream gradients, each may be NULL; entropy = H_r from the forward). */
int dsb_categorical_stats_bwd(const float* logits, const float* teacher, const int64_t* action, const float* lse,
                              const float* entropy, const float* lse_t, const float* g_logp, const float* g_ent,
                              const float* g_kl, const float* g_mean, float* grad_logits, int64_t rows, int C,
                              dsb_stream_t stream);

/* ---- selected-units pointer network: one sampling step  (SelectedUnitsHead._query loop body, action_arg_head.py:267-306) ----
 * weights16: host array of 16 device pointers in this order: query_fc1 W,b; query_fc2 W,b; lstm W_ih, W_hh;
 * layernorm_i w,b; layernorm_h w,b; layernorm_c w,b; embed_fc1 W,b; embed_fc2 W,b.
 * Per row n (one CTA): query MLP, LN-LSTM(32) cell, logits over the S = E+1 key slots (slot entity_num = end token),
 * mask / temperature / softmax / argmax(p/q), end_flag + selected_units_num bookkeeping, running mean of the selected
 * keys and the embedding MLP -> ae = emb0 + embedding.  State tensors are updated in place: ae [N,1024], h/c [N,32],
 * mask u8 [N,S], ksum [N,32], count i32 [N], end_flag u8 [N], num i64 [N].  prev = previous step's result (step > 0).
 * q [N,S]: Exp(1) draws of this step.  all_ended (device int, preset to 1) is cleared by any row still selecting. */
int dsb_su_sample_step(const void* const* weights16, const float* emb0, float* ae, const float* key, float* h, float* c,
                       uint8_t* mask, float* ksum, int* count, uint8_t* end_flag, int64_t* num, const int64_t* entity_num,
                       const int64_t* prev, const float* q, float* logits_out, int64_t* result, int* all_ended, int N,
                       int S, int step, float temperature, dsb_stream_t stream);

/* ---- selected-units pointer network, teacher-forced (training) path  (SelectedUnitsHead._train_query,
 *      head/action_arg_head.py:168-216; csrc/su_train.cu) ----
 * kfull [P, E, ld] fp32: the stacked key projection, this head's 32 key columns first; end_embedding [32] is the learned end
 * token that occupies slot entity_num (:118-129); selected_units [P, su_ld] int64 labels, S <= 64 steps are processed.
 * su_prefix_mean: mean[p, i, :] = mean (selected_units_num[p] != 0) or sum of the keys of the distinct units labelled at steps
 *   <= i, the end token and everything after it excluded (:196-199); count [P, S] is kept for the backward, which ADDS the key
 *   gradients into grad_kfull (same [P, E, ld] layout, zeroed by the caller, shared with su_logits and the target-unit head).
 * su_lstm: the 32-wide LayerNorm-LSTM over the S steps from zero state; ig [P, S, 128] = LN_i(q W_ih^T); one warp per row.
 *   Saves gates / hg [P,S,128], pre_c [P,S,32], stats [P,S,4] = (mean_h, rstd_h, mean_c, rstd_c).  Backward returns d_ig and
 *   d_hg (gradient of the raw h W_hh^T: the caller forms dW_hh = d_hg^T h_prev) and ADDS the LayerNorm parameter gradients.
 * su_logits: logits[p, i, e] = h[p, i] . key_e for the E + 1 slots, -1e9 where the reference masks: slot > entity_num, slot
 *   chosen at an earlier step, end token at step 0 (:179-186).  Backward: grad_hs, key gradients ADDED into grad_kfull, the end
 *   token's gradient ADDED into grad_end_embedding [32]. */
int dsb_su_prefix_mean_fwd(const float* kfull, int ld, const int64_t* selected_units, int su_ld, const int64_t* entity_num,
                           const int64_t* selected_units_num, float* mean, int* count, int64_t P, int E, int S,
                           dsb_stream_t stream);
int dsb_su_prefix_mean_bwd(const float* grad_mean, const int64_t* selected_units, int su_ld, const int64_t* entity_num,
                           const int64_t* selected_units_num, const int* count, float* grad_kfull, int ld, int64_t P, int E, int S,
                           dsb_stream_t stream);
int dsb_su_lstm_fwd(const float* ig, const float* w_hh, const float* gamma_h, const float* beta_h, const float* gamma_c,
                    const float* beta_c, float* hs, float* cs, float* gates, float* hg, float* pre_c, float* stats, int64_t P,
                    int S, dsb_stream_t stream);
int dsb_su_lstm_bwd(const float* grad_hs, const float* w_hh, const float* gamma_h, const float* gamma_c, const float* beta_c,
                    const float* cs, const float* gates, const float* hg, const float* pre_c, const float* stats, float* d_ig,
                    float* d_hg, float* dgamma_h, float* dbeta_h, float* dgamma_c, float* dbeta_c, int64_t P, int S,
                    dsb_stream_t stream);
int dsb_su_logits_fwd(const float* hs, const float* kfull, int ld, const float* end_embedding, const int64_t* selected_units,
                      int su_ld, const int64_t* entity_num, float* logits, int64_t P, int E, int S, dsb_stream_t stream);
int dsb_su_logits_bwd(const float* grad_logits, const float* hs, const float* kfull, int ld, const float* end_embedding,
                      const int64_t* selected_units, int su_ld, const int64_t* entity_num, float* grad_hs, float* grad_kfull,
                      float* grad_end_embedding, int64_t P, int E, int S, dsb_stream_t stream);

/* ---- masked categorical sampling  (torch.multinomial(softmax(x),1) sites: head/action_type_head.py:57-58,
 *      head/action_arg_head.py:46-47,79-80,147-148,361-362,448-449) ----
 * index[r] = argmax_j softmax(logits[r])_j / q[r,j]  (first max wins), q ~ Exp(1) supplied by the caller so the
 * RNG stream is the reference's; also returns logp[r] = log_softmax(logits[r])[index[r]] (model.py:67-71). */
int dsb_sample_categorical(const float* logits, const float* q, int64_t* index, float* logp, int64_t rows, int C,
                           dsb_stream_t stream);

/* ---- bilinear x2 up-sampling, align_corners=False  (F.interpolate at head/action_arg_head.py:439-440) ----
 * in [NC, H, W] f32 -> out [NC, 2H, 2W]; backward is a deterministic gather (no atomics). */
int dsb_upsample_bilinear2x_fwd(const float* in, float* out, int64_t NC, int H, int W, dsb_stream_t stream);
int dsb_upsample_bilinear2x_bwd(const float* grad_out, float* grad_in, int64_t NC, int H, int W, dsb_stream_t stream);
/* channels-last variants: in [N, H, W, C] -> out [N, 2H, 2W, C] */
int dsb_upsample_bilinear2x_nhwc_fwd(const float* in, float* out, int64_t N, int H, int W, int C, dsb_stream_t stream);
int dsb_upsample_bilinear2x_nhwc_bwd(const float* grad_out, float* grad_in, int64_t N, int H, int W, int C,
                                     dsb_stream_t stream);

/* ---- location-head tail: conv3x3(upsample2x(x), one output channel)  (head/action_arg_head.py:436-443) ----
 * second half of the factorisation out = b + sum_tap shift_tap(upsample2x(z_tap)), z = x . w[C,9] at low resolution:
 * z [N, H, W, 9] -> out [N, 2H, 2W] (bias: 1 float, may be NULL); backward gz [N, H, W, 9] from grad_out. */
int dsb_upshift9_fwd(const float* z, const float* bias, float* out, int64_t N, int H, int W, dsb_stream_t stream);
int dsb_upshift9_bwd(const float* grad_out, float* grad_z, int64_t N, int H, int W, dsb_stream_t stream);

/* ---- location-head up-sampling stages: y = act(conv3x3(upsample_bilinear2x(x), w[C,Cin,3,3]) + b), C in {32, 64} ----
 * (head/action_arg_head.py:436-443).  z [N*H*W, ldz] holds the low-resolution projection z[p, tap*C + co] =
 * sum_ci w[co,ci,tap] x[p,ci] (one tensor-core GEMM); dsb_upconv_fwd applies the nine shifted bilinear up-samplings, bias
 * and ReLU and writes y [N,2H,2W,ldo] as fp32 and/or a bf16 (hi, lo) pair.  dsb_upconv_bwd is its transpose: from the
 * (ReLU-masked) g = dL/dy [N,2H,2W,ldg] it writes dL/dz as the (hi, lo) pair [N*H*W, ldz] (pad columns zeroed). */
int dsb_upconv_fwd(const float* z, int ldz, const float* bias, int relu, float* out, void* out_hi, void* out_lo, int ldo,
                   int64_t N, int H, int W, int C, dsb_stream_t stream);
int dsb_upconv_bwd(const float* g, int ldg, void* gz_hi, void* gz_lo, int ldz, int64_t N, int H, int W, int C,
                   dsb_stream_t stream);

/* ---- location-head tail projection: z[p, t] = sum_c x[p, c] w[c, t], C = 32 channels -> the 9 taps of the last conv ----
 * (the channel contraction of `upsample.2`, commuted in front of the up-sampling; see dsb_upshift9_*).  x [pixels, 32],
 * w [32, 9], z [pixels, 9] fp32.  Backward writes grad_x [pixels, 32] and ADDS the weight gradient into grad_w [32, 9]. */
int dsb_proj9_fwd(const float* x, const float* w, float* z, int64_t pixels, int C, dsb_stream_t stream);
int dsb_proj9_bwd(const float* x, const float* grad_z, const float* w, float* grad_x, float* grad_w, int64_t pixels, int C,
                  dsb_stream_t stream);

/* ---- 2x2 stride-2 max-pool, channels-last  (spatial_encoder.py:74-79 `F.max_pool2d` between the down-sampling convs) ----
 * x [N,H,W,C] fp32 -> out [N,H/2,W/2,C] (+ optional bf16 pair for the next convolution) and a one-byte argmax (0..3, window
 * scan order, first maximum) per output element; the backward scatters grad_out through it (grad_x fully written). */
int dsb_maxpool2_nhwc_fwd(const float* x, float* out, void* out_hi, void* out_lo, uint8_t* argmax, int64_t N, int H, int W, int C,
                          dsb_stream_t stream);
int dsb_maxpool2_nhwc_bwd(const float* grad_out, const uint8_t* argmax, float* grad_x, int64_t N, int H, int W, int C,
                          dsb_stream_t stream);

/* ---- GatedResBlock tail  (module_utils.py:228-229, location head K14) ----
 * out = relu(tanh(r * sigmoid(g)) * sp[0] + x) (+ skip: the next block's `x + map_skip`), fp32 [n] each, optionally also as
 * the bf16 (hi, lo) pair the next convolutions read.  Backward recomputes the gate from (r, g, x), writes grad_r / grad_g /
 * grad_x (grad of skip == grad_out) and ADDS the scalar gradient of sp into grad_sp[0]. */
int dsb_gate_update_fwd(const float* r, const float* g, const float* x, const float* skip, const float* sp, float* out,
                        void* out_hi, void* out_lo, int64_t n, dsb_stream_t stream);
int dsb_gate_update_bwd(const float* grad_out, const float* r, const float* g, const float* x, const float* sp, float* grad_r,
                        float* grad_g, float* grad_x, float* grad_sp, int64_t n, dsb_stream_t stream);

/* ---- column sums of a bf16 (hi, lo) pair [rows, N] ADDED into out [N]: the bias gradient of a layer whose masked gradient only
 * exists as the pair a dX GEMM epilogue emitted (dsb_gemm_args.relu_mask) ---- */
int dsb_colsum_pair(const void* hi, const void* lo, float* out, int64_t rows, int N, dsb_stream_t stream);

/* ---- fp32 -> (hi, lo) bf16 split used by the split-precision tensor-core GEMM ---- */
int dsb_split_bf16(const float* x, void* hi, void* lo, int64_t n, dsb_stream_t stream);

/* ---- ReLU backward + bf16 split + bias gradient in one pass ----
 * g = gy * (y > 0) (y NULL: g = gy; y may be fp32 or, with y_is_bf16, the bf16 hi half of the ReLU output); hi/lo receive the split of g, g_out (optional) g itself, colsum (optional)
 * per-block partial column sums [dsb_relu_bwd_split_blocks(rows, N), N] whose sum over blocks is the bias gradient.
 * With colsum_atomic != 0, colsum is instead a [N] vector (the bias gradient itself) that the column sums are atomically
 * added to.  hi/lo (together) and g_out may each be NULL when that output is not wanted. */
int dsb_relu_bwd_split_blocks(int64_t rows, int N);
int dsb_relu_bwd_split(const float* gy, const void* y, int y_is_bf16, float* g_out, void* hi, void* lo, float* colsum,
                       int colsum_atomic, int64_t rows, int N, dsb_stream_t stream);

/* ---- operand packing for small / odd-shaped / integer-typed activations (scalar_encoder.py:99-132 `.float()` inputs; head MLPs
 * whose K or N is not a tile multiple) ----
 * x [rows, K] (row pitch ld_in elements) of dtype 0 u8, 1 i16, 2 i8, 3 f16, 4 f32, 5 i64 -> bf16 hi (and lo = bf16(x - hi),
 * optional: 0/1 flags and counts <= 256 are exact in hi) [rows, Kp], columns >= K zero.  Kp % 8 == 0. */
int dsb_pack_pair(const void* x, int dtype, int64_t rows, int K, int64_t ld_in, void* hi, void* lo, int Kp,
                  dsb_stream_t stream);

/* ---- GLU gate  out = sigmoid(gate) * x  (module_utils.py:508-524), fp32 [n], n % 4 == 0 ---- */
int dsb_glu_gate_fwd(const float* gate, const float* x, float* out, int64_t n, dsb_stream_t stream);
int dsb_glu_gate_bwd(const float* grad_out, const float* gate, const float* x, float* grad_gate, float* grad_x, int64_t n,
                     dsb_stream_t stream);

/* ---- act(W . one_hot(idx) + b) as a gather  (action_type_head.py:61-63, action_arg_head.py:49-52,82-85: the action
 * embeddings fed back into the auto-regressive chain; scalar_encoder.py:105-116: nn.Embedding lookups) ----
 * out[p, j] = act(W[j * stride_out + idx[p] * stride_class] + bias[j]), j < N, idx in [0, C).  fc weight [N, C]: stride_out = C,
 * stride_class = 1; embedding table [C, N]: stride_out = 1, stride_class = N.  An id outside [0, C) is clamped; bit 4 of
 * error_flag (optional) records it unless clamp_max != 0 and the id is too LARGE (scalar_encoder.py:110-114 clamps those).
 * Backward ADDS into grad_W (same strides) and grad_bias [N] (optional) with atomics (rows of one class collide). */
int dsb_onehot_linear_fwd(const float* W, const float* bias, const int64_t* idx, float* out, int64_t P, int N, int C,
                          int64_t stride_out, int64_t stride_class, int relu, int clamp_max, int* error_flag,
                          dsb_stream_t stream);
int dsb_onehot_linear_bwd(const float* grad_out, const float* out, const int64_t* idx, float* grad_W, float* grad_bias,
                          int64_t P, int N, int C, int64_t stride_out, int64_t stride_class, int relu, dsb_stream_t stream);

/* ---- TargetUnitHead logits  (action_arg_head.py:343-363) ----
 * logits[p, e] = (e < entity_num[p] ? key[p, e, 0:32] . query[p, 0:32] : -1e9) / temperature; key rows have pitch ldk floats
 * (>= 32: the key projection shares its GEMM output with the selected-units head), E % 4 == 0.  One warp per 4 entities.
 * Backward: grad_query [P, 32] and grad_key [P, E, 32] (row pitch ldgk; masked entities get zeros). */
int dsb_target_unit_fwd(const float* key, int ldk, const float* query, const int64_t* entity_num, float* logits, int64_t P,
                        int E, float temperature, dsb_stream_t stream);
int dsb_target_unit_bwd(const float* grad_logits, const float* key, int ldk, const float* query, const int64_t* entity_num,
                        float* grad_key, int ldgk, float* grad_query, int64_t P, int E, float temperature, dsb_stream_t stream);

/* ---- beginning-build-order encoder pieces  (BeginningBuildOrderEncoder, obs_encoder/scalar_encoder.py:19-57: a 3-layer
 *      pre-LN transformer over 20 tokens of width 64, 2 heads of 8) ----
 * dsb_bo_tokens: token features one-hot(action, num_actions) | one-hot(position, L) | 10-bit x | 10-bit y of bo_location
 *   (MSB first, x = loc % spatial_x, y = loc / spatial_x) written as the exact bf16 A operand hi [B * L, Kp] of the embedding GEMM.
 * dsb_ln_small_*: LayerNorm over rows of D = 32 / 64 / 96 (one warp per row); backward ADDS dgamma / dbeta.
 * dsb_attn_small_*: unmasked softmax(Q K^T / sqrt(HD)) V for S <= 32 tokens, HD = 8 or 16; qkv [B, S, 3 * H * HD] (q | k | v,
 *   head-major), out [B, S, H * HD]; one warp per (sequence, head); the backward recomputes the probabilities. */
int dsb_bo_tokens(const int16_t* beginning_order, const int16_t* bo_location, int spatial_x, void* hi, int64_t B, int L,
                  int num_actions, int Kp, dsb_stream_t stream);
int dsb_ln_small_supported(int D);
int dsb_ln_small_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int64_t rows, int D, float eps,
                     dsb_stream_t stream);
int dsb_ln_small_bwd(const float* gy, const float* x, const float* gamma, const float* stats, float* gx, float* dgamma,
                     float* dbeta, int64_t rows, int D, dsb_stream_t stream);
int dsb_attn_small_fwd(const float* qkv, float* out, int64_t B, int S, int H, int HD, dsb_stream_t stream);
int dsb_attn_small_bwd(const float* qkv, const float* grad_out, float* grad_qkv, int64_t B, int S, int H, int HD,
                       dsb_stream_t stream);

/* ---- tcgen05 GEMM family  (fc_block nn_module.py:231-270; attention module_utils.py:88-111; their backward) ----
 * dsb_gemm_bf16_split:  C[M,N] = act( A[M,K] . W[N,K]^T + bias[N] ),  A and W as bf16 (hi, lo) pairs, K contiguous,
 * K % 64 == 0, N % 128 == 0, M arbitrary.  terms = 1: hi*hi only (plain bf16); terms = 3: hi*hi + hi*lo + lo*hi
 * (fp32-class product, ~2^-16 relative).  fp32 accumulation in TMEM.  C fp32 [M,N]; optional c_hi/c_lo (bf16 [M,N])
 * receive the split of the result so a following GEMM needs no separate split pass. */
int dsb_gemm_bf16_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                        float* c, void* c_hi, void* c_lo, int64_t M, int N, int K, int terms, int relu,
                        dsb_stream_t stream);

/* dsb_gemm_ex: the general form.  For batch b = bo*inner + bi and split s the kernel computes the [m, n] block
 *     C_b,s = alpha * sum_{k in split s} A_b[i,k] * B_b[j,k]   (+ bias[j], ReLU when splits == 1)
 * where the operands are sub-blocks of 2-D row-major bf16 tensors addressed by coordinate arithmetic only:
 *   K-major operand  (x_mn = 0): element (i, k) at row  row_outer*bo + row_inner*bi + i,  col col_base + col_inner*bi + k
 *   MN-major operand (x_mn = 1): element (i, k) at row  row_outer*bo + row_inner*bi + k,  col col_base + col_inner*bi + i
 * and C_b,s is written to rows c_row_outer*bo + c_row_inner*bi + c_row_split*s + [0, m), columns
 * c_col_base + c_col_inner*bi + [0, n) of the fp32 tensor c [c_rows, c_cols].
 * Examples: S = Q K^T for (obs, head): A = B = the QKV activation, a_col_base 0 / b_col_base 256, col_inner 128,
 * row_outer 512; O = P V: B = QKV with b_mn = 1; dW = dY^T X: both MN-major, splits > 1 over the token dimension.
 * Implicit-GEMM convolution (3x3 pad 1 or 1x1 over NHWC bf16 pairs, conv2d_block nn_module.py:119-174):
 *   a_conv = 1: A is the activation [conv_imgs, conv_h, conv_w, conv_c]; row i = output pixel, k = (tap, channel);
 *               the 3x3 halo is TMA zero fill.  B = weights [n = Cout, k = taps*C] (K-major), m = imgs*H*W.
 *   b_conv = 1: weight gradient: A = dY [pixels, Cout] (a_mn = 1), B = the activation read tap-shifted (b_mn = 1),
 *               n = taps*C, k = imgs*H*W pixels (split-K).
 * residual (optional fp32, same layout as c) is added before the ReLU.  bn: output tile width 64 or 128 (0 = auto).
 * Requirements: n % bn == 0, k % (64*splits) == 0, m % 128 == 0 when batch > 1, c_row_split % 128 == 0 when splits > 1. */
typedef struct dsb_gemm_args {
    const void *a_hi, *a_lo, *b_hi, *b_lo;   /* bf16 tensors (lo may be NULL when terms == 1 or the operand is flagged exact) */
    int64_t a_rows, a_cols, b_rows, b_cols;  /* their full 2-D shapes (cols contiguous) */
    int32_t a_mn, b_mn;
    int32_t a_col_base, a_col_inner, a_row_outer, a_row_inner;
    int32_t b_col_base, b_col_inner, b_row_outer, b_row_inner;
    const float* bias;
    float alpha;
    int32_t relu, terms;
    float* c;                                /* fp32 result; may be NULL when only the bf16 pair is wanted (no split-K then) */
    int64_t c_rows, c_cols;
    void *c_hi, *c_lo;                       /* optional bf16 (hi, lo) split of C, same [c_rows, c_cols] layout */
    int64_t m;
    int32_t n, k;
    int32_t batch, inner, splits;
    int32_t c_row_outer, c_row_inner, c_row_split, c_col_base, c_col_inner;
    const float* residual;                   /* optional fp32 [c_rows, c_cols] added before the ReLU */
    int32_t bn;
    int32_t a_conv, b_conv, conv_h, conv_w, conv_c, conv_taps;
    int64_t conv_imgs;
    int32_t c_accumulate;   /* != 0: C += result (TMA reduce-add; use with splits > 1 and c_row_split = 0 on a zeroed C) */
    int32_t mc;             /* 0 auto, 1 single CTAs, 2 clusters of two CTAs sharing B tiles by TMA multicast,
                               4 CTA pairs issuing cta_group::2 MMAs (M = 256 across two SMs) */
    int32_t a_exact;        /* terms == 3 only: A is exactly representable in bf16 (a_lo may be NULL): the a_lo x b_hi product
                               and the a_lo loads are skipped */
    int32_t b_exact;        /* same for B */
    const void* relu_mask;  /* optional bf16 [c_rows, c_cols] (the saved bf16 output of a ReLU layer): the result is zeroed where it
                               is 0 - the ReLU derivative applied in the epilogue of the dX GEMM that produces the gradient */
    float* colsum;          /* optional fp32 [c_cols]: the column sums of the (masked) result are ADDED here - the bias gradient of
                               the layer the gradient flows into, formed in the epilogue (warp butterfly + one atomic per column) */
} dsb_gemm_args;
int dsb_gemm_ex(const dsb_gemm_args* args, dsb_stream_t stream);

/* ---- masked softmax of the entity self-attention  (module_utils.py:99-107) ----
 * scores fp32 [rows, S] (already scaled by 1/sqrt(d) in the Q.K^T epilogue); row r belongs to observation
 * r / rows_per_obs; keys >= entity_num[obs] are masked to -1e9 as in the reference.  P leaves as a bf16 (hi, lo) pair.
 * Backward: dS = P * (dP - sum_k dP_k P_k) (0 at masked keys), again as a (hi, lo) pair for the dQ / dK products. */
int dsb_attn_softmax_fwd(const float* scores, const int64_t* entity_num, int rows_per_obs, void* p_hi, void* p_lo,
                         int64_t rows, int S, dsb_stream_t stream);
int dsb_attn_softmax_bwd(const void* p_hi, const void* p_lo, const float* dp, const int64_t* entity_num,
                         int rows_per_obs, void* ds_hi, void* ds_lo, int64_t rows, int S, dsb_stream_t stream);

/* ---- fused (residual +) LayerNorm  (module_utils.py:130-139; res_block.py:68-141; lstm.py:142-143) ----
 * y[r] = LN(x[r] (+ residual[r])) * gamma + beta over D = 128/256/384/512/1536 features, eps as nn.LayerNorm.
 * sum_out (optional, only with residual) receives x + residual (the LayerNorm input the backward needs);
 * y_hi / y_lo (optional) receive the bf16 split of y for a following tensor-core GEMM; stats [rows, 2] = (mean, rstd).
 * Backward: gx = dL/d(LN input); pgamma / pbeta are per-block partial sums [dsb_layernorm_bwd_blocks(rows), D], or with
 * atomic != 0 the [D] gradients of gamma / beta themselves, which the kernel adds to atomically. */
int dsb_layernorm_supported(int D);
int dsb_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* sum_out,
                      float* y, void* y_hi, void* y_lo, float* stats, int64_t rows, int D, float eps, dsb_stream_t stream);
int dsb_layernorm_bwd_blocks(int64_t rows);
int dsb_layernorm_bwd(const float* gy, const float* xin, const float* gamma, const float* stats, float* gx, float* pgamma,
                      float* pbeta, int atomic, int64_t rows, int D, dsb_stream_t stream);

/* ---- fused LayerNorm-LSTM cell  (LayerNormLSTMCell.forward, model/lstm.py:138-153, after the two matmuls) ----
 * ig [B,4H] = LN_i(x W_ih^T), hg [B,4H] = h W_hh^T (raw), c_in [B,H]; gate order in/forget/cell/out.
 * forward saves gates (pre-activation), stats_h, pre_c, stats_c for the backward; the backward returns d_ig, d_hg
 * (wrt the raw recurrent product), d_cin and ACCUMULATES the LayerNorm parameter gradients into dgamma_h / dbeta_h
 * [4H] and dgamma_c / dbeta_c [H] (caller zero-initialises).  H = 128 or 384. */
int dsb_lstm_cell_fwd(const float* ig, const float* hg, const float* c_in, const float* gamma_h, const float* beta_h,
                      const float* gamma_c, const float* beta_c, float* h_out, float* c_out, float* gates, float* stats_h,
                      float* pre_c, float* stats_c, int B, int H, float eps, dsb_stream_t stream);
int dsb_lstm_cell_bwd(const float* gh, const float* gcy, const float* gates, const float* hg, const float* stats_h,
                      const float* c_in, const float* pre_c, const float* stats_c, const float* gamma_h,
                      const float* gamma_c, const float* beta_c, float* d_ig, float* d_hg, float* d_cin, float* dgamma_h,
                      float* dbeta_h, float* dgamma_c, float* dbeta_c, int B, int H, dsb_stream_t stream);

/* ---- on-device learner-batch assembly from compact trajectories  (collate_fn / padding_entity_info,
 *      rl_training/rl_dataloader.py:45-76,206-245; csrc/batch_expand.cu) ----
 * dsb_expand_ragged: dst[r, s, e] = (s < steps[r] && e < width[r]) ? src[row_offset[r] + s * width[r] + e] : fill for r < rows,
 * s < S, e < W; steps may be NULL (one step per row).  elem_bytes 1 / 2 / 4 (4: float when fill_is_float, else int32).  Pads
 * entity fields with 0 to [rows, 512], the target-unit teacher logits with -1e9 to [rows, 512], the selected-units teacher
 * logits with -1e9 to [rows, 64, 513] and the selected-units behaviour log-probs / labels to [rows, 64].
 * dsb_sequence_mask: dst[r, j] = j < lengths[r] + add (uint8 0 / 1): selected_units_mask (add 0, W 64),
 * selected_units_logits_mask (add 1, W 513), target_units_logits_mask (add 0, W 512).
 * dsb_unpack_planes: one uint16 per pixel (bits [0,2) visibility_map, [2] creep, [3,6) player_relative, [6] alerts, [7] pathable,
 * [8] buildable) -> the six uint8 planes spatial_encoder.py:51-71 reads. */
int dsb_expand_ragged(const void* src, const int64_t* row_offset, const int* steps, const int* width, void* dst, int64_t rows,
                      int S, int W, int elem_bytes, double fill, int fill_is_float, dsb_stream_t stream);
int dsb_sequence_mask(const int64_t* lengths, int add, uint8_t* dst, int64_t rows, int W, dsb_stream_t stream);
int dsb_unpack_planes(const uint16_t* packed, uint8_t* visibility, uint8_t* creep, uint8_t* player_relative, uint8_t* alerts,
                      uint8_t* pathable, uint8_t* buildable, int64_t n, dsb_stream_t stream);

/* ---- persistent LayerNorm-LSTM layer: every time step of one layer in one launch each way  (LSTMLayer / LayerNormLSTMCell,
 *      model/lstm.py:138-167; csrc/lstm_seq.cu) ----
 * ig_all [L,B,4H] = LN_i(x W_ih^T) for all steps, (h0, c0) [B,H], w_hh_t [H,4H] = W_hh^T (forward) / w_hh [4H,H] (backward).
 * Forward writes hs, cs [L,B,H] and the tensors the backward needs: gates [L,B,4H] (pre-activation), hg [L,B,4H] (raw
 * h W_hh^T), stats_h / stats_c [L,B,2] (mean, rstd), pre_c [L,B,H].  Backward takes grad_hs [L,B,H] (may be NULL) and the
 * gradient of the last cell state (may be NULL), writes d_ig (gradient of ig_all), d_hg (gradient of the raw recurrent
 * product: the caller forms dW_hh = d_hg^T h_prev with one GEMM), dh0, dc0 and ADDS the LayerNorm parameter gradients into
 * dgamma_h / dbeta_h [4H], dgamma_c / dbeta_c [H].  H = 128 or 384. */
int dsb_lstm_seq_fwd(const float* ig_all, const float* h0, const float* c0, const float* w_hh_t, const float* gamma_h,
                     const float* beta_h, const float* gamma_c, const float* beta_c, float* hs, float* cs, float* gates, float* hg,
                     float* stats_h, float* pre_c, float* stats_c, int L, int B, int H, float eps, dsb_stream_t stream);
int dsb_lstm_seq_bwd(const float* grad_hs, const float* grad_c_last, const float* gates, const float* hg, const float* stats_h,
                     const float* pre_c, const float* stats_c, const float* cs, const float* c0, const float* w_hh,
                     const float* gamma_h, const float* gamma_c, const float* beta_c, float* d_ig, float* d_hg, float* dh0,
                     float* dc0, float* dgamma_h, float* dbeta_h, float* dgamma_c, float* dbeta_c, int L, int B, int H,
                     dsb_stream_t stream);

/* ---- fused grad-norm -> clip -> Adam over the flat arena  (rl_learner.py:73-80,125,132; grad_clip.py:141-144) ----
 * step 1: dsb_sumsq partial sums of grad^2 into `partial` [>= dsb_sumsq_partials()] then a finishing reduction
 *         into norm_out[0] = sqrt(sum) * scale (all on device, no host sync); scale = 1/world makes it the norm of the
 *         AVERAGED gradient, which is what the reference clips and logs (dist_helper.py:421-431, rl_learner.py:125).
 * step 2: dsb_adam_step reads norm_out on device: g = grad * grad_scale * min(1, max_norm/(norm + 1e-6)) (no clipping when
 *         norm is NULL or max_norm <= 0; grad_scale = 1/world folds the DP average), then torch.optim.Adam(beta1, beta2,
 *         eps, weight_decay: g += weight_decay * param, base_learner.py:164-168) with bias correction for step `t`
 *         (1-based); optionally refreshes the bf16 hi/lo shadow of the weights.  skip_flag (optional device float): when
 *         skip_flag[0] != 0 the launch leaves every buffer untouched - the data-parallel learner all-reduces one extra
 *         "my batch was invalid" slot together with the gradients, so a bad batch on ANY rank freezes the step on ALL ranks
 *         without a host round trip (the reference raises inside forward, entity_encoder.py:69-72). */
int dsb_sumsq_partials(void);
int dsb_grad_norm(const float* grad, int64_t n, float* partial, float* norm_out, float scale, dsb_stream_t stream);
int dsb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* norm,
                  float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int t, void* shadow_hi, void* shadow_lo, const float* skip_flag, dsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTAR_B200_H_ */
