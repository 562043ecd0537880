/*
 * satk.h — C ABI of libsatk.so: the B200 (sm_100a) kernels behind the teacher-forced training
 * hot path of Self-Attention Tacotron.
 *
 * The reference (nii-yamagishilab/self-attention-tacotron) is pure Python/TF1 and has NO native
 * seam; nothing in it dictates an FFI.  The entry points below are therefore
 * the operator boundary this implementation defines: one entry point per row group of SURVEY.md
 * §8(a), each comment naming the reference code it replaces (paths relative to /root/reference).
 * The host side (Python, `self-attention-tacotron_b200/engine.py`) mirrors the reference's
 * module interfaces (encoder / decoder / model_fn) and calls these with raw device pointers.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named `h_*`; the caller owns all buffers, the
 *     library allocates nothing;
 *   - tensors are row-major contiguous fp32 unless stated; masks are uint8 (1 = keep);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - return value: 0 on success, negative satk_status otherwise; `satk_last_error()` returns a
 *     thread-local message.  No exception crosses the ABI.  No global mutable state.
 */
#ifndef SATK_H_
#define SATK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SATK_OK = 0,
  SATK_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  SATK_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  SATK_ERR_UNSUPPORTED = -3  /* device is not sm_100 or resources do not fit */
} satk_status;

const char* satk_last_error(void);
int satk_version(void);
/* fills: [0]=SM count, [1]=cc major, [2]=cc minor, [3]=max co-resident 16-CTA clusters of the
 * attention-RNN kernel, [4]=max co-resident clusters of the LSTM kernel (H=256) */
int satk_device_info(int* out5);
/* sizeof() of the descriptor structs, in declaration order (gemm, lstm_fwd, lstm_bwd, attn_fwd, attn_bwd):
 * lets a foreign-language binding verify its struct layout */
int satk_struct_sizes(int* out5);
/* same for the decode-step descriptors (rowgemm, attn_step, sa_step, sa_tail, mlp_chain) */
int satk_struct_sizes_decode(int* out5);

/* ------------------------------------------------------------------------------------------
 * Dense tile: C = epilogue( alpha * sum_tap op(A_tap) * op(B_tap) ) (+ beta*C)
 * Replaces every tf.layers.Dense / tf.tensordot / tf.matmul / tf.layers.Conv1D on the path:
 * PreNet (module.py:394,426,1509), Conv1d banks and projections (module.py:46-68,78-83),
 * HighwayNet (module.py:72,91), LSTM input projections, attention memory layers
 * (forward_attention.py:59-64), MultiHeadAttention projections and QK^T / PV
 * (self_attention.py:45-65,108-128), Projection (module.py:639-643), and all their gradients.
 * ------------------------------------------------------------------------------------------ */
typedef enum { SATK_ACT_NONE = 0, SATK_ACT_RELU = 1, SATK_ACT_TANH = 2, SATK_ACT_SIGMOID = 3 } satk_act;

typedef struct {
  int M, N, K;
  int transA;              /* 0: A is [M,K] (lda = row stride); 1: A is [K,M] */
  int transB;              /* 0: B is [K,N]; 1: B is [N,K] */
  /* On the tcgen05 tile: (transA, transB) = (0, 1) — both operands contiguous along the reduction index — is the general form;
   * (1, 0) is the weight-gradient form C = A^T.B (X^T.dY, module.py's dense / conv layers differentiated): both operands are read
   * row-major as they are through MN-major shared-memory descriptors (no epilogue options, beta in {0, 1}, taps = 1; shift0 /
   * shift_per_batch1 / kshift0 / kshift_per_batch1 all shift A's reduction row, rows outside A read as zero).  Other combinations
   * and small shapes run on the SIMT tile. */
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  float alpha, beta;       /* beta applies to the existing C (0: overwrite) */
  const float* bias;       /* [N] or NULL, added before the activation */
  int act;                 /* satk_act */
  const float* residual;   /* [M,N] (ld = ldres) added AFTER the activation, or NULL */
  long long ldres;
  const uint8_t* keep_mask;/* [M,N] (ld = N) dropout keep mask applied after activation, or NULL */
  float keep_scale;        /* 1/keep_prob */
  /* batching: grid.z = batch1*batch2; pointer offsets z1*s?1 + z2*s?2 (elements) */
  int batch1, batch2;
  long long sA1, sA2, sB1, sB2, sC1, sC2;
  /* 1-D convolution as shifted GEMMs (tf Conv1D SAME, A.3).  taps>1 (or seq_len>0): the row index
   * of A is shifted by (shift0 + tap*tap_dir) WITHIN sequences of seq_len rows; rows shifted
   * outside their sequence read as zero.  For transA=0 the shift applies to the M index; for
   * transA=1 to the K index (weight gradient).  B advances by sBtap per tap (0 = same B). */
  int taps, shift0, tap_dir, seq_len;
  long long sBtap;
  int shift_per_batch1;    /* extra row shift z1*shift_per_batch1 (weight gradient of a conv: one tap per batch entry) */
  int split_k;             /* >1: partition K over grid.y slices and atomically add into C
                              (C must hold the value to accumulate onto; beta ignored; no epilogue
                              other than alpha) */
  int causal_skip;         /* 1: batched QK^T / PV with causal structure — skip tiles fully above the diagonal */
  /* weight gradient of a 1-D convolution on the tensor-core tile (operands already transposed, reduction index
   * contiguous): batch entry z1 reads A[m, k + kshift0 + z1*kshift_per_batch1] (zero outside [0,K)); one output per tap. */
  int kshift0, kshift_per_batch1;
  /* convolution BANK in one launch (module.py:46-53 and its input gradient): bank_widths = W > 0 makes z-batch entry w = 0..W-1 the
   * SAME-padded conv of width w+1 over the same A: taps = w+1, first tap at row shift -(w/2)*tap_dir, weights = entries
   * [w(w+1)/2 + tap] of B's tap dimension (stride sBtap: the W kernels stored back to back), A read at reduction offset
   * w*bank_a_kstep, result written (beta = 0) or added (beta = 1) at output column w*bank_c_nstep.  N is the width of one conv. */
  int bank_widths, bank_a_kstep, bank_c_nstep;
  /* batched products of the self-attention block (self_attention.py:45-65 and their gradients) on the tensor-core tile:
   * zcoord = 1 makes the batch1*batch2 entries share ONE 2-D view of each operand — A is [a_rows, a_cols] (row stride lda),
   * B is [b_rows, b_cols] (row stride ldb) — and entry z reads A at (m + z*za_row, k + z*za_k) and B at (n + z*zb_row, k + z*zb_k);
   * coordinates outside a view read as zero.  An operand is stored with the reduction index contiguous ([index rows, reduction
   * columns]: A with transA = 0, B with transB = 1) or with the reduction index as its ROW ([reduction rows, index columns]: transA = 1 /
   * transB = 0, read through MN-major descriptors — P.V and dS.K take V / K that way, P^T.dO and dS^T.Q both operands).
   * The result of entry z is slab z of a stacked output (sC1 = slab stride in elements, rows >= M / columns >= N clipped)
   * when sC1 != 0, otherwise columns [z*zc_col, z*zc_col + N) of one [M, c_cols] matrix.  Heads of a time-major
   * [T, B*D] activation are k-shifts (z*d_head) or, in its transpose, row shifts; stacked [z][T][T] score matrices are
   * row shifts (z*T) or, transposed as a whole, k-shifts.  No epilogue other than alpha; beta = 0.
   * causal_skip: 1 = skip tiles entirely above the diagonal (QK^T, dP), 2 = reduce over k < m0 + tile only (P.V, dS.K),
   * 3 = reduce over k >= m0 only (P^T.dO, dS^T.Q). */
  int zcoord, za_row, za_k, zb_row, zb_k, zc_col;
  long long a_rows, a_cols, b_rows, b_cols, c_cols;
  /* tensor-core tile only: 0 = "3xTF32" (hi/lo split of both operands, three MMAs per k-step: fp32-class accuracy, the
   * default everywhere), 1 = one TF32 pass on the raw fp32 operands (10-bit mantissas; 3x fewer MMAs, no split pass).
   * Which GEMM groups tolerate 1 inside the 1e-3 parity budget is measured by tools/ab_tf32.py (profiles/r02_ab_tf32.md). */
  int precision;
} satk_gemm_desc;

/* engine: 0 = auto, 1 = fp32 SIMT tile, 2 = tcgen05 3xTF32 tile (TMA-fed; falls back with an
 * error if the shape is not supported by the tensor-core path) */
int satk_gemm(const satk_gemm_desc* d, int engine, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-bound pieces
 * ------------------------------------------------------------------------------------------ */
/* Embedding lookup, models.py:283,351 (A.1): out[r,:] = table[ids[r]-offset,:] */
int satk_embedding_fwd(const long long* ids, int rows, int offset, const float* table, int dim, float* out, void* stream);
int satk_embedding_bwd(const long long* ids, int rows, int offset, const float* dout, int dim, float* dtable, void* stream);

/* Batch-norm statistics over rows (tf.layers.batch_normalization, training=True; A.3).
 * x [rows, C] with row stride ldx -> mean[C], var[C] (biased).  Optionally updates the moving
 * statistics in place: mov = mom*mov + (1-mom)*stat (variance Bessel-corrected if bessel!=0). */
int satk_bn_stats(const float* x, long long ldx, int rows, int C, float* mean, float* var,
                  float* mov_mean, float* mov_var, float momentum, int bessel, void* stream);
/* y = act(gamma*(x-mean)*rsqrt(var+eps)+beta) (+residual); optional fused MaxPooling1D(2,1,SAME)
 * along sequences of seq_len positions (module.py:54,80): y[t] = max(z[t], z[t+1]), y[T-1] = z[T-1].
 * Position of row r is (r / pos_stride) % seq_len; its neighbours are rows r +- pos_stride
 * (pos_stride = 1 for batch-major [B,T,C], = B for time-major [T,B,C]). */
int satk_bn_apply(const float* x, long long ldx, int rows, int C, const float* mean, const float* var,
                  const float* gamma, const float* beta, float eps, int act, const float* residual,
                  int maxpool_seq_len, int pos_stride, float* y, long long ldy, void* stream);
/* Backward of satk_bn_apply (+ optional maxpool, relu) for batch statistics (training) or moving
 * statistics (use_batch_stats=0).  dy [rows,C] -> dx [rows,C]; accumulates dgamma/dbeta (+=).
 * scratch: 2*C floats, zeroed by the call. */
int satk_bn_bwd(const float* x, long long ldx, int rows, int C, const float* mean, const float* var,
                const float* gamma, const float* beta, float eps, int act, int maxpool_seq_len, int pos_stride,
                int use_batch_stats, const float* dy, long long lddy, float* dx, long long lddx,
                float* dgamma, float* dbeta, float* scratch, void* stream);

/* HighwayNet combine (A.4; module.py:91): y = H*T + x*(1-T), H=relu(.), T=sigmoid(.) precomputed. */
int satk_highway_fwd(const float* H, const float* T, const float* x, float* y, long long n, void* stream);
/* dHpre = dy*T*(H>0); dTpre = dy*(H-x)*T*(1-T); dx = dy*(1-T) */
int satk_highway_bwd(const float* H, const float* T, const float* x, const float* dy,
                     float* dHpre, float* dTpre, float* dx, long long n, void* stream);

/* Pointwise backward through y = act(z) * keep * keep_scale (activation followed by dropout):
 * dz = dy * keep_scale * act'(y / keep_scale), zero where keep_mask == 0.  For relu the mask may be
 * NULL: a dropped unit has y == 0 and relu'(0) = 0 already zeroes it. */
int satk_act_bwd(const float* y, const float* dy, float* dz, long long n, int act, const uint8_t* keep_mask,
                 float keep_scale, void* stream);
/* dropout with an explicit keep mask (tf.layers.dropout behind a batch-normalised layer: the PostNetV2 convolutions,
 * models/models.py:92-100): y = keep_mask ? x * keep_scale : 0; also its own backward pass (dx from dy).  y may alias x. */
int satk_mask_scale(const float* x, const uint8_t* keep_mask, float keep_scale, float* y, long long n, void* stream);
/* column sums: out[c] += sum_r x[r,c]  (bias gradients) */
int satk_colsum_acc(const float* x, long long ldx, int rows, int C, float* out, void* stream);
int satk_add(const float* a, const float* b, float* out, long long n, void* stream);          /* out = a + b */
int satk_axpy(float alpha, const float* x, float* y, long long n, void* stream);              /* y += alpha*x */
int satk_transpose(const float* x, int rows, int cols, float* y, void* stream);               /* y[c,r] = x[r,c] */
/* strided transpose: y[c*ldy + r] = x[r*ldx + c]; feeds the weight-gradient products (X^T dY) to the tcgen05 tile, which
 * wants both operands contiguous along the reduction (row) dimension */
int satk_transpose_strided(const float* x, long long ldx, int rows, int cols, float* y, long long ldy, void* stream);
/* n transposes in one launch over two flat buffers with identical offsets: desc[3*i] = {offset, rows, cols} (device,
 * int32): dst[off + c*rows + r] = src[off + r*cols + c].  Keeps the K-contiguous copy of the weights that the
 * tcgen05 tile consumes ([in,out] -> [out,in]) in sync after each optimiser step. */
int satk_transpose_batched(const float* src, float* dst, const int* desc, int n, void* stream);
int satk_mask_rows(const float* x, const long long* lengths, int B, int T, int C, int time_major,
                   float* y, void* stream);                                                   /* zero rows t>=len[b] */
/* softsign(x) = x/(1+|x|) forward/backward (MultiSpeakerPreNet, multi_speaker_modules.py:22) */
int satk_softsign_fwd(const float* x, float* y, long long n, void* stream);
int satk_softsign_bwd(const float* x, const float* dy, float* dx, long long n, void* stream);
/* add a per-batch row vector to every time step: y[t,b,:] += v[b,:] (time-major) and its reduction */
int satk_add_rowvec_tb(float* y, const float* v, int T, int B, int C, void* stream);
int satk_sum_over_t(const float* dy, int T, int B, int C, float* dv, void* stream);
/* uint8 keep-mask generator: counter-based hash RNG (not TF's Philox stream; only the Bernoulli law matters) */
int satk_bernoulli_mask(uint8_t* out, long long n, float keep_prob, unsigned long long seed, void* stream);
/* the same mask as satk_bernoulli_mask(seed = seed_dev[0] * 1000003 + salt), the step's seed read from DEVICE memory: the launch can sit
 * inside a captured CUDA graph and still draw fresh masks on every replay (the host updates seed_dev[0] before it) */
int satk_bernoulli_mask_dev(uint8_t* out, long long n, float keep_prob, const unsigned long long* seed_dev, unsigned long long salt,
                            void* stream);

/* Row softmax with optional causal mask + dropout, self_attention.py:45-65,80-86.
 * S [rows_total = nmat*T, T] in place -> probabilities P (kept for backward / alignments);
 * Pd (may be NULL) = P * keep_mask * keep_scale. */
int satk_softmax_fwd(float* S, int nmat, int T, int causal, const uint8_t* keep_mask, float keep_scale, float* Pd, void* stream);
/* dS = P * (dP' - sum(dP' * P)),  dP' = dPd * keep_mask * keep_scale */
int satk_softmax_bwd(const float* P, const float* dPd, int nmat, int T, int causal, const uint8_t* keep_mask, float keep_scale,
                     float* dS, void* stream);

/* TransformerTrainingHelper (helpers.py:13-55,224-225): time-major decoder inputs
 * out[t,b,:] = mel[b, (t-1)*r + (r-n_feed) ... , :] flattened (zeros at t=0). */
int satk_teacher_inputs(const float* mel, int B, int Tm, int n_mels, int r, int n_feed, float* out, void* stream);

/* Losses (models.py:467-482; A.9) forward + gradient in one pass.
 * pred_tm: time-major mel prediction [Td,B,r*n_mels]; stop_tm [Td,B]; target batch-major.
 * out3 = {mel_loss, done_loss, loss}; dpred_tm/dstop_tm receive dLoss/d(pred). */
int satk_losses(const float* pred_tm, const float* stop_tm, const float* mel, const float* done,
                const float* spec_mask, const float* bin_mask, int B, int Tm, int n_mels, int r,
                float* out3, float* dpred_tm, float* dstop_tm, float* scratch4, void* stream);

/* Optimiser step on the flat buffer (models.py:485-498,595-598): global-norm clip (1.0),
 * Adam (eps outside the sqrt), gradient pre-scale (1/world_size after the all-reduce).
 * sumsq: scratch of SATK_SUMSQ_SCRATCH floats; [0] receives the sum of squares (deterministic: per-block partials are added
 * in a fixed order, so data-parallel replicas clip identically). */
#define SATK_SUMSQ_SCRATCH 640
int satk_grad_sumsq(const float* g, long long n, float* sumsq, void* stream);
int satk_adam_clip(float* p, const float* g, float* m, float* v, long long n, const float* sumsq, float grad_scale,
                   float clip_norm, float lr, float beta1, float beta2, float eps, int step, void* stream);

/* l2_regularization_loss (modules/regularizers.py:11-18, models/models.py:470-478) on the flat parameter buffer:
 * loss_acc[0] += scale * sum_i mask[i] * p[i]^2 / 2 (tf.nn.l2_loss) when loss_acc != NULL, and
 * g[i] += scale * mask[i] * p[i] when g != NULL.  mask[i] in {0, 1} selects the variables that are not black-listed. */
int satk_l2_reg(const float* p, const float* mask, long long n, float scale, float* g, float* loss_acc, void* stream);

/* ------------------------------------------------------------------------------------------
 * Recurrent core
 * ------------------------------------------------------------------------------------------ */
/* ZoneoutLSTMCell over a sequence (tacotron2 ZoneoutLSTMCell A.6 / TF LSTMCell A.5), used for the
 * encoder BiLSTM (module.py:93-110) and decoder LSTM-2/LSTM-3 (DecoderRNNV2, module.py:1525-1534).
 * The input projection (x.Wx+b) is dense over time and precomputed: xg [T,B,4H] (i,j,f,o).
 * One thread-block cluster of H/16 CTAs owns 4 batch rows; Wh lives in registers; h is exchanged
 * through distributed shared memory. */
typedef struct {
  int T, B, H;                 /* H in {128, 256} */
  int reverse;                 /* 1: backward direction of bidirectional_dynamic_rnn (length-reversed) */
  const float* xg;             /* [T,B,4H] */
  const float* Wh;             /* [H,4H]   rows of the TF kernel belonging to h */
  const long long* lengths;    /* [B] or NULL (all T) */
  const uint8_t* mask_c;       /* [T,B,H] indexed by PROCESSING step, or NULL -> eval interpolation */
  const uint8_t* mask_h;
  float zc, zh;
  float forget_bias;
  float* out;                  /* [T,B,ld_out] cell output (un-zoned h_new) in columns [0,H), zero past length */
  long long ld_out;            /* row stride of out (>= H): lets both BiLSTM directions write one [T,B,2H] tensor */
  /* saved for backward (position-indexed, may be NULL for inference): */
  float* gates;                /* [T,B,4H] post-nonlinearity i,j,f,o */
  float* c_prev;               /* [T,B,H] cell state BEFORE the step */
  float* h_prev;               /* [T,B,H] hidden state BEFORE the step */
} satk_lstm_fwd_desc;
int satk_lstm_seq_fwd(const satk_lstm_fwd_desc* d, void* stream);
/* developer aid: per-phase cycle averages of the last launch of kernel family `which` (0 lstm fwd, 1 attention-RNN fwd,
 * 2 attention-RNN bwd); only meaningful in -DSATK_PHASE_TIMING builds */
int satk_debug_phase_cycles(int which, long long* out16);

typedef struct {
  int T, B, H;
  int reverse;
  const float* Wh;
  const long long* lengths;
  const uint8_t* mask_c;
  const uint8_t* mask_h;
  float zc, zh;
  const float* gates;          /* saved by forward */
  const float* c_prev;
  const float* dout;           /* [T,B,ld_dout] gradient wrt cell output */
  long long ld_dout;
  float* dgates;               /* [T,B,4H] gradient wrt pre-activation gates (position-indexed, 0 past length) */
  /* optional [B] int32 (forward direction, lengths == NULL only): dout is exactly zero for the steps t >= step_end[b] of each
   * utterance (masked losses behind a causal decoder, see satk_attn_rnn_bwd_desc.step_end), so a cluster starts its walk at
   * the largest step_end of its rows and writes zeros to the dgates rows it skips.  NULL: all T steps. */
  const int* step_end;
} satk_lstm_bwd_desc;
int satk_lstm_seq_bwd(const satk_lstm_bwd_desc* d, void* stream);

/* Attention RNN: LSTM-1 + attention mechanism(s) for all Td steps in one launch.
 * Replaces the body of dynamic_decode's while_loop for layer 0 of DecoderRNNV2 under teacher
 * forcing: AttentionWrapper(ZoneoutLSTMCell, [ForwardAttention, BahdanauAttention])
 * (module.py:1011-1042,1514-1524; forward_attention.py:88-136; A.7,A.8).
 * One cluster of 16 CTAs owns 4 utterances for all Td steps: LSTM-1's recurrent kernel lives in
 * registers, the cluster's keys/values/query weights in shared memory, and the per-step exchange
 * (hidden state, partial energies, context) goes through distributed shared memory. */
typedef struct {
  int Td, B, Tt;
  int H;                       /* 256 */
  int A1, A2;                  /* score units; A2 = 0 for the single-attention model */
  int M1, M2;                  /* memory depths (256 / 32 or 0) */
  int att_kernel;              /* taps of the location convolution (<=32) */
  int mode;                    /* 0 additive, 1 location_sensitive, 2 forward */
  int cumulative;
  const float* xg;             /* [Td,B,4H] = prenet_out.W1x + b1 */
  const float* Wrec;           /* [ctx+H, 4H] rows of dec.lstm1.W after the pre-net rows */
  const uint8_t* mask_c;       /* [Td,B,H] or NULL */
  const uint8_t* mask_h;
  float zc, zh, forget_bias;
  const long long* lengths;    /* [B] source lengths */
  const float* keys1;          /* [Tt,B,A1] (time-major) memory_layer(values1) */
  const float* values1;        /* [Tt,B,M1] */
  const float* Wq1;            /* [H,A1] query_layer */
  const float* v1;             /* [A1] attention_variable / attention_v */
  const float* b1;             /* [A1] attention_bias (forward_attention.py:22) or NULL */
  const float* loc_conv_w;     /* [att_kernel, att_filters] location_features_convolution kernel (C_in = 1) */
  const float* loc_conv_b;     /* [att_filters] */
  const float* loc_layer_w;    /* [att_filters, A1] location_features_layer */
  int att_filters;             /* <= 8 */
  const float* keys2;          /* [Tt,B,A2] */
  const float* values2;        /* [Tt,B,M2] */
  const float* Wq2;            /* [H,A2] */
  const float* v2;             /* [A2] */
  /* outputs */
  float* x2;                   /* [Td,B,H+M1+M2] = concat(out1, ctx1, ctx2): input rows of LSTM-2 */
  float* align1;               /* [Td,B,Tt] alignments that build the context (alpha for forward attention) */
  float* align2;               /* [Td,B,Tt] or NULL */
  /* saved for backward (may be NULL): */
  float* gates;                /* [Td,B,4H] */
  float* c_prev;               /* [Td,B,H] */
  float* h_prev;               /* [Td,B,H] */
  float* soft1;                /* [Td,B,Tt] softmax alignments a_t (forward attention: state field 0) */
  float* q_save;               /* [Td,B,A1+A2] processed queries (query_layer outputs) */
  /* forward attention with the transition agent (forward_attention.py:111-114): u_t = sigmoid([ctx1_t, q1_t] . agent_w + agent_b)
   * replaces the constant 0.5 in the recursion of step t+1.  agent_w NULL: no agent. */
  const float* agent_w;        /* [M1+A1] transition_factor_projection kernel */
  const float* agent_b;        /* [1] */
  float* u_save;               /* [Td,B] transition factor USED by step t (u_{t-1}; 0.5 at t=0); saved for backward, may be NULL */
  float* state_final;          /* [B,Tt] location-attention state after the last step (sum of all alignments when cumulative);
                                  required by the backward pass when cumulative != 0, else may be NULL */
} satk_attn_rnn_fwd_desc;
int satk_attn_rnn_fwd(const satk_attn_rnn_fwd_desc* d, void* stream);

typedef struct {
  satk_attn_rnn_fwd_desc f;    /* same tensors as forward (its outputs are inputs here) */
  float* dx2;                  /* IN/OUT [Td,B,H+M1+M2].  in: gradient wrt x2 (from LSTM-2's input projection).
                                  out: columns [H:] hold the TOTAL gradient wrt (ctx1, ctx2) of each step
                                  (external + recurrent); the caller turns it into dvalues with one batched
                                  GEMM  dvalues[b] = align[:,b,:]^T . dctx[:,b,:] */
  float* dgates;               /* [Td,B,4H] gradient wrt LSTM-1 pre-activation gates */
  float* dq;                   /* [Td,B,A1+A2] gradient wrt processed queries (dense dWq afterwards) */
  float* dkeys1;               /* [Tt,B,A1]  (=) ; column sums give d(attention_bias) */
  float* dkeys2;               /* [Tt,B,A2]  (=) */
  float* dv1;                  /* [A1] (+=, atomics) */
  float* dv2;                  /* [A2] (+=) */
  float* dloc_conv_w;          /* [att_kernel, att_filters] (+=) */
  float* dloc_conv_b;          /* [att_filters] (+=) */
  float* dloc_layer_w;         /* [att_filters, A1] (+=) */
  float* dagent_w;             /* [M1+A1] (+=, atomics); required when f.agent_w is set */
  float* dagent_b;             /* [1] (+=) */
  /* optional [B] int32: number of decoder steps of each utterance that can carry a gradient (1 + the last step whose loss masks
   * are non-zero, models.py:467-482).  The caller guarantees that dx2 is exactly zero for the steps t >= step_end[b] (the losses
   * are masked there and every layer between the losses and x2 is causal in t), so their gradients are exactly zero: a cluster
   * starts its walk at the largest step_end of its utterances instead of Td-1 and writes zeros to the dgates / dq rows it
   * skips (dx2 is left as it is: zero).  Ignored with cumulative != 0 (the location state is walked back from the final state).
   * NULL: all Td steps. */
  const int* step_end;
  /* optional workspace of Td*B*(2*SATK_DE_ROW(Tt) + 8*Tt) floats (d(energies) [Td,B,2,SATK_DE_ROW(Tt)] - rows padded to whole
   * 128-byte lines -, then the location features [Td,B,Tt,8] that satk_attn_energy_grad computes once for all steps).  When it is set and the configuration is the dual-source decoder of the shipped
   * models (forward / location-sensitive first mechanism without cumulative weights or transition agent, <= 5 location
   * filters, Tt <= 192), the second-generation kernels run: the sequential kernel (one wave of 16-CTA clusters at B = 32)
   * only walks the recurrence and leaves d(energies) of both mechanisms here, and dkeys / dv / d(location layer / conv),
   * which do not feed the recurrence, come from a second, fully parallel launch over all (step, utterance) pairs
   * (forward_attention.py:13-26,98-100).  NULL: first-generation kernel, everything in one launch. */
  float* de_ws;
  /* with de_ws: SATK_EG_SYNC_INTS(B) ints (unit queue of the energy-gradient workers and the per-utterance progress flags the
   * recurrence publishes for them); no initialisation needed.  NULL: first-generation kernel. */
  int* sync_ws;
} satk_attn_rnn_bwd_desc;
#define SATK_DE_ROW(Tt) (((Tt) + 31) / 32 * 32)
#define SATK_EG_SYNC_INTS(B) (4 + 2 * (B))
int satk_attn_rnn_bwd(const satk_attn_rnn_bwd_desc* d, void* stream);
/* The two launches of the second-generation path separately (satk_attn_rnn_bwd issues both on one stream): the sequential
 * recurrence (fills dx2[:, H:], dgates, dq, de_ws) and the parallel energy gradients (reads de_ws; fills dkeys1/2 and adds
 * onto dv1/2, dloc_*), so that a caller can overlap the second one with whatever does not consume dkeys.
 * Both return SATK_ERR_UNSUPPORTED when the configuration is not covered (see de_ws). */
int satk_attn_rnn_bwd_recurrence(const satk_attn_rnn_bwd_desc* d, void* stream);
int satk_attn_energy_grad(const satk_attn_rnn_bwd_desc* d, void* stream);
/* The same in two parts: SATK_EG_FEATURES fills the location-feature half of de_ws (it reads only what the forward pass saved, so
 * it may be issued before / beside the recurrence), SATK_EG_GRADIENTS is the gradient launch proper (both mechanisms in one grid). */
#define SATK_EG_FEATURES 1
#define SATK_EG_GRADIENTS 2
int satk_attn_energy_grad_parts(const satk_attn_rnn_bwd_desc* d, int parts, void* stream);
/* Both launches as an overlapped pair on one stream: the gradient launch is a programmatic dependent of the recurrence (it starts
 * when every recurrence CTA is resident, on the SMs the clusters leave idle, and consumes d(energies) chunk by chunk behind the
 * recurrence's progress flags), so little of it is left when the recurrence ends.  features != 0: the location features are
 * computed first (0: the caller has issued SATK_EG_FEATURES already and ordered `stream` behind it). */
int satk_attn_rnn_bwd_overlapped(const satk_attn_rnn_bwd_desc* d, int features, void* stream);
/* `features` is a bit set: SATK_EG_FEATURES as above; SATK_EG_PREPARED: the caller has already issued satk_attn_energy_grad_prepare
 * (zeroing of the dkeys accumulators, the work queue and the progress flags) earlier on `stream` — hoisting those few memsets away from
 * the recurrence launch lets its clusters claim their SMs the moment the producer of dx2 ends. */
#define SATK_EG_PREPARED 4
int satk_attn_energy_grad_prepare(const satk_attn_rnn_bwd_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Free-running decoder step (PREDICT mode, predict_mel.py:36-74): the inference-branch cells of
 * rnn_wrappers.py:47-124,188-214 and module.py:762-778 with the decoder self-attention served from a
 * key/value cache instead of re-attending over the whole history every step.  Every kernel reads the
 * step index t from device memory (`t_ptr`), so one step can be captured in a CUDA graph and replayed.
 * ------------------------------------------------------------------------------------------ */
/* Skinny dense layer(s): C_i[M,N_i] = act_i(A[M,K] . W_i[K,N_i] + bias_i) (+ residual_i), up to 3 matrices sharing A.
 * W_i is the TF kernel layout [in,out].  `*_tstride` (elements) are multiplied by t = *t_ptr and added to the base
 * pointer, so histories / caches [Tmax, B, .] are addressed without host involvement. */
typedef struct {
  int M, K;
  const float* A; long long lda; long long a_tstride;
  const int* t_ptr;            /* device int (NULL: t = 0) */
  int nmat;                    /* 1..3 */
  const float* W[3];
  const float* bias[3];        /* or NULL */
  float* C[3]; long long ldc[3]; long long c_tstride[3];
  int N[3]; int act[3];
  const float* residual[3]; long long ldres[3]; long long res_tstride[3];   /* added after the activation, or NULL */
  /* double-buffered rows: A += (t & 1) * a_pstride, C_i += (t & 1) * c_pstride[i]  (0: single buffer).  The recurrent input
   * rows [x | h] are read by every CTA of a launch while the LSTM epilogue of the same launch produces the next h, so the
   * concatenated rows live in two buffers alternating with the parity of t. */
  long long a_pstride; long long c_pstride[3];
  /* lstm_H > 0: the single matrix is an LSTM kernel [K, 4H] (gate order i,j,f,o) and the epilogue is the ZoneoutLSTMCell
   * pointwise update in inference mode (A.5, A.6): c/h [M,H] are updated in place with (1-z)*new + z*old, the cell output
   * (un-zoned h) goes to lstm_out[(t&1)*out_pstride + row*ld_out + u] and the new h state to
   * lstm_hdst[((t+1)&1)*hdst_pstride + row*ld_hdst + u] (next step's input row); C / act / residual are unused. */
  int lstm_H;
  float* lstm_c; float* lstm_h;
  float zc, zh, forget_bias;
  float* lstm_out; long long ld_out; long long out_pstride;
  float* lstm_hdst; long long ld_hdst; long long hdst_pstride;
} satk_rowgemm_desc;
int satk_rowgemm(const satk_rowgemm_desc* d, void* stream);

/* One step of the attention mechanism(s) for every utterance (forward_attention.py:88-122,:13-26; TF BahdanauAttention
 * A.8; transition agent :111-114).  State (aprev, alpha, u) is updated in place; contexts [ctx1|ctx2] are written to up to
 * two destinations (the next LSTM-1 input row and the LSTM-2 input row).  A cluster of 8 CTAs serves one utterance. */
typedef struct {
  int B, Tt, A1, A2, M1, M2;
  int att_kernel, att_filters; /* 0 for additive attention */
  int mode;                    /* 0 additive, 1 location_sensitive, 2 forward */
  int cumulative, use_agent;
  const int* t_ptr;            /* device step index (row of align1/align2), NULL: 0 */
  const long long* lengths;    /* [B] */
  const float* q; long long ldq;               /* [B, A1+A2] processed queries (query_layer outputs); unused when Wq1 is set */
  /* optional fused query layers: q = q_x[b] . [Wq1 | Wq2] computed by the cluster (Wq1 NULL: read q).  q_x rows are double-buffered
   * on the parity of t like the recurrent input rows of satk_rowgemm. */
  const float* Wq1; const float* Wq2;          /* [q_in, A1], [q_in, A2] */
  const float* q_x; long long q_x_ld; long long q_x_pstride; int q_in;
  const float* keys1; const float* values1;    /* [Tt,B,A1], [Tt,B,M1] time-major */
  const float* v1; const float* b1;            /* [A1]; b1 may be NULL */
  const float* loc_conv_w; const float* loc_conv_b; const float* loc_layer_w;
  const float* keys2; const float* values2; const float* v2;
  const float* agent_w; const float* agent_b;  /* [M1+A1], [1] transition_factor_projection */
  float* aprev;                /* [B,Tt] previous (or cumulative) alignments */
  float* alpha;                /* [B,Tt] forward variable */
  float* u;                    /* [B] transition factor */
  float* ctx_dst0; long long ld0; long long pstride0;   /* written at parity (t+1)&1: LSTM-1 input row of the NEXT step */
  float* ctx_dst1; long long ld1; long long pstride1;   /* written at parity t&1: LSTM-2 input row of THIS step */
  float* align1;               /* [Tmax,B,Tt] or NULL */
  float* align2;
  /* forced-alignment mode (teacher_forcing_attention.py:13-78, models/models.py:411-427): when forced1 is set, row t of
   * forced1 [T,B,Tt] (and forced2 for the second mechanism) REPLACES the computed alignments of both mechanisms - no query, no energies,
   * no softmax, the recurrent attention state is left alone; contexts and the alignment history are produced as usual. */
  const float* forced1; const float* forced2;
} satk_attn_step_desc;
int satk_attn_step(const satk_attn_step_desc* d, void* stream);

/* Newest query against the cached keys / values of decoder steps 0..t (causal row t of self_attention.py:45-65). */
typedef struct {
  int B, D, heads, Tmax;
  const int* t_ptr;
  const float* q; long long ldq;   /* [B,D] */
  const float* Kc; const float* Vc;/* [Tmax,B,D]; rows 0..t valid */
  float* out; long long ldo;       /* [B,D] heads concatenated */
  float* probs;                    /* [B,heads,Tmax,Tmax] row t receives the alignment, or NULL */
} satk_sa_step_desc;
int satk_sa_step(const satk_sa_step_desc* d, void* stream);

/* Chain of up to three dense layers on one row per cluster (decoder pre-net of a free-running step, module.py:1509-1511,
 * multi_speaker_modules.py:27-32): h_{l+1} = act_l(h_l . W_l + bias_l) (+ residual_l[b]); widths <= 256.  Input row b is
 * x[t*x_tstride + b*x_ld ...], the last layer is written to out[(t&1)*out_pstride + b*out_ld ...]. */
typedef struct {
  int B, K0, nlayers;
  const int* t_ptr;
  const float* x; long long x_ld, x_tstride;
  const float* W[3]; const float* bias[3]; int N[3]; int act[3];
  const float* residual[3]; long long ldres[3];
  float* out; long long out_ld, out_pstride;
} satk_mlp_chain_desc;
int satk_mlp_chain(const satk_mlp_chain_desc* d, void* stream);

/* Fused tail of a decoder step, one cluster of 8 CTAs per utterance: for each self-attention hop the K/V/Q projections of the newest
 * decoder output (K, V appended to the caches), causal attention over rows 0..t, output projection, tanh transform + residual
 * (TransformerWrapper, rnn_wrappers.py:111-124; SelfAttentionTransformer.call, module.py:363-371), then the mel and stop projections
 * (OutputAndStopTokenTransparentWrapper, rnn_wrappers.py:188-214).  hops = 0: projections only (single-attention model). */
typedef struct {
  int B, D, heads, Tmax, hops;
  const int* t_ptr;
  const float* x; long long ldx;       /* [B, D] LSTM-3 output of this step */
  const float* Wk[4]; const float* bk[4]; const float* Wv[4]; const float* bv[4]; const float* Wq[4]; const float* bq[4];
  const float* Wo[4]; const float* bo[4]; const float* Wt[4]; const float* bt[4];     /* [D,D] kernels (TF layout), [D] biases */
  float* Kc[4]; float* Vc[4];          /* [Tmax,B,D] caches; row t is written here */
  float* probs[4];                     /* [B,heads,Tmax,Tmax] row t receives the alignment, or NULL */
  const float* W_out; const float* b_out; int n_out;    /* [D, n_out] mel projection */
  const float* W_stop; const float* b_stop;             /* [D, 1] */
  float* mel_dst; long long mel_tstride;                /* mel_dst[(t+1)*mel_tstride + b*n_out + c] (row 0 = go frame) */
  float* stop_dst;                                      /* [Tmax, B] */
  /* optional end-of-step bookkeeping (replaces satk_decode_tick): the last CTA to finish records the first finished step in
   * *done_step (see satk_decode_tick) and increments *tick_t; tick_counter is a zero-initialised device word.  NULL: disabled. */
  unsigned int* tick_counter; int* tick_t; int* done_step; int min_iters; int use_stop;
} satk_sa_tail_desc;
int satk_sa_tail(const satk_sa_tail_desc* d, void* stream);

/* End of a step: records the first step at which sigmoid(stop[t,b]) > 0.5 for all b and t > min_iters into *done_step
 * (initialised to -1 by the caller; stop may be NULL), then *t_ptr += 1. */
int satk_decode_tick(int* t_ptr, const float* stop, int B, int min_iters, int* done_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SATK_H_ */
