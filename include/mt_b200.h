/*
 * mt_b200.h -- C ABI of libmt_b200.so: the B200 (sm_100a) implementation of the fusion-model hot path of
 * frankaging/Multimodal-Transformer (per-modality Transformer encoder stacks + Memory-Fusion recurrence,
 * forward and backward).
 *
 * The reference has no FFI (it is pure PyTorch); each entry point below names the reference Python code it
 * replaces (paths relative to /root/reference/transformer/).  Conventions:
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns every buffer; nothing here
 *     allocates device memory.  Scratch is passed in as `ws` with the size returned by *_ws_bytes().
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - return value: 0 = ok, otherwise an MT_ERR_* code (mt_error_string() gives text).  Never throws/exits.
 *   - dtype MT_F32: activations and weights fp32 (FFMA GEMMs; the 1e-5 parity mode).
 *     dtype MT_BF16: GEMM operands bf16 (tcgen05/TMEM GEMMs fed by TMA), fp32 accumulation, fp32 residual
 *     stream / LayerNorm / softmax statistics / LSTM state.  Biases and LayerNorm parameters are always fp32.
 *   - parameter blocks are FLAT arrays in the canonical orders documented at each call; `params_lp` is the
 *     same layout in bf16 (only read when dtype == MT_BF16).  Gradients are fp32 in the same layout.
 */
#ifndef MT_B200_H
#define MT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_F32 0
#define MT_BF16 1

#define MT_OK 0
#define MT_ERR_ARG 1        /* bad shape / null pointer / unsupported size */
#define MT_ERR_ALIGN 2      /* pointer or leading dimension not aligned as required */
#define MT_ERR_CUDA 3       /* a CUDA runtime call failed (see mt_last_cuda_error) */
#define MT_ERR_WS 4         /* workspace too small */
#define MT_ERR_UNSUPPORTED 5

#define MT_ACT_NONE 0
#define MT_ACT_RELU 1
#define MT_ACT_TANH 2

#define MT_MAX_MODS 4

const char* mt_error_string(int code);
const char* mt_last_cuda_error(void);
/* library / device probe: returns MT_OK when device `dev` is an sm_100 part. */
int mt_check_device(int dev);
int mt_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t mt_launch_count(void);

/* Per-launch profiler (bench.py's roofline leg): after mt_prof_start every kernel launch of the library is followed
 * by an event record on its stream; record i's duration is the time since record i-1 (kernels of one stream
 * serialise).  name = "<host launch function>:<line>"; flops / bytes = the ALGORITHMIC work of that launch as annotated
 * at the launch site (0 when not annotated).  mt_prof_stop returns the number of records. */
int mt_prof_start(int max_records, void* stream);
int mt_prof_stop(void);
int mt_prof_get(int i, char* name, int name_cap, float* ms, double* flops, double* bytes);

/* ---------------------------------------------------------------------------------------------------
 * Linear:  y[M,N] = act(x[M,K] W[N,K]^T + b) [* rowmask[m]]        replaces nn.Linear call sites
 * MFT/multiTransformer.py:270,296 (embed), SFT/models.py:137-138 (fusionLayer + tanh), SFT/multiTransformer.py:
 * 431-433 (Dropout -> Linear -> ReLU; in_drop_p > 0 applies the input dropout).  x,y in `dtype` (x_f32 != 0: x is
 * fp32 and is cast on the fly -- the hot-path inputs arrive as fp32); W is ALWAYS the fp32 master weight; b fp32.
 * y_f32 != 0: y is written as fp32 even in bf16 mode (residual-stream input of an encoder stack).
 * ------------------------------------------------------------------------------------------------- */
int mt_linear_fwd(int dtype, int M, int N, int K, const void* x, int x_f32, const float* W, const float* b, void* y, int y_f32,
                  int act, const float* rowmask, float in_drop_p, uint64_t seed, uint32_t site, void* ws, size_t ws_bytes,
                  void* stream);
size_t mt_linear_ws_bytes(int dtype, int M, int N, int K, int x_f32, float in_drop_p);
/* backward: dy[M,N] (dtype, or fp32 when dy_f32), y = forward output (needed when act != NONE; fp32 when y_f32) ->
 * dx[M,K] (dtype; may be NULL), dW[N,K] fp32, db[N] fp32 (all overwritten). */
int mt_linear_bwd(int dtype, int M, int N, int K, const void* x, int x_f32, const float* W, const void* y, int y_f32, const void* dy,
                  int dy_f32, int act, const float* rowmask, float in_drop_p, uint64_t seed, uint32_t site, void* dx, float* dW,
                  float* db, void* ws, size_t ws_bytes, void* stream);
/* Modality concat in front of the early-fusion / embed Linear (SFT/models.py:136-138, B2-Trans/models.py:130-132 torch.cat(outputs, 2)):
 * dst[:, off_s : off_s + widths[s]] = srcs[s] ([M, widths[s]] contiguous, fp32 if src_f32[s] else bf16), converted to dst's dtype
 * (row stride ld_dst elements); mt_concat_bwd hands each modality its contiguous slice of the gradient.  dsrcs[s] may be NULL. */
int mt_concat_fwd(int M, int n_src, const void* const* srcs, const int* widths, const int* src_f32, void* dst, int ld_dst, int dst_f32,
                  void* stream);
int mt_concat_bwd(int M, int n_src, void* const* dsrcs, const int* widths, const int* dsrc_f32, const void* ddst, int ld, int ddst_f32,
                  void* stream);
size_t mt_linear_bwd_ws_bytes(int dtype, int M, int N, int K, int x_f32, float in_drop_p);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm (unbiased std, eps added to std):  MFT/multiTransformer.py:81-91.
 * x fp32 [M,d]; y in `dtype` (y_f32 forces fp32).  d % 128 == 0, d <= 1024.
 * ------------------------------------------------------------------------------------------------- */
int mt_layernorm_fwd(int dtype, int M, int d, const float* x, const float* a_2, const float* b_2, float eps, void* y,
                     int y_f32, void* stream);
/* dx = (dres ? dres : 0) + LN'(dy); da/db are ACCUMULATED into (caller zeroes). dy in `dtype` (dy_f32 forces fp32). */
int mt_layernorm_bwd(int dtype, int M, int d, const float* x, const float* a_2, float eps, const void* dy, int dy_f32,
                     const float* dres, float* dx, float* da, float* db, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused multi-head self-attention core (scores, query-row mask, softmax, dropout, PV, head merge):
 * attention() MFT/multiTransformer.py:22-34 + the head split/merge of MultiHeadedAttention.forward :54-64.
 * qkv [B,T,3d] (q | k | v along the last dim, head hd at columns hd*dk), mask fp32 [B,T] (0 => the whole
 * query row is filled with -1e9, i.e. uniform attention over ALL T keys), out [B,T,d].
 * lse fp32 [B,h,T] is written when non-NULL (needed by backward).  dk = d/h in {16,32,64,128}.
 * ------------------------------------------------------------------------------------------------- */
int mt_attention_fwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse,
                     float p_drop, uint64_t seed, uint32_t site, void* stream);
/* Ragged inference variant (no dropout, no lse): keys j >= key_len[b] (device int [B]) are excluded for narrative b; see
 * MtEncoderCfg.key_len. */
int mt_attention_ragged_fwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, const int* key_len, void* out,
                            void* stream);
/* dqkv [B,T,3d] overwritten.  ws: mt_attention_bwd_ws_bytes (B*h*T floats). */
int mt_attention_bwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, const void* out,
                     const float* lse, const void* dout, void* dqkv, float p_drop, uint64_t seed, uint32_t site,
                     void* ws, size_t ws_bytes, void* stream);
size_t mt_attention_bwd_ws_bytes(int B, int T, int h);
/* test hook: route bf16 attention through the FFMA engine (A/B the tensor-core engine); returns the previous setting. */
int mt_attention_force_ffma(int on);
/* test hook: route bf16 attention with T <= 128 through the tiled any-T tensor-core kernel instead of the whole-head one. */
int mt_attention_force_tiled(int on);
/* tcgen05 / TMEM engine for T <= 128, 32-wide heads, an even number of heads (bf16 only; mt_attention_fwd / _bwd pick it
 * automatically).  Direct entries for tests and probes: MT_ERR_UNSUPPORTED outside that envelope.  key_len may be NULL.
 * mt_attention_tc_bwd: aux = workspace of mt_attention_tc_bwd_ws_bytes; dbias (optional, fp32 [3d], ACCUMULATED) receives the
 * column sums of dqkv (the QKV projection's bias gradient). */
int mt_attention_tc_fwd(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, float p_drop,
                        uint64_t seed, uint32_t site, const int* key_len, void* stream);
int mt_attention_tc_bwd(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                        const void* dout, void* dqkv, float p_drop, uint64_t seed, uint32_t site, float* dbias, void* ws,
                        size_t ws_bytes, void* stream);
size_t mt_attention_tc_bwd_ws_bytes(int B, int T, int h);
/* test hook: disable the tcgen05 engine (A/B against the mma.sync kernels); returns the previous setting. */
int mt_attention_force_no_tc(int on);
/* materialise p_attn [B,h,T,T] fp32 (the reference keeps it as MultiHeadedAttention.attn, :59); debug/inspection. */
int mt_attention_probs(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, float* probs,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Encoder stack: Encoder.forward MFT/multiTransformer.py:73-76 over N EncoderLayer (:106-116) =
 * SublayerConnection (:103-104) around MultiHeadedAttention (:47-65) and PositionwiseFeedForward (:19-20),
 * then the final LayerNorm.
 * Flat parameter layout (floats), per layer l = 0..N-1:
 *   w_qkv[3d*d] (linears.0/1/2 weights stacked on the output dim), b_qkv[3d], w_o[d*d], b_o[d],
 *   w_1[dff*d], b_1[dff], w_2[d*dff], b_2[d], ln1_a[d], ln1_b[d], ln2_a[d], ln2_b[d];
 * then ln_f_a[d], ln_f_b[d].                                   mt_encoder_param_count() floats in total.
 * x fp32 [B,T,d] (residual stream in), mask fp32 [B,T], y [B,T,d] in `dtype` (y_f32 forces fp32).
 * training != 0 keeps the activations backward needs inside `ws` (pass the same ws to mt_encoder_bwd).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, T, d, h, dff, n_layers;
  int dtype;
  int training;      /* keep the activation stash backward needs (independent of dropout) */
  float p_drop;      /* dropout probability applied in this call; pass 0 in eval mode */
  uint64_t seed;
  int stack_id;      /* selects the dropout sites: site = (stack_id*64 + layer)*8 + k */
  int y_f32;
  int grid_share;    /* 0/1: this stack owns the GPU; s > 1: it runs concurrently with s - 1 other streams (the other modality stacks):
                        its GEMM / LayerNorm kernels launch 1/s of the resident CTA slots so the stacks co-reside on the SMs */
  const int* key_len; /* NULL, or device int [B] for RAGGED INFERENCE (mt_encoder_fwd with training == 0, p_drop == 0): narrative b only
                        has its first key_len[b] windows, so attention keys beyond them are excluded and the first key_len[b] output
                        rows equal a forward of that narrative alone (the reference evaluates with batch_size = 1, MFT/train.py:169,218,
                        where no padded window exists; in a padded TRAINING batch padded windows are live keys, appendix A.12) */
} MtEncoderCfg;

size_t mt_encoder_param_count(int d, int dff, int n_layers);
/* mt_encoder_stack_fwd / _bwd (the names proposed in SURVEY 8(b)) are aliases of mt_encoder_fwd / _bwd. */
size_t mt_encoder_ws_bytes(const MtEncoderCfg* cfg);
int mt_encoder_fwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x,
                   const float* mask, void* y, void* ws, size_t ws_bytes, void* stream);
/* dy [B,T,d] (dtype / fp32 per cfg->y_f32) -> dx fp32 [B,T,d]; grads (fp32, flat layout) are OVERWRITTEN. */
int mt_encoder_bwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x,
                   const float* mask, const void* dy, float* dx, float* grads, void* ws, size_t ws_bytes,
                   void* stream);
/* Grouped form: n_stacks (<= 4) modality stacks of IDENTICAL shape in one call (the three per-modality encoders of MultiTransformer,
 * MFT/multiTransformer.py:278-299, which the reference runs one after the other).  x / y / dy / dx are [n_stacks, B, T, d] (stack g at
 * rows g*B*T), the parameter block of stack g starts at params + g*param_stride floats (params_lp + g*param_stride bf16 elements; gradients
 * likewise), param_stride % 8 == 0.  seeds[g] / stack_ids[g] select stack g's dropout streams exactly as cfg->seed / cfg->stack_id do for
 * a single stack (NULL: cfg->seed for all, cfg->stack_id + g), so the result equals n_stacks single-stack calls bit for bit in fp32 mode and
 * up to the GEMM engine's summation order in bf16 mode.  One launch per projection / LayerNorm / weight gradient serves all stacks. */
size_t mt_encoder_group_ws_bytes(const MtEncoderCfg* cfg, int n_stacks);
int mt_encoder_group_fwd(const MtEncoderCfg* cfg, int n_stacks, const uint64_t* seeds, const int* stack_ids, const float* params,
                         const void* params_lp, size_t param_stride, const float* x, const float* mask, void* y, void* ws, size_t ws_bytes,
                         void* stream);
int mt_encoder_group_bwd(const MtEncoderCfg* cfg, int n_stacks, const uint64_t* seeds, const int* stack_ids, const float* params,
                         const void* params_lp, size_t param_stride, const float* x, const float* mask, const void* dy, float* dx, float* grads,
                         void* ws, size_t ws_bytes, void* stream);
int mt_encoder_stack_fwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x,
                         const float* mask, void* y, void* ws, size_t ws_bytes, void* stream);
int mt_encoder_stack_bwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x,
                         const float* mask, const void* dy, float* dx, float* grads, void* ws, size_t ws_bytes,
                         void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Memory Fusion Network: MFN.forward MFT/multiTransformer.py:181-248 (LSTHM cells :208, delta-memory
 * attention over [c_{t-1} || c_t] :210-219, cHat :220, gamma gates + memory update :221-224, output
 * head :237-248) with the x-projection hoisted into one GEMM, the recurrence in one persistent kernel and
 * the head batched over all steps.
 * Flat parameter layout: for each modality m (in order): w_ih[4H_m*D_m], w_hh[4H_m*H_m], b_ih[4H_m],
 * b_hh[4H_m]; then (W,b) of att1_fc1 [A1,2H], att1_fc2 [2H,A1], att2_fc1 [A2,2H], att2_fc2 [MEM,A2],
 * gamma1_fc1 [G,2H+MEM], gamma1_fc2 [MEM,G], gamma2_fc1, gamma2_fc2, out_fc1 [O,H+MEM], out_fc2 [1,O].
 * Inputs x[m]: [B,T,D_m] in `dtype` addressed as x[m] + b*stride_b[m] + t*stride_t[m] (elements), last dim
 * contiguous -- so the reference's [T,B,D] permuted views need no copy.  out fp32 [B,T] = head * mask.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, T, n_mods;
  int in_dim[MT_MAX_MODS];
  int hid[MT_MAX_MODS];
  int mem_dim, h_att1, h_att2, h_gamma, h_out;
  int dtype;
  int training;           /* keep the stash backward needs */
  float p_gamma, p_out;   /* dropout probs applied in this call (0.2, 0.5 in train mode; pass 0 in eval) */
  uint64_t seed;
  int bwd_phase;          /* mt_mfn_bwd only: 0 = whole backward; 1 = everything except the batched weight / bias gradients (dx and the
                             reverse-time recurrences: what upstream layers wait for); 2 = only those weight / bias gradients, which
                             nothing downstream depends on -- a caller may enqueue phase 2 on another stream once phase 1 is enqueued */
} MtMfnCfg;

size_t mt_mfn_param_count(const MtMfnCfg* cfg);
size_t mt_mfn_ws_bytes(const MtMfnCfg* cfg);
int mt_mfn_fwd(const MtMfnCfg* cfg, const float* params, const void* params_lp, const void* const* x,
               const int64_t* stride_b, const int64_t* stride_t, const float* mask, float* out, float* h_last,
               float* c_last, float* mem_last, void* ws, size_t ws_bytes, void* stream);
/* test hook: run the bf16 recurrences on the FFMA kernels instead of the tensor-core ones; returns the previous setting. */
int mt_mfn_force_ffma(int on);
/* dout fp32 [B,T] -> dx[m] [B,T,D_m] (dtype, contiguous), grads fp32 flat (OVERWRITTEN). */
int mt_mfn_bwd(const MtMfnCfg* cfg, const float* params, const void* params_lp, const void* const* x,
               const int64_t* stride_b, const int64_t* stride_t, const float* mask, const float* dout, void* const* dx,
               float* grads, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Step-wise LSTM decoder with output feedback + MLP head: SFT/multiTransformer.py:465-483 (NLPTransformer)
 * and MFT/multiTransformer.py:357-375 (UniTransformer).
 * Flat layout: w_ih[4E*2E], w_hh[4E*E], b_ih[4E], b_hh[4E], h0[E], c0[E], w_out0[Hd*E], b_out0[Hd], w_out2[Hd], b_out2[1].
 * enc [B,T,E] in `dtype`; out fp32 [B,T] = head * mask.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, T, E, Hd;
  int dtype;
  int training;
} MtLstmHeadCfg;
size_t mt_lstm_head_param_count(const MtLstmHeadCfg* cfg);
size_t mt_lstm_head_ws_bytes(const MtLstmHeadCfg* cfg);
int mt_lstm_head_fwd(const MtLstmHeadCfg* cfg, const float* params, const void* params_lp, const void* enc,
                     const float* mask, float* out, void* ws, size_t ws_bytes, void* stream);
int mt_lstm_head_bwd(const MtLstmHeadCfg* cfg, const float* params, const void* params_lp, const void* enc,
                     const float* mask, const float* dout, void* denc, float* grads, void* ws, size_t ws_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Window front-end: what MultiCNNTransformer.forward does per modality BEFORE the hot path, for all windows of
 * the batch in one call instead of one python iteration per narrative (MFT/models.py:117-132):
 *   CNN      MFT/models.py:57-79   Conv1d(D -> E, kernel k, bias) over the K vectors of a window, then a global max over the
 *                                  K - k + 1 positions                                                     (stage bit 1)
 *   Highway  MFT/models.py:27-55   g = sigmoid(Wg c + bg); g * (Wp c + bp) + (1 - g) * c, then Dropout(0.3) :105,129   (stage bit 2)
 * x fp32: stages & 1 -> [n_win, K, D] raw window vectors; stages == 2 -> [n_win, E] pooled features.  conv_w is nn.Conv1d's
 * [E, D, k]; the Highway weights are nn.Linear's [E, E].  out fp32 [n_win, E].  The conv is ONE GEMM over overlapping rows of x
 * (no im2col copy), tcgen05 in bf16 mode; the workspace carries what backward needs (same cfg, same ws).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int dtype;
  int n_win;         /* B * T windows */
  int K, D, E, k;    /* vectors per window, vector width, embedding width, conv kernel size (K, D, k unused when stages == 2) */
  int stages;        /* 1 = CNN, 2 = Highway + dropout, 3 = both */
  int training;      /* keep what backward needs */
  float dropout_p;   /* dropout probability applied in this call (0.3 in the reference's train mode; pass 0 in eval mode) */
  uint64_t seed;
  uint32_t site;     /* dropout site id (one per modality) */
} MtWindowCnnCfg;
size_t mt_window_cnn_ws_bytes(const MtWindowCnnCfg* cfg);
int mt_window_cnn_fwd(const MtWindowCnnCfg* cfg, const float* x, const float* conv_w, const float* conv_b, const float* wproj,
                      const float* bproj, const float* wgate, const float* bgate, float* out, void* ws, size_t ws_bytes, void* stream);
/* dout fp32 [n_win, E] -> parameter gradients (fp32, OVERWRITTEN); dx fp32 [n_win, E] only when stages == 2 (may be NULL): raw
 * inputs are data and get no gradient. */
int mt_window_cnn_bwd(const MtWindowCnnCfg* cfg, const float* x, const float* dout, float* dx, float* dconv_w, float* dconv_b,
                      float* dwproj, float* dbproj, float* dwgate, float* dbgate, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Batched evaluation metrics: per-narrative concordance correlation coefficient and Pearson r of the first
 * lengths[b] predictions of every narrative of a padded batch, in one launch -- eval_ccc MFT/train.py:42-50 and
 * scipy.stats.pearsonr as called at MFT/train.py:236 (population moments, fp64 accumulation).  pred / target fp32 [B, T];
 * ccc / pearson fp64 [B] (pearson may be NULL); sq_err fp64 [1] (may be NULL): sum over valid steps of (pred - target)^2,
 * OVERWRITTEN (the MSELoss(sum) the evaluation loop accumulates, MFT/train.py:229).
 * ------------------------------------------------------------------------------------------------- */
int mt_ccc_batched(const float* pred, const float* target, const int* lengths, int B, int T, double* ccc, double* pearson,
                   double* sq_err, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * GPU-side batcher: the padded corpus stays on the device; a batch is an index gather instead of the
 * reference's torch.tensor(nested python lists) per batch (generateTrainBatch / generateInputChunkHelper MFT/train.py:59-108).
 * mt_batch_gather: dst[b, 0:prefix] = src[idx[b] * row_stride + 0:prefix] for b < B (fp32; idx device int [B]; prefix <= row_stride:
 * the first T_batch windows of a narrative are a contiguous prefix of its row).  mt_length_mask: mask[b,t] = t < lengths[b] (:103-106).
 * ------------------------------------------------------------------------------------------------- */
int mt_batch_gather(const float* src, size_t row_stride, const int* idx, int B, size_t prefix, float* dst, void* stream);
int mt_length_mask(const int* lengths, int B, int T, float* mask, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Utilities on flat buffers.
 * ------------------------------------------------------------------------------------------------- */
/* out = x + dropout(y) (SublayerConnection.forward MFT/multiTransformer.py:103-104, stand-alone path; x may be NULL)
 * and its gradient wrt y: out = g * dropout_factor.  fp32, element index = flat index. */
int mt_residual_dropout_fwd(const float* x, const float* y, float* out, size_t n, float p, uint64_t seed, uint32_t site,
                            void* stream);
int mt_dropout_bwd(const float* g, float* out, size_t n, float p, uint64_t seed, uint32_t site, void* stream);
/* fp32 -> bf16 shadow copy of a parameter arena. */
int mt_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream);
int mt_cast_bf16_to_f32(const void* src, float* dst, size_t n, void* stream);
/* loss = sum((pred-target)^2) * inv_norm ; dpred = 2*(pred-target)*inv_norm   (MFT/train.py:135-139).
 * loss is ACCUMULATED into *loss (caller zeroes). */
int mt_mse_loss_fwd_bwd(const float* pred, const float* target, size_t n, float inv_norm, float* loss, float* dpred,
                        void* stream);
/* Adam with L2 weight decay folded into the gradient (torch.optim.Adam semantics, MFT/train.py:557) on a flat
 * arena: one launch for all parameters.  step >= 1. */
int mt_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, void* stream);

/* Captured-graph variants: the per-step scalars live in DEVICE memory so one CUDA graph of the whole train step can be
 * replayed with a different loss normaliser / learning rate / step count each time.  step_dev holds the 1-based step
 * count of THIS update; p_lp (optional) receives the bf16 shadow of the updated parameters in the same launch. */
int mt_mse_loss_fwd_bwd_dev(const float* pred, const float* target, size_t n, const float* inv_norm_dev, float* loss, float* dpred,
                            void* stream);
int mt_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, const float* lr_dev, float lr, float beta1, float beta2,
                     float eps, float weight_decay, const int64_t* step_dev, void* p_lp, void* stream);
/* Every dropout site adds *dev_ptr (a device uint64, may be NULL = 0) to its seed: bump it between replays of a captured
 * graph to draw fresh masks.  Process-wide. */
int mt_set_seed_offset_ptr(const uint64_t* dev_ptr);
/* Busy-wait kernel (ms <= 2000): lets the host run ahead of the device so per-launch event timings are back to back. */
int mt_spin(float ms, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (the reference has no distributed code): one process per GPU, narratives sharded over
 * ranks, ONE grouped NCCL all-reduce (SUM, fp32) over the flat gradient arenas per step, enqueued on the caller's stream (capturable).
 * NCCL is resolved at run time from the libnccl.so.2 already loaded in the process; mt_comm_available() == 0 when there is none.
 * Rank 0 calls mt_comm_unique_id, the host broadcasts the 128 bytes, every rank calls mt_comm_init (collective) on its own device.
 * ------------------------------------------------------------------------------------------------- */
int mt_comm_available(void);
int mt_comm_unique_id(char* id128);
int mt_comm_init(const char* id128, int rank, int world, void** comm);
int mt_comm_destroy(void* comm);
int mt_allreduce_grads(void* comm, float* const* bufs, const size_t* counts, int n_bufs, void* stream);
/* Overlapped form ("overlap with the remaining backward").  mt_comm_overlap_arm arms the NEXT mt_encoder_group_bwd call: as
 * soon as that call has enqueued the weight gradients of layers >= split_layer (split_layer < 0: n_layers / 3) and of the final norm of
 * every stack, they are all-reduced (SUM, fp32) on the library's communication stream while the backward of the remaining layers still
 * runs on the caller's stream.  mt_comm_overlap_join makes `stream` wait for that all-reduce and reports the ranges it covered (room for
 * MT_MAX_MODS entries; *n_ranges = 0 when none was started), so the caller passes only the rest to mt_allreduce_grads.  Both streams may
 * be under CUDA-graph capture (fork / join through events). */
int mt_comm_overlap_arm(void* comm, int split_layer);
int mt_comm_overlap_join(void* stream, float** ptrs, size_t* counts, int* n_ranges);

/* test hook: keep the bf16 LSTM-decoder forward (mt_lstm_head_fwd) on the FFMA kernel instead of the cluster / tensor-core kernel
 * (csrc/mt_lstm_head_mma.cu: E == 256, bf16 mode); returns the previous setting. */
int mt_lstm_head_force_ffma(int on);

/* Generic GEMM (exposed for tests and profiling):  C[M,N] = A·B^T-style contraction, see csrc/mt_gemm.cuh.
 * a_kmajor: A element (m,k) at A[m*lda+k] (else A[k*lda+m]); b_kmajor: B element (n,k) at B[n*ldb+k] (else B[k*ldb+n]). */
int mt_gemm(int dtype, int M, int N, int K, const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor,
            void* C, int ldc, int c_f32, const float* bias, int act, int split_k_atomic, void* stream);
/* Row-stream GEMM engine of the encoder projections (csrc/mt_gemm_rs.cu; exposed for tests and profiling): weight-resident, grouped
 * over G modality stacks (nn.Linear call sites MFT/multiTransformer.py:19-20,47-65 and their input gradients).
 *   C[g*Mg + m, n] = epi( sum_k A[g*Mg + m, k] B_g(n, k) ),  A bf16 [G*Mg, K];  B = G weight matrices back to back, each [N,K] (b_kmajor)
 * or [K,N];  C bf16 or fp32 [G*Mg, N];  bias / colsum / ln_a / ln_b: G x N fp32 back to back or NULL;  drop_p > 0: output dropout with
 * site `site + 512 g` on the group-local element index;  gate bf16 [G*Mg, N]: out = gate > 0 ? out * gate_scale : 0;  residual fp32
 * [G*Mg, N];  colsum is ACCUMULATED;  ln_out (bf16 [G*Mg, N], needs N == 256 and c_f32): LayerNorm(C) with unbiased std, eps 1e-6 on std.
 * MT_ERR_UNSUPPORTED outside the encoder's (N, K, feature) combinations. */
int mt_gemm_rs(int G, int Mg, int N, int K, const void* A, const void* B, int b_kmajor, void* C, int c_f32, const float* bias, int act,
               float drop_p, uint64_t seed, uint32_t site, const void* gate, float gate_scale, const float* residual, float* colsum,
               void* ln_out, const float* ln_a, const float* ln_b, void* stream);
/* debug hook of the row-stream engine: CTA 0 writes clock64 stamps (16 x uint64 per tile, first 32 tiles: TMA issue of k-blocks 0-3, their
 * arrival as seen by the MMA thread, accumulator free, epilogue start / end of the two warp groups) into dev_buf (>= 4 KB); NULL = off. */
int mt_gemm_rs_trace(void* dev_buf);
/* which engine a (dtype, shape) GEMM would use: 0 = FFMA SIMT, 1 = tcgen05 */
int mt_gemm_engine(int dtype, int M, int N, int K, int a_kmajor, int b_kmajor);
/* test hook: route every GEMM through the FFMA engine (A/B the tensor-core engine); returns the previous setting. */
int mt_gemm_force_simt(int on);
/* tuning hook: 0 = 256-wide weight-resident tiles where they apply (two CTAs per SM with a streaming ring elsewhere), 1 = one CTA per
 * SM everywhere, 2 = never use the 256-wide tile (default, measured fastest on the MFT step); returns the previous mode. */
int mt_gemm_tc_mode(int mode);
/* tuning knobs; returns the previous value (-1: bad key).  key 0 / 1 / 2 = grid share of the tcgen05 GEMM / the T <= 128 attention /
 * the LayerNorm kernels: a share s > 1 launches 1/s of the resident CTA slots, so kernels of concurrent streams (the modality stacks
 * of MultiTransformer) co-reside on the SMs instead of queueing behind each other.  key 3 = programmatic dependent launch of the encoder
 * chain (LayerNorm, tcgen05 GEMMs, tcgen05 attention, keep-bit draw: grid launch and prologue overlap the previous kernel's tail): 0 off
 * (default), 1 = launches outside stream capture only (eager MFT step 7.27 -> 7.05 ms), 2 = always (a captured graph gets 1 % slower);
 * EXPERIMENTAL -- with a row-stream GEMM and the attention forward both launched this way the bf16 forward was measured run-to-run
 * non-deterministic (csrc/mt_common.cuh, tools/fwd_determinism.py); key 14 = bit mask of the kernel families key 3 applies to; key 15 bit 0: the attention forward does not
 * trigger its dependents early (with it the chain was measured deterministic again); key 15 bit 1: the attention backward keeps its light
 * preparation launch (A/B; by default the output projection's input-gradient GEMM writes all four per-query scalars from its epilogue).
 * key 5 != 0: the encoder's projections skip the weight-resident row-stream engine (A/B against the streaming engine); key 6 != 0: no
 * LayerNorm fused into the FFN output projection's epilogue; key 7 != 0: no 256-row / 256-wide tiles for the L2-bound GEMMs; key 8: debug mask of the
 * MFN recurrence kernels (bit 6: the first-cut kernels, for A/B timing; bit 5: clock trace); key 9 != 0: the tcgen05 attention kernels hash
 * their dropout draws themselves instead of reading the keep bits drawn once per step and layer; key 10 != 0: streaming GEMMs over the stacked
 * modality rows launch per stack; key 11 != 0: the output projection fuses the sublayer-1 LayerNorm across a CTA pair (opt-in); key 12 != 0: the keep bits are
 * drawn by the stand-alone kernel instead of inside the LayerNorm forward pass that precedes the attention; key 13 != 0: bf16 mode carries the residual-stream gradient
 * between the sublayers of a stack in bf16 instead of fp32 (opt-in: fewer bytes, measured slower). */
int mt_tune(int key, int value);
/* debug hook: CTA 0 of every tcgen05 GEMM writes per-tile clock64 stamps (8 x uint64 per tile, first 64 tiles: TMA issue, MMA
 * tile start, first operands landed, last k-block landed, epilogue sees the accumulator, accumulator released, last pass
 * starts; then per-k-block issue / landing stamps of the first 16 tiles) into dev_buf (>= 8 KB); NULL switches it off. */
int mt_gemm_debug_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* MT_B200_H */
