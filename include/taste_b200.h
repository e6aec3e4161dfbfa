/* taste_b200.h — C ABI of the B200-native TASTE speech-tokenization path (libtaste_b200.so).
 *
 * The reference (dienruei123/TASTE-SpokenLM) has NO native/FFI layer on this path: everything below replaces
 * PyTorch module calls.  Each entry point cites the reference interface it stands in for, using the
 * abbreviations of SURVEY.md:
 *   WF  taste_speech/modules_taste/cosyvoice/whisper_frontend.py
 *   JES taste_speech/modules_taste/audio_joint_encoder_segmenter.py
 *   CW  taste_speech/modules_taste/cosyvoice/customized_whisper.py
 *   MT  taste_speech/modeling_taste.py
 *   AQ  taste_speech/modules_taste/audio_quantizer.py
 *   RVQ taste_speech/modules_taste/vq/residual_vq.py     VQ  .../vq/vector_quantize_pytorch.py
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless a name ends in `_host`;
 *   - the library never allocates or frees device memory: the caller passes a workspace sized by
 *     taste_ws_bytes(); every launch is enqueued on the caller's stream (`stream` is a cudaStream_t passed as
 *     void*); no host synchronisation inside;
 *   - return value: 0 = success; negative = argument error (TASTE_E_*); positive = cudaError_t / CUresult;
 *     taste_last_error() returns a thread-local message for the last non-zero return;
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every compute entry point fails.
 */
#ifndef TASTE_B200_H_
#define TASTE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TASTE_ABI_VERSION 4

/* Library flavour: the 16-bit type of every tensor-core operand buffer below that is documented as "bf16" (packed
 * weights, log-mel features for the stem, encoder states h_last / h_target, GEMM / attention operands).
 * 0 = bf16 (libtaste_b200.so; BASELINE config 2), 1 = fp16 (libtaste_b200_f16.so, built with -DTASTE_F16=1: the
 * reference's own GPU dtype under torch.cuda.amp.autocast(), JES:133 / JES:336).  fp32 buffers are fp32 in both. */
int taste_operand_dtype(void);

#define TASTE_E_ARG        (-1)  /* null pointer / bad size */
#define TASTE_E_SHAPE      (-2)  /* geometry not supported by the kernels (see taste_handle_create) */
#define TASTE_E_WORKSPACE  (-3)  /* workspace too small */
#define TASTE_E_NO_DEVICE  (-4)  /* no sm_100 device / driver entry point missing */

#define TASTE_N_FFT      400
#define TASTE_HOP        160
#define TASTE_N_SAMPLES  480000   /* WF:30 pad_samples */
#define TASTE_N_FRAMES   3000     /* JES:97 expected_seq_length */
#define TASTE_N_MELS     128
#define TASTE_ENC_FRAMES 1500
#define TASTE_DFT_LD     224      /* padded leading dim of the DFT tables */
#define TASTE_DFT_K      448      /* tensor-core DFT: samples per frame slab (400, zero-weighted up to 7 x 64) */
#define TASTE_DFT_N      512      /* ... output columns: Re X[0..200] at 0, Im X[0..200] at 256, rest zero */

typedef struct taste_handle_s* taste_handle_t;

/* Geometry (distil-large-v3 + CFG:146-155 by default). */
typedef struct {
  int32_t d_model;          /* 1280; multiple of 128 */
  int32_t heads;            /* 20; head_dim must be 64 */
  int32_t ffn;              /* 5120; multiple of 128 */
  int32_t enc_layers;       /* 32 */
  int32_t dec_layers;       /* 2 */
  int32_t vocab;            /* 51866 */
  int32_t max_target_pos;   /* 448 */
  int32_t codebook_dim;     /* 256 (fixed by the RVQ kernel) */
  int32_t codebook_size;    /* 512 (fixed by the RVQ kernel) */
  int32_t num_quantizers;   /* 4; <= 8 */
  int32_t target_layer;     /* 6: hidden state ENTERING this encoder layer is the value source (JES:192-193) */
  int32_t reserved;
} taste_dims_t;

/* One Whisper encoder layer (CW:649-716).  Matrices are bf16 row-major [out, in]; vectors fp32. */
typedef struct {
  const float* ln1_w;  const float* ln1_b;     /* self_attn_layer_norm */
  const void*  wqkv;   const float* bqkv;      /* [3D, D]: q rows pre-scaled by head_dim^-0.5 (CW:342), k bias = 0 (CW:315) */
  const void*  wo;     const float* bo;        /* out_proj */
  const float* ln2_w;  const float* ln2_b;     /* final_layer_norm */
  const void*  w1;     const float* b1;        /* fc1 [FF, D] */
  const void*  w2;     const float* b2;        /* fc2 [D, FF] */
  /* Optional LayerNorm-folded copies (all six or none).  With them the encoder never materialises LayerNorm(h): the
   * QKV / fc1 GEMMs read the raw bf16 rows and apply  rstd * (acc - mean * c[n]) + b'[n]  in their epilogue, with
   *   W'[n][k] = W[n][k] * ln_w[k]  (bf16),  c[n] = sum_k W'[n][k],  b'[n] = b[n] + sum_k W[n][k] * ln_b[k].      */
  const void*  wqkv_ln;  const float* bqkv_ln;  const float* cqkv_ln;     /* self_attn_layer_norm folded into QKV */
  const void*  w1_ln;    const float* b1_ln;    const float* c1_ln;       /* final_layer_norm folded into fc1 */
} taste_enc_layer_t;

/* One aggregator (Whisper decoder) layer (CW:719-833). */
typedef struct {
  const float* ln1_w;  const float* ln1_b;     /* self_attn_layer_norm */
  const void*  wqkv;   const float* bqkv;      /* causal self-attention, packed as above */
  const void*  wo;     const float* bo;
  const float* lnx_w;  const float* lnx_b;     /* encoder_attn_layer_norm */
  const void*  wq_x;   const float* bq_x;      /* encoder_attn.q_proj (pre-scaled) */
  const void*  wk_x;                           /* encoder_attn.k_proj, no bias; applied to the LAST encoder state */
  const void*  wv_x;   const float* bv_x;      /* encoder_attn.v_proj; applied to the layer-`target_layer` input */
  const void*  wo_x;   const float* bo_x;
  const float* ln2_w;  const float* ln2_b;     /* final_layer_norm */
  const void*  w1;     const float* b1;
  const void*  w2;     const float* b2;
} taste_dec_layer_t;

typedef struct {
  taste_dims_t dims;
  /* log-mel tables (WF:56-85; A1): folded DFT twiddles and the sparse Slaney filterbank */
  const float*   dft_cos;       /* [200][TASTE_DFT_LD]: cos(2*pi*k*n/400), n = row+1 (1..199; row 199 unused), k = col */
  const float*   dft_sin;       /* [200][TASTE_DFT_LD]: sin(2*pi*k*n/400) */
  const float*   hann;          /* [400] periodic Hann */
  const int32_t* mel_start;     /* [128] first non-zero bin of each mel filter */
  const int32_t* mel_count;     /* [128] number of non-zero bins */
  const float*   mel_weight;    /* [128][TASTE_MEL_MAXW] non-zero weights, zero padded */
  /* tensor-core DFT (nullable: then the fp32 FMA kernel runs): bf16 [TASTE_DFT_N][3 * TASTE_DFT_K], row j = output
   * column j, K slabs {hi(t), lo(t), hi(t)} of t[j][n] = hann[n] * cos|sin(2*pi*bin*n/400) (split bf16, see logmel.cu) */
  const void*    dft_w_bf16;
  /* encoder stem (JES:174-181) */
  const void*  conv1_w;  const float* conv1_b;   /* bf16 [D, 3*128], k = tap*128 + c */
  const void*  conv2_w;  const float* conv2_b;   /* bf16 [D, 3*D],   k = tap*D + c */
  const float* enc_pos;                          /* fp32 [1500, D] embed_positions.weight */
  const taste_enc_layer_t* enc;                  /* HOST array [enc_layers] */
  const float* enc_ln_w;  const float* enc_ln_b; /* encoder.layer_norm */
  /* aggregator (CW:1160-1437) */
  const float* tok_emb;                          /* fp32 [vocab, D] */
  const float* dec_pos;                          /* fp32 [max_target_pos, D] */
  const taste_dec_layer_t* dec;                  /* HOST array [dec_layers] */
  const float* dec_ln_w;  const float* dec_ln_b;
  /* RVQ (RVQ:102-170, VQ:266-340); all fp32 */
  const float* rvq_win_t;     /* [D][256]      project_in.weight transposed */
  const float* rvq_bin;       /* [256] */
  const void* rvq_code_split; /* bf16 [Q][2][512][256]: hi = bf16(e), lo = bf16(e - hi) planes of the codebooks (tensor-core distance pass) */
  const float* rvq_code;      /* [Q][512][256] codebooks (gather) */
  const float* rvq_code_sq;   /* [Q][512]      |e|^2 */
  const float* rvq_wout_t;    /* [256][D]      project_out.weight transposed */
  const float* rvq_bout;      /* [D] */
} taste_weights_t;

#define TASTE_MEL_MAXW 16

/* --- library ---------------------------------------------------------------------------------------------- */
int         taste_abi_version(void);
const char* taste_last_error(void);

/* Build an immutable handle from a weights descriptor (the struct and its host arrays are copied; the device
 * buffers they point to stay owned by the caller and must outlive the handle).
 * Stands in for TasteAudioTower.__init__ + checkpoint load (MT:34-95, JES:281-328). */
int taste_handle_create(const taste_weights_t* w, taste_handle_t* out);
int taste_handle_destroy(taste_handle_t h);

/* Workspace bytes needed by the calls below for a batch of `batch` utterances whose assembled transcripts
 * (prefix + tokens + 1) total `sum_tokens` rows. */
size_t taste_ws_bytes(taste_handle_t h, int batch, int sum_tokens);

/* --- R1: WhisperFrontend.forward (WF:87-113) -------------------------------------------------------------- */
/* wav: fp32 [batch, wav_stride]; n_samples: int32 [batch] (samples valid in each row; <= 480000 are used, the
 * rest of the 30 s window is zero, WF:98-99).  feats_f32: [batch,3000,128] (nullable); feats_bf16 (nullable):
 * same layout in bf16 for the encoder stem.  At least one output must be given. */
int taste_logmel_f32(taste_handle_t h, const float* wav, const int32_t* n_samples, int batch, int64_t wav_stride,
                     float* feats_f32, void* feats_bf16, void* ws, size_t ws_bytes, void* stream);

/* --- R2/R3: WhisperAudioEncoderForJoint.forward (JES:133-223) --------------------------------------------- */
/* feats_f32 [batch,3000,128] (or feats_bf16 if feats_f32 is NULL) -> h_last, h_target: bf16 [batch,1500,D]
 * (final-LayerNorm state and the hidden state entering layer `target_layer`). */
int taste_encoder_fwd(taste_handle_t h, const float* feats_f32, const void* feats_bf16, int batch, void* h_last_bf16,
                      void* h_target_bf16, void* ws, size_t ws_bytes, void* stream);

/* --- R4/R5: token assembly (MT:144-152) + WhisperDecoder.forward with dict K/V (JES:377-388, CW:1200-1437) - */
/* Token assembly (MT:144-152) on the device: asr_token_ids int64 [batch,tmax] (padded), token_lengths int32 [batch],
 * cu_tokens int32 [batch+1] with cu[b+1]-cu[b] = T_b + 5  ->  tokens int32 packed [cu[batch]]:
 * <sot>,<en>,<transcribe>,<notimestamps>, ids[b,:T_b], then the entry that follows in the reference's padded row
 * (ids[b,T_b], or <eot> when T_b == tmax). */
int taste_assemble_tokens(const int64_t* asr_token_ids, const int32_t* token_lengths, const int32_t* cu_tokens, int batch,
                          int tmax, int32_t* tokens, void* stream);
/* tokens: int32 packed [sum_tokens] assembled ids; cu_tokens: int32 [batch+1] row offsets.
 * dec_out: fp32 packed [sum_tokens, D] = decoder final LayerNorm state at every assembled position. */
int taste_aggregator_fwd(taste_handle_t h, const void* h_last_bf16, const void* h_target_bf16, const int32_t* tokens,
                         const int32_t* cu_tokens, int batch, int sum_tokens, int max_tokens, float* dec_out, void* ws,
                         size_t ws_bytes, void* stream);

/* --- R6: prefix skip + word pooling + EOS drop (JES:393-458, MT:170-172) ----------------------------------- */
/* dec_out packed as above (row 4+t of utterance b is transcript position t); word_ids int32 [batch,tmax] (padded
 * with the caller's pad value, normally 0 — runs are formed on the padded row exactly as JES:437-458);
 * token_lengths int32 [batch].  z: fp32 [batch,tmax,D]; rows t >= token_lengths[b] are written as 0. */
int taste_word_pool_f32(const float* dec_out, const int32_t* cu_tokens, const int32_t* word_ids,
                        const int32_t* token_lengths, int batch, int tmax, int d_model, float* z, void* stream);

/* --- R7: RVQAudioQuantizer.forward -> ResidualVQ.forward (AQ:109-124, RVQ:359-490) ------------------------- */
/* z fp32 [batch,tmax,in_dim]; lengths int32 [batch] (mask = t < lengths[b], modules_taste/utils.py:5-8; NULL =
 * all valid).  in_dim == d_model: project_in is applied (RVQ:371); in_dim == 256: `z` is already a code
 * (ResidualVQ.get_indices_from_code, RVQ:258-357).  indices int64 [batch,tmax,Q] (-1 at masked rows);
 * quantized fp32 [batch,tmax,D] (nullable) = project_out(sum of codes) (RVQ:470).
 * ws: taste_rvq_ws_bytes(batch * tmax) bytes of device scratch (projected input and code sums; residuals themselves
 * never leave the SM between levels). */
size_t taste_rvq_ws_bytes(int n_rows);
int taste_rvq_encode_f32(taste_handle_t h, const float* z, const int32_t* lengths, int batch, int tmax, int in_dim,
                         int64_t* indices, float* quantized, void* ws, size_t ws_bytes, void* stream);
/* ResidualVQ.get_output_from_indices / get_code_from_indices (RVQ:183-242): indices int64 [n,Q] (-1 -> zero code).
 * out fp32 [n, D] if project_out != 0 else [n, 256]. */
int taste_rvq_decode_f32(taste_handle_t h, const int64_t* indices, int n, int project_out, float* out, void* stream);

/* --- (f)1: TasteForCausalLM.extract_vq epilogue (MT:1438-1450, MT:1877-1881) ------------------------------- */
/* asr_indices int64 [batch,tmax,Q] -> llm_indices int64 [batch,lmax,Q] (-1 where no word-start match). */
int taste_map_to_llm_tokens(const int64_t* asr_indices, const int32_t* asr_word_ids, const int32_t* asr_lengths,
                            const int32_t* llm_word_ids, const int32_t* llm_lengths, int batch, int tmax, int lmax,
                            int num_q, int64_t* llm_indices, void* stream);

/* --- (f)2: corpus ingest, `resampler(speech_pt).mean(0)` of process_one_sample (DS:52-60) --------------------- */
/* torchaudio.transforms.Resample(orig_sr, 16000) (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99; third
 * party, restated in oracle/taste_oracle.py) followed by the channel mean, for a batch of decoded PCM arrays.
 * in: fp32, utterance b is [channels[b], n_in[b]] row-major at in + in_offsets[b] (in_offsets int64 [batch+1]).
 * orig_reduced / new_reduced: the two rates divided by their gcd; width, taps: the polyphase kernel of
 * functional._get_sinc_resample_kernel in sparse form - phase p uses taps[p*taps_ld .. +taps_per_phase) against
 * xpad[f*orig + tap_start[p] + k] (xpad = signal left-padded with `width` zeros).  Identity (orig == new): 1, 1, 0,
 * taps {1}, tap_start {0}.  Output row b of wav [batch, wav_stride]: the first min(ceil(new*n/orig), wav_stride)
 * samples (the rest of the row is not written: taste_logmel_f32 treats it as zero given n_samples); n_out
 * (nullable) int32 [batch] = ceil(new*n/orig).  max_out bounds the grid (max over b of the output length);
 * total_in_elems / total_out_elems are used for accounting only. */
int taste_resample_mean_f32(const float* in, const int64_t* in_offsets, const int32_t* channels, const int32_t* n_in,
                            int batch, int orig_reduced, int new_reduced, int width, const float* taps,
                            const int32_t* tap_start, int taps_per_phase, int taps_ld, int max_out,
                            int64_t total_in_elems, int64_t total_out_elems, float* wav, int64_t wav_stride,
                            int32_t* n_out, void* stream);

/* --- building blocks exported for kernel-level parity tests and micro-benchmarks ---------------------------- */
/* C[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias).  A, W bf16 row-major; bias fp32 [N] (nullable).
 * epilogue: 0 = bf16 out; 1 = GELU(erf) then bf16 out; 2 = fp32 out += (residual add in place, CW:692, 702);
 * 3 = fp32 out.  Requires N % 128 == 0, K % 64 == 0. */
int taste_gemm_bf16(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int epilogue,
                    void* stream);
/* General form of taste_gemm_bf16 with the optional LayerNorm-folding operands (see taste_enc_layer_t):
 *   producer: stats_out [m][n/128][2] fp32 and out_bf16 [m,n] (epilogue 2 only): also emits the bf16 copy of the fp32
 *             output rows and their per-128-column (sum, sum of squares);
 *   consumer: ln_stats [m][ln_nseg][2], ln_colsum [n], bias = b' (epilogue 0 or 1): applies the folded LayerNorm.
 * Requires n % 256 == 0 (CTA-pair kernel).  Unused pointers are NULL. */
typedef struct {
  const void* a; const void* w; const float* bias; void* out;
  int32_t m, n, k, epilogue;
  const float* ln_stats; int32_t ln_nseg; int32_t reserved; const float* ln_colsum;
  float* stats_out; void* out_bf16;
} taste_gemm_ex_t;
int taste_gemm_ex(const taste_gemm_ex_t* g, void* stream);
/* Diagnostic switches (taste_*_set_mode): process-wide, read at launch time, NOT synchronised - set them once before
 * any concurrent use; they exist for A/B timing, for tests that pit two formulations against each other, and as the
 * documented escape hatch below.  Everything else in this ABI is safe to call concurrently on different streams.
 *
 * Encoder: 0 = fold self_attn_layer_norm into the fc2 -> QKV GEMM pair when the folded weights are present and the batch
 * has >= 2048 rows (final_layer_norm stays a kernel), 1 = always run the separate two-pass LayerNorm kernels, 2 = fold both
 * LayerNorms (measured slower; A/B only).  The fold rounds the RAW residual rows to the 16-bit operand type and removes
 * the row mean after the GEMM, so its rounding error scales with |row mean| / row std (tests: Whisper-like rows with
 * +-300 outlier channels and a half-sigma common offset stay inside the LayerNorm -> GEMM budget); a checkpoint whose
 * residual rows carry a common-mode offset far above their spread should run with mode 1. */
int taste_encoder_set_mode(int mode);
/* Front-end formulation for A/B timing and tests: 0 = split-bf16 DFT on the tcgen05 GEMM kernel when dft_w_bf16 is
 * present (default), 1 = the fp32 folded-DFT FMA kernel.  Process-wide. */
int taste_logmel_set_mode(int mode);
/* Tile-shape override for A/B timing and tests: 0 = automatic (CTA pairs, 256 x 256 tiles, when one wave of pair tiles
 * exists), 1 = always the single-CTA 128 x {256,128} kernel.  Process-wide. */
int taste_gemm_set_mode(int mode);
/* y = LayerNorm(x) (eps 1e-5, CW:660): x fp32 [rows, d]; y bf16 (out_bf16 != 0) or fp32. */
int taste_layernorm_f32(const float* x, const float* w, const float* b, void* y, int rows, int d, int out_bf16,
                        void* stream);
/* The aggregator's cross-attention (CW:361-366 with the dict K/V of JES:377-388): ragged queries against fixed-length keys,
 * no mask.  q / o are PACKED [total_q, ld] matrices, utterance b owning rows cu_q[b] .. cu_q[b+1] (at most max_q_len of
 * them); k / v are [batch * kv_len, ld].  Runs on the tcgen05 / TMA attention kernel (one 128-query tile per work item). */
int taste_attention_ragged_bf16(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo,
                                const int32_t* cu_q, int total_q, int max_q_len, int kv_len, int batch, int heads,
                                void* stream);
/* softmax(Q K^T [+ causal mask]) V per head (CW:377-394); q is expected pre-scaled.  bf16, head_dim 64.
 * Row b of q/o starts at cu_q[b] (or b*q_len if cu_q is NULL) and has cu_q[b+1]-cu_q[b] (or q_len) rows; same
 * for k/v with cu_kv / kv_len.  ld* are row strides in elements. */
int taste_attention_bf16(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo,
                         const int32_t* cu_q, const int32_t* cu_kv, int q_len, int kv_len, int batch, int heads,
                         int causal, void* stream);

/* --- instrumentation (no reference counterpart: SURVEY.md section 5 "tracing / profiling: none") ------------- */
/* Kernels launched by this library in the calling process since load. */
unsigned long long taste_launch_count(void);
/* When enabled, every launch is bracketed by CUDA events on its own stream; taste_prof_collect() synchronises those
 * events and returns one entry per kernel class that launched since the last taste_prof_reset(). */
typedef struct {
  const char* name;
  long long   launches;
  double      total_ms;   /* sum of per-launch device durations */
  double      flops;      /* algorithmic FLOPs of those launches */
  double      bytes;      /* algorithmic HBM bytes of those launches */
} taste_prof_entry_t;
int taste_prof_enable(int on);
int taste_prof_reset(void);
int taste_prof_collect(taste_prof_entry_t* out, int max_entries, int* n_out);

/* Kernel override for A/B timing and tests: 0 = automatic (tcgen05/TMEM kernel for fixed-length non-causal problems
 * with >= 256 queries, mma.sync kernel for the ragged / causal aggregator shapes), 1 = always the mma.sync kernel. */
int taste_attention_set_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* TASTE_B200_H_ */
