/*
 * vaura_b200 — C ABI of the B200-native V-AURA generation hot path.
 *
 * The reference (ilpoviertola/V-AURA) is pure Python/PyTorch and has no FFI of its own; its plugin
 * boundary is `instantiate_from_config({"target","params"})` (utils/utils.py:11-22) plus the
 * nn.Module call signatures.  This header is the C boundary that sits *under* the Python mirror of
 * that interface (vaura_b200/model.py, sampler.py, codec.py); every entry point names the reference
 * code it replaces.  Conventions:
 *   - plain C types only; every tensor is a raw DEVICE pointer owned by the caller (weights, KV
 *     pages, workspaces, sequences, outputs).  The library allocates no device memory.
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*); nothing executes
 *     on any other stream and there is no host synchronisation.  vaura_sampler_generate captures the
 *     decode step into a CUDA graph (on a private capture-origin stream, so the legacy default stream
 *     is usable as `stream`) and replays it asynchronously on `stream`.
 *   - return value: 0 = ok, otherwise one of VAURA_ERR_*; the message of the last error on the
 *     calling thread is returned by vaura_last_error().  Nothing throws across the ABI.
 *   - one host thread per handle; handles are independent (one process per GPU creates its own).
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails with
 *     VAURA_ERR_CUDA.
 */
#ifndef VAURA_B200_H
#define VAURA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAURA_OK 0
#define VAURA_ERR_INVALID 1      /* bad argument / shape */
#define VAURA_ERR_UNSUPPORTED 2  /* shape the kernels are not built for (e.g. head_dim != 96) */
#define VAURA_ERR_CUDA 3         /* CUDA runtime error; message has the cudaError string */
#define VAURA_ERR_WORKSPACE 4    /* caller-provided workspace too small */

#define VAURA_PRECISION_AUTO 0    /* BF16 from 16 sequence rows, and from 3 rows when the call samples (top-k / top-p /
                                     temperature: no bit-exactness contract); FP32ACT otherwise (greedy, rows <= 2) */
#define VAURA_PRECISION_FP32ACT 1 /* bf16 weights, fp32 activations + fp32 KV (HBM-bound small batch).  Decode steps: fp32-exact
                                     products (cluster / persistent kernels).  Multi-position passes (prompt prefill, teacher-forced
                                     forward) run on tcgen05 with every fp32 operand split into three bf16 terms (fp32-equivalent,
                                     greedy tokens stay bit-exact); the prompt prefill of a call that samples rounds the GEMM
                                     operands to bf16 once instead (residual stream, q, K/V and attention still fp32; logits
                                     within the BF16 tolerance; csrc/knobs.h: VAURA_PREFILL_BF16=0 turns that off) */
#define VAURA_PRECISION_BF16 2    /* bf16 weights + bf16 activations/KV, tcgen05 GEMMs, fp32 accumulate */

#define VAURA_KV_F32 0
#define VAURA_KV_BF16 1

typedef struct vaura_sampler vaura_sampler; /* opaque: AR transformer + sampling */
typedef struct vaura_codec vaura_codec;     /* opaque: DAC token->waveform decoder */

/* ---- library --------------------------------------------------------------------------------- */
int vaura_version(void);                 /* ABI version, currently 1 */
const char* vaura_arch(void);            /* "sm_100a" */
const char* vaura_last_error(void);      /* thread-local, never NULL */

/* Number of kernels this library has launched in the calling process (graph replays count their
 * kernel nodes).  bench.py reports the difference over its timed region as "gpu_launches". */
unsigned long long vaura_launch_count(void);

/* Weight-streaming GEMV, the dominant kernel of the fp32-activation decode path, as a stand-alone op:
 * y[r][n] = sum_k W[n][k] * x[r][k]; W bf16 [N][K] row-major, x f32 [R][K], y f32 [R][N]; fp32 accumulate.
 * (What nn.Linear does at llama.py:228/:259/:176-177/:504 for one position.)  N even, K % 8 == 0. */
int vaura_gemv_bf16w(const uint16_t* W, const float* x, float* y, int32_t N, int32_t K, int32_t R, void* stream);

/* tcgen05 linear layer, the dominant kernel of the bf16 (rows >= 16 / prefill) path, as a stand-alone op:
 * y[r][n] = sum_k A[r][k] * W[n][k]; A bf16 [R][K], W bf16 [N][K], y f32 [R][N]; fp32 accumulate in TMEM.
 * K % 64 == 0, N % block_n == 0, block_n in {16, 32, 64, 128}. */
int vaura_linear_bf16(const uint16_t* A, const uint16_t* W, float* y, int32_t R, int32_t N, int32_t K, int32_t block_n,
                      void* stream);

/* ---- AR transformer ("sampler"), replaces models/modules/sampler/llama.py:286-586 -------------- */
typedef struct {
  int32_t num_layers;    /* 24 */
  int32_t d_model;       /* 1536 */
  int32_t nhead;         /* 16  (head_dim = d_model / nhead must be 96) */
  int32_t ffn_dim;       /* 4096 (llama.py:164-169) */
  int32_t vocab;         /* 1024; special/BOS id == vocab */
  int32_t num_codebooks; /* 9 */
  int32_t block_size;    /* 256 RoPE rows (llama.py:317, :364-368) */
  int32_t cond_dim;      /* 512 = d_model / cond_feature_channel_scaler */
  int32_t cond_in;       /* 768 AVCLIP width */
  int32_t cond_tokens;   /* 32 */
  int32_t audio_tokens_per_video_frame; /* 7 (scripts/generate.py:216) */
  float norm_eps;        /* 1e-5 */
} vaura_sampler_dims;

/* All device pointers; bf16 stored as uint16_t.  Packing is done by the host mirror
 * (vaura_b200/weights.py); layouts:                                                              */
typedef struct {
  const uint16_t* wqkv;     /* [L][3*d][d]   layers.{i}.attention.wqkv.weight                      */
  const uint16_t* wo;       /* [L][d][d]     layers.{i}.attention.wo.weight                        */
  const uint16_t* w13;      /* [L][2*F][d]   rows interleaved: 2j = w1 row j, 2j+1 = w3 row j      */
  const uint16_t* w2;       /* [L][d][F]     layers.{i}.feed_forward.w2.weight                     */
  const uint16_t* w_heads;  /* [K*V][d]      lm_heads.{k}.weight stacked (llama.py:503-504)        */
  const float* attn_norm;   /* [L][d] */
  const float* ffn_norm;    /* [L][d] */
  const float* final_norm;  /* [d] */
  const float* tok_tables;  /* [K][V+1][d-cond_dim]  folded emb_k @ W_k^T + b_k (llama.py:60-73)   */
  const float* rope;        /* [block_size][head_dim/2][2] cos,sin (llama.py:593-603)              */
  const float* fc1;         /* [cond_dim][cond_in]  cls_embeddings.projection.fc1.weight           */
  const float* fc2;         /* [cond_dim][cond_dim] cls_embeddings.projection.fc2.weight           */
  const float* empty_video_emb; /* [cond_dim] (llama.py:336-338) */
  /* Optional second copy for the cluster-persistent decode kernel (rows <= 2; csrc/decode_cluster.cu), as 12288-byte
   * slots of 16x16 bf16 tiles in mma.m16n8k16 A-fragment register order, in consumption order:
   *   [16 heads][4 ranks][L][18 q|k|v + 6 wo slots]   shared by the two clusters that serve a head
   *   [128 CTAs][L x (16 w13 + 8 w2) + 18 heads slots]
   * (vaura_b200/weights.py: pack_cluster_stream documents the exact index map).  NULL disables that kernel. */
  const uint8_t* wstream;
} vaura_sampler_weights;

/* Paged KV cache, caller-owned.  Element (layer l, kv in {0=K,1=V}, sequence b, position p, head h, dim e):
 *   page = page_table[b*max_pages_per_seq + p/page_size]
 *   index = ((((l*2 + kv)*num_pages + page)*nhead + h)*page_size + p%page_size)*head_dim + e        */
typedef struct {
  void* pages;
  const int32_t* page_table; /* device, [rows][max_pages_per_seq]; rows = batch * (cfg ? 2 : 1) */
  int32_t num_pages;
  int32_t page_size;         /* 16 or 32 */
  int32_t max_pages_per_seq;
  int32_t dtype;             /* VAURA_KV_F32 | VAURA_KV_BF16 (must match the precision mode) */
} vaura_kv_cache;

int vaura_sampler_create(const vaura_sampler_dims* dims, const vaura_sampler_weights* weights,
                         vaura_sampler** out);
void vaura_sampler_destroy(vaura_sampler* s);

/* Conditioning projection, once per clip (replaces AVCLIPEmbedder/MLP llama.py:79-141 and the
 * per-step _repeat_and_pad_video llama.py:555-586): rows_out[r][0..Tv) = fc2(gelu_tanh(fc1(feats[r])));
 * rows_out[r][Tv] = empty_video_emb.  feats [rows][Tv][cond_in] f32, rows_out [rows][Tv+1][cond_dim] f32. */
int vaura_sampler_cond_project(vaura_sampler* s, const float* feats, int32_t rows, int32_t tv,
                               float* rows_out, void* stream);

size_t vaura_sampler_workspace_bytes(const vaura_sampler* s, int32_t rows, int32_t max_positions,
                                     int32_t precision);

/* The decode loop of VAURAModel.generate (models/vaura_model.py:502-547) with _sample_next_token
 * (:775-827) fused on the device: for offset in [start_offset, S): run the transformer on the columns
 * not yet in the KV cache (a prefill when more than one), CFG-combine, sample / argmax, apply the
 * delay-pattern mask-fix (:536-537) and the prompt-preserving write-back (:540-544).             */
typedef struct {
  int32_t batch;         /* B clips */
  int32_t use_cfg;       /* 1: rows = 2B, [cond; uncond] (vaura_model.py:786-795) */
  int32_t timesteps;     /* T = max_new_tokens; S = T + num_codebooks columns */
  int32_t start_offset;  /* first column to sample = prompt_len + 1 */
  int32_t end_offset;    /* one past the last column to sample; S for a full run */
  int32_t use_sampling;  /* 0: argmax (also when temp <= 0, vaura_model.py:816,:824-825) */
  float temp;
  int32_t top_k;         /* used when top_p <= 0 and top_k > 0 */
  float top_p;           /* > 0 takes precedence over top_k (vaura_model.py:818-823) */
  float cfg_scale;
  uint64_t seed;         /* Philox key; counter = (clip_id, offset, codebook, stream_id) */
  const int32_t* clip_ids; /* device [B] or NULL (=> 0..B-1) */
  int32_t* sequence;     /* device [B][K][S] int32, in/out; -1 = not generated yet (vaura_model.py:482) */
  const float* cond_rows;/* device [rows][cond_tokens+1][cond_dim] from vaura_sampler_cond_project */
  float* logits_out;     /* optional device [S][B][K][V] post-CFG logits, entry [offset] = the logits that
                            produced column offset; NULL to skip */
  int32_t precision;     /* VAURA_PRECISION_* */
  uint32_t stream_id;    /* 4th Philox counter word.  Calls that would otherwise repeat (seed, clip id, column,
                            codebook) - the windows of a chunked long clip, successive batches without clip ids -
                            pass different values so that they do not draw the same uniforms (the reference's
                            torch.multinomial advances a global generator, utils/utils.py:139-160) */
} vaura_generate_params;

int vaura_sampler_generate(vaura_sampler* s, const vaura_generate_params* p, const vaura_kv_cache* kv,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Teacher-forced forward = Transformer.forward (llama.py:520-539): logits for every position.
 * sequence [rows][K][S] int32 (no -1), cond_rows as above, logits_out [rows][K][S][V] f32.          */
int vaura_sampler_forward(vaura_sampler* s, const int32_t* sequence, const float* cond_rows, int32_t rows,
                          int32_t S, float* logits_out, const vaura_kv_cache* kv, int32_t precision,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Sampling stage alone (utils/utils.py:139-196 + vaura_model.py:810-825) for distribution tests:
 * logits [rows*(use_cfg?2:1)][K][V] f32 -> tokens_out [rows][K] int32; probs_out (optional)
 * [rows][K][V] = filtered, renormalised probabilities in vocabulary order.                         */
int vaura_sample_logits(const float* logits, int32_t rows, int32_t K, int32_t V, int32_t use_cfg,
                        float cfg_scale, int32_t use_sampling, float temp, int32_t top_k, float top_p,
                        uint64_t seed, const int32_t* clip_ids, int32_t offset, int32_t* tokens_out,
                        float* probs_out, void* stream);

/* Device time of the decode-step launches of the last vaura_sampler_generate call on this handle: CUDA events recorded on
 * the caller's stream right before the first and right after the last step launch (first pass, memsets and host glue are
 * outside).  Synchronises on the second event.  Used by bench.py for the step kernel's average launch duration. */
int vaura_sampler_last_loop_ms(vaura_sampler* s, float* ms_out, int32_t* steps_out);

/* ---- codec decode, replaces DacModelWrapper.decode (models/modules/dac/model.py:41-48) ------------ */
typedef struct {
  int32_t latent_dim;    /* 1024 */
  int32_t decoder_dim;   /* 1536 */
  int32_t n_blocks;      /* 4 */
  int32_t rates[8];      /* 8,8,4,2 */
  int32_t n_codebooks;   /* 9 */
  int32_t codebook_size; /* 1024 */
} vaura_codec_dims;

/* Folded (weight-norm removed) fp16 weights, channels-last GEMM layouts (vaura_b200/weights.py):
 *   code_tables [Kc][codebook_size][latent] f16 = codebook_k @ out_proj_k^T + bias_k (every table carries its own bias)
 *   conv weights  [taps][Cout][Cin] f16;  conv-transpose [stride][2][Cout][Cin] f16 (polyphase)
 *   biases / snake alphas f32.  `blob` is one device allocation; the offsets table indexes it.     */
typedef struct {
  const void* blob;
  const int64_t* offsets; /* host array, slot order: vaura_b200/weights.py:pack_codec, listed in csrc/cabi.cu above struct vaura_codec */
  int32_t n_offsets;
} vaura_codec_weights;

int vaura_codec_create(const vaura_codec_dims* dims, const vaura_codec_weights* w, vaura_codec** out);
void vaura_codec_destroy(vaura_codec* c);
size_t vaura_codec_workspace_bytes(const vaura_codec* c, int32_t batch, int32_t frames);
/* codes [B][Kc][T] int32 -> wav [B][T*hop] f16 */
int vaura_codec_decode(vaura_codec* c, const int32_t* codes, int32_t batch, int32_t frames, uint16_t* wav_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- codec encode, replaces DacModelWrapper.encode (models/modules/dac/model.py:30-39: preprocess + dac `encode`,
 *      codes only).  Same dims struct as the decoder plus encoder_dim; its own weight blob (vaura_b200/weights.py:pack_codec_encoder):
 *      0 conv_in W f32 [C0][7]  1 conv_in bias;  per block i (base 2 + 21 i): three residual units (+6 j: alpha1 | conv7 W f16
 *      [7][C][C] | bias | alpha2 | conv1 W f16 [1][C][C] | bias), +18 block snake alpha, +19 strided conv as three taps over
 *      frames of `stride` samples W f16 [3][2C][stride*C], +20 bias;  tail (base 2 + 21 n): final snake alpha | conv k3 W f16
 *      [3][latent][Cn] | bias | in_proj W f32 [Kc][Dc][latent] | in_proj bias [Kc][Dc] | l2-normalised codebooks f32
 *      [Kc][V][Dc] | out_proj(codebook) tables f32 [Kc][V][latent] | tap-offset tables int32.                          */
typedef struct vaura_codec_encoder vaura_codec_encoder;
int vaura_codec_encoder_create(const vaura_codec_dims* dims, int32_t encoder_dim, int32_t codebook_dim,
                               const vaura_codec_weights* w, vaura_codec_encoder** out);
void vaura_codec_encoder_destroy(vaura_codec_encoder* c);
size_t vaura_codec_encoder_workspace_bytes(const vaura_codec_encoder* c, int32_t batch, int32_t samples);
/* wav [B][samples] f32 (samples a multiple of the hop length: the caller pads, DAC.preprocess) -> codes [B][Kc][samples/hop]
 * int32; latent_out (optional) [B][samples/hop][latent] f16 = the encoder output before quantisation */
int vaura_codec_encode(vaura_codec_encoder* c, const float* wav, int32_t batch, int32_t samples, int32_t* codes_out,
                       uint16_t* latent_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Segment-AVCLIP visual features, replaces MotionFormer.forward in the shipped configuration
 *      (models/modules/feature_extractors/avclip/motionformer.py:252-342; extract_features, factorize_space_time,
 *      agg_space_module = TransformerEncoderLayer, agg_time_module = Identity, add_global_repr = False;
 *      motionformer_src/divided_224_16x4.yaml) ---------------------------------------------------------------- */
typedef struct vaura_avclip vaura_avclip;     /* opaque: video segments -> per-frame features */
typedef struct {
  int32_t embed_dim;   /* 768 */
  int32_t depth;       /* 12 */
  int32_t num_heads;   /* 12 (head width must be 64) */
  int32_t mlp_ratio;   /* 4 */
  int32_t img_size;    /* 224 */
  int32_t patch_size;  /* 16 */
  int32_t in_chans;    /* 3 */
  int32_t frames;      /* 16 frames per segment */
  int32_t tubelet;     /* 2 (VIT.PATCH_SIZE_TEMP) */
} vaura_avclip_dims;

/* bf16 matrices [out][in], fp32 vectors; `blob` is one device allocation indexed by the host offsets table.  Slot order
 * (vaura_b200/weights.py:pack_avclip): 0 tubelet W [D][C*2*16*16]  1 tubelet bias  2 position table f32 [t*n][D]
 * (pos_embed[1 + patch] + temp_embed[frame])  3 CLS row f32 [D] (cls_token + pos_embed[0]);
 * per block i (base 4 + 18 i): norm3 w,b | timeattn qkv W,b | timeattn proj W,b | norm1 w,b | attn qkv W,b | attn proj W,b |
 * norm2 w,b | fc1 W,b | fc2 W,b;  tail (base 4 + 18 depth): norm w,b | agg cls_token | agg norm1 w,b | in_proj W,b |
 * out_proj W,b | agg norm2 w,b | linear1 W,b | linear2 W,b.                                                         */
typedef struct {
  const void* blob;
  const int64_t* offsets;
  int32_t n_offsets;
} vaura_avclip_weights;

int vaura_avclip_create(const vaura_avclip_dims* dims, const vaura_avclip_weights* w, vaura_avclip** out);
void vaura_avclip_destroy(vaura_avclip* a);
/* workspace for processing `segments` segments at a time (forward walks its input in chunks that fit) */
size_t vaura_avclip_workspace_bytes(const vaura_avclip* a, int32_t segments);
/* frames [segments][C][T][H][W] f32 (normalised RGB) -> features [segments][T / tubelet][D] f32 */
int vaura_avclip_forward(vaura_avclip* a, const float* frames, int32_t segments, float* features_out, void* workspace,
                         size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAURA_B200_H */
