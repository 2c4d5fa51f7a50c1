// Internal launch interfaces between the C-ABI glue (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "knobs.h"

namespace vaura {

enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_QKV = 3,
       // fp32-activation prefill on the tensor cores (A operand = fp32 activations split into three bf16 terms):
       EPI_QKV_F32 = 4,        // RoPE, q -> fp32 [R][d], K/V -> fp32 pages
       EPI_SWIGLU_SPLIT3 = 5   // silu(w1 x) * w3 x in fp32, stored as three bf16 terms [R][3F] (the next GEMM's A operand)
};

struct EmbedArgs {
  const int32_t* seq;       // [B][K][S]
  const float* cond_rows;   // [rows][cond_tokens+1][cond_dim]
  const float* tables;      // [K][V+1][d - cond_dim]
  float* h;                 // [rows*npos][d]
  const StepState* state;   // nullptr -> use pos0
  int pos0, npos;
  int batch, K, S, vocab, d_model, cond_dim, cond_tokens, atpvf;
};

struct GemvArgs {
  const uint16_t* W;   // [N][K] bf16
  const float* x;      // row r at x + r*ldx
  const float* norm_w; // RMSNorm weight (NORM variants)
  float* out;          // row r at out + r*ldo
  const float* rope;   // EPI_QKV
  KvView kv;           // EPI_QKV
  const StepState* state;
  int pos0, npos;
  int N, K, R, ldx, ldo;
  int layer, d_model;
  int perm_S, perm_V;  // EPI_STORE: when perm_S > 0, row = b*perm_S + j is stored at [b][n / perm_V][j][n % perm_V]
  float eps;
};

struct AttnArgs {
  const float* q;   // [rows*npos][d]
  float* out;       // [rows*npos][d]
  uint16_t* out3;   // optional: the same rows as three bf16 terms [rows*npos][3 d] (hi | mid | lo), see split3()
  int out_terms;    // 0 / 3: three terms; 1: out3 is [rows*npos][d], the rows rounded to bf16 (bf16-activation prefill)
  KvView kv;
  const StepState* state;
  int pos0, npos;
  int layer, d_model;
  float scale;
};

struct SampleArgs {
  const float* logits;   // [rows_eff][K][V]; cond rows first, then uncond rows
  int32_t* sequence;     // [B][K][S] or nullptr (then tokens_out is written)
  int32_t* tokens_out;   // [B][K] optional
  float* probs_out;      // [B][K][V] optional
  float* logits_out;     // [S][B][K][V] optional (entry [offset])
  const int32_t* clip_ids;
  StepState* state;      // nullptr -> use offset, no advance
  int offset;
  int B, K, V, S, T;
  int use_cfg, use_sampling, top_k;
  float cfg_scale, temp, top_p;
  uint32_t seed_lo, seed_hi;
  uint32_t stream_id;    // 4th Philox counter word: distinguishes calls that reuse (clip id, column, codebook)
};

struct ConvArgs {
  const __half* in;        // [B][Tin][Cin]
  const __half* W;         // [nphase][ntaps][Cout][Cin]
  const int* tap_off;      // device [nphase][ntaps]: input index = q + off
  const float* bias;       // [Cout]
  const float* alpha;      // [Cout] Snake alpha for out_act
  const __half* residual;  // [B][Tout][Cout] or nullptr
  __half* out_raw;         // [B][Tout][Cout] or nullptr
  __half* out_act;         // [B][Tout][Cout] or nullptr (Snake applied)
  int Tin, Tq, Tout, Cin, Cout, ntaps, nphase, ostride;  // t_out = q*ostride + phase
};

cudaError_t launch_from_codes(const int32_t* codes, const __half* tables, __half* z, int B, int Kc, int T, int Vc,
                              int latent, cudaStream_t st);
cudaError_t launch_conv_gemm(const ConvArgs& a, int B, cudaStream_t st);
cudaError_t launch_conv_out_tanh(const __half* in, const float* W, const float* bias, __half* wav, int B, int T, int C,
                                 cudaStream_t st);
cudaError_t launch_enc_conv_in(const float* wav, const float* W, const float* bias, const float* alpha, __half* out_raw,
                               __half* out_act, int B, int L, int C, cudaStream_t st);
cudaError_t launch_rvq_encode(const __half* z, const float* w_in, const float* b_in, const float* cb_norm, const float* tables,
                              int32_t* codes, int B, int Kc, int T, int Vc, int latent, int Dc, cudaStream_t st);
struct LinearTcArgs {
  const void* A;        // bf16 [R][lda]
  const void* W;        // bf16 [N][K]
  float* out_f32;
  void* out_bf16;
  const float* rope;
  KvView kv;
  const StepState* state;
  int R, N, K, lda, ldo, block_n, epi;
  int perm_S, perm_V, pos0, npos, layer, d_model;
  int ksplit;  // EPI_RESID only: split K over this many CTAs, partials reduced with red.global.add
  int pdl;     // programmatic dependent launch (see TcShape::pdl)
  int w_k;     // 0: W is [N][K].  > 0: W is [N][w_k] and A is [R][K] with K a multiple of w_k: the K blocks of W are
               // re-read for every w_k-wide section of A (A = the bf16 terms of a split fp32 operand side by side)
  int aux;     // EPI_SWIGLU_SPLIT3: F (distance between the three terms of one output element)
};
cudaError_t launch_linear_tc(const LinearTcArgs& a, cudaStream_t st);
// bf16 path helpers (decode_bf16.cu)
cudaError_t launch_rmsnorm_bf16(const float* h, const float* w, void* out_bf16, int R, int D, size_t ldh, float eps, int pdl,
                                cudaStream_t st);
struct AttnBf16Args {
  const void* q;   // bf16 [rows*npos][d]
  void* out;       // bf16 [rows*npos][d]
  KvView kv;
  const StepState* state;
  int pos0, npos, layer, d_model;
  float scale;
  int pdl;
};
cudaError_t launch_attn_bf16(const AttnBf16Args& a, int nhead, int rows, cudaStream_t st);
// fp32 -> three bf16 terms (x = t1 + t2 + t3 to 24 significant bits): the A operand of the fp32-equivalent tensor-core GEMMs
cudaError_t launch_rmsnorm_split3(const float* h, const float* w, void* out3, int R, int D, size_t ldh, float eps, cudaStream_t st);

// fused decode step of the bf16 path, rows <= 64 (gemm_tcgen05.cu: decode_step_fused_bf16)
struct FusedStepArgs {
  const float *attn_norm, *ffn_norm, *final_norm, *rope;
  float* h;                  // [R][D] residual stream (written by embed_kernel before this kernel)
  __nv_bfloat16 *xn, *q, *attn, *act;  // [R][D], [R][D], [R][D], [R][F]
  float* logits;             // [R][NH]
  float* part;               // [max(wo_ksplit, w2_ksplit)][R][D] split-K partial sums of the two residual GEMMs (fp32)
  KvView kv;                 // bf16 pages
  StepState* state;
  int R, L, D, F, H, NH;     // rows, layers, d_model, ffn, heads, K*V logits per row
  int wo_ksplit, w2_ksplit;
  float eps, scale;
  // raw weight pointers (the GEMM operands themselves go through tensor maps): L2 prefetch of the next phase's tile
  const __nv_bfloat16 *w_qkv, *w_o, *w_13, *w_2, *w_heads;
  int l2_prefetch;  // 0 = off
  unsigned long long* timing;  // optional: timestamps (ns) of CTA `timing_cta` before / after every device-wide barrier
  unsigned long long* step_times;  // [kMaxCtx + 16]: %globaltimer at the start of the launch that samples column `offset`
  int timing_cta;
  // fuse_io: the kernel also builds the embedding rows (first phase) and samples / writes back the tokens (last phase)
  int fuse_io;
  const int32_t* seq;       // [B][K][S]
  const float* cond_rows;   // [rows][cond_tokens+1][cond_dim]
  const float* tables;      // [K][V+1][d - cond_dim]
  int batch, Kc, S, vocab, cond_dim, cond_tokens, atpvf;
  SampleArgs sample;        // state = nullptr: the column comes from this kernel's own state read
};

bool fused_step_supported(int R, int D, int F, int NH);
constexpr int kFusedKsplit = 6;  // K slices of the wo / w2 tiles of decode_step_fused_bf16 (24 x 6 = 144 tiles: one per CTA)
size_t fused_part_bytes(int R, int D);
cudaError_t launch_decode_fused_bf16(const FusedStepArgs& a, const void* wqkv, const void* wo, const void* w13, const void* w2,
                                     const void* w_heads, cudaStream_t st);

// fused decode step, second design (decode_fused2.cu: decode_step_fused2): rows <= 64, clusters of two CTAs
struct Fused2Args {
  const float *attn_norm, *ffn_norm, *final_norm, *rope;
  float* h_t;                // [D][64] residual stream, transposed (feature-major), fp32
  __nv_bfloat16* hb;         // [R][D] bf16(h * next norm weight): the B operand of the next GEMM
  __nv_bfloat16* q_t;        // [D][64] RoPE'd q, feature-major
  __nv_bfloat16 *attn, *act; // [R][D], [R][F]
  float* ssq_part;           // [<= 64 slots][64] partial sums of squares of the residual rows
  float* w2_part;            // [D/64][3][2][32][64] partial w2 sums of the three K thirds
  unsigned* w2_cnt;          // [D/64][2] arrival counters of the K thirds (monotonic, start at a multiple of 3)
  float* logits;             // [R][NH]
  KvView kv;                 // bf16 pages
  StepState* state;
  int R, L, D, F, H, NH;
  float eps, scale;
  unsigned long long* timing;      // optional: %globaltimer before / after every device-wide barrier wait of CTA `timing_cta`
  unsigned long long* step_times;  // [kMaxCtx + 16]: %globaltimer at the start of the launch that samples column `offset`
  int timing_cta;
  int flags;                // bit 0: the weight ring does not run ahead of the device-wide barriers (experiment)
  const int32_t* seq;       // [B][K][S]
  const float* cond_rows;   // [rows][cond_tokens+1][cond_dim]
  const float* tables;      // [K][V+1][d - cond_dim]
  int batch, Kc, S, vocab, cond_dim, cond_tokens, atpvf;
  SampleArgs sample;        // state = nullptr: the column comes from this kernel's own state read
};
bool fused2_supported(int R, int L, int D, int F, int H, int NH, int sms, int page_size, int max_pages);
size_t fused2_workspace_bytes(int D);
cudaError_t launch_decode_fused2(const Fused2Args& a, const void* wqkv, const void* wo, const void* w13, const void* w2,
                                 const void* w_heads, cudaStream_t st);
// [outer][rows][K] 16-bit tensor seen as (64, rows, K / 64, outer); one box = nblk K blocks of box_rows rows, landing in shared
// memory as nblk consecutive 128B-swizzled [box_rows x 64] sub-tiles (gemm_tcgen05.cu).  `map` is a CUtensorMap.
bool tc_make_map_kblocks(void* map, const void* base, uint64_t K, uint64_t rows, uint64_t outer, uint64_t row_stride_el,
                         uint64_t outer_stride_el, int box_rows, int nblk);

// ---- Segment-AVCLIP visual tower (avclip.cu) ----------------------------------------------------------------------------
enum { VIT_STORE_BF16 = 0, VIT_RESID_F32 = 1, VIT_PATCH = 2 };
struct VitLinearArgs {
  const void* A;       // bf16 [M][lda]
  const void* W;       // bf16 [N][K]
  const float* bias;   // [N]
  void* out_bf16;      // VIT_STORE_BF16
  float* out_f32;      // VIT_RESID_F32 (in place) / VIT_PATCH
  const float* pos;    // VIT_PATCH
  int M, N, K, lda, ldo, mode, gelu;
  int rows_in, rows_out, row_off;
};
cudaError_t launch_vit_linear(const VitLinearArgs& a, cudaStream_t st);
cudaError_t launch_vit_patchify(const float* frames, void* A, int S, int C, int T, int H, int W, int tub, int ps, cudaStream_t st);
cudaError_t launch_vit_broadcast_row(float* dst, const float* src, int D, int count, size_t row_stride, cudaStream_t st);
cudaError_t launch_vit_layernorm(const float* x, const float* g, const float* b, void* out_bf16, int rows, int D, float eps,
                                 cudaStream_t st);
cudaError_t launch_vit_final_norm_agg(const float* x, const float* agg_cls, const float* gf, const float* bf, const float* g1,
                                      const float* b1, void* out_bf16, int S, int t, int n, int D, float eps, cudaStream_t st);
cudaError_t launch_vit_time_attn(const void* qkv, void* out, int S, int t, int n, int heads, cudaStream_t st);
cudaError_t launch_vit_cls_attn(const void* qkv, void* out, int seqs, int len, int heads, int out_rows_per_seq, cudaStream_t st);
cudaError_t launch_vit_space_attn(const void* qkv, void* out, int S, int t, int n, int heads, cudaStream_t st);

// fused ResidualUnit (conv k7 -> Snake -> conv k1 -> + x) of the codec for narrow layers (gemm_tcgen05.cu: gemm_ru_fused_kernel)
struct RuArgs {
  const __half* act;       // [B][T][C] Snake(x, alpha1): the unit's activated input
  const __half* x;         // [B][T][C] the unit's raw input (residual)
  const __half *W7, *W1;   // [7][C][C], [1][C][C]
  const float *bias7, *alpha2, *bias1, *alpha_next;
  __half* out_raw;         // [B][T][C] x + unit(x) or nullptr (may alias x: every element is read and written by one thread)
  __half* out_act;         // [B][T][C] Snake(out, alpha_next)
  int T, C;
};
bool ru_fused_supported(int C);
cudaError_t launch_ru_fused(const RuArgs& a, const int* taps7_host, int B, cudaStream_t st);
bool conv_tc_supported(int Cin, int Cout, int ntaps, int nphase);
cudaError_t launch_conv_tc(const ConvArgs& a, const int* tap_off_host, int B, cudaStream_t st);
cudaError_t init_decode_kernels();

// persistent decode-step kernel (decode_persistent.cu)
struct PersistArgs {
  const uint16_t *wqkv, *wo, *w13, *w2, *w_heads;
  const uint8_t* wstream;       // fragment-ordered weight streams (cluster variant, decode_cluster.cu)
  long long* xfix;              // cluster variant: residual buffers [2L+1][rows][D] (word = fixed-point sum << 6 | count)
  const float *attn_norm, *ffn_norm, *final_norm, *tok_tables, *rope;
  const int32_t* seq;
  const float* cond_rows;
  float *h, *q, *act, *logits, *attn_part;
  KvView kv;
  StepState* state;
  unsigned long long* timing;  // optional: phase timestamps (ns) of CTA `timing_cta`
  unsigned long long* step_times;  // [kMaxCtx + 16]: %globaltimer at the start of the launch that samples column `offset`
  int timing_cta;
  int pace_cycles;  // cluster variant: units of L2 prefetch ahead of the shared-memory fill (< 0: default)
  int tail_units;   // cluster variant: units of the NEXT step's stream head prefetched into L2 during the sampling tail
  SampleArgs sample;
  int L, D, F, H, Kc, V, S, batch, cond_dim, cond_tokens, atpvf, slot_cap, prefetch_ahead;
  float eps, scale;
};

bool persistent_supported(int rows, int D, int F, int page_size);
size_t persistent_attn_part_bytes(int rows, int H);
cudaError_t launch_decode_persistent(PersistArgs& a, int rows, cudaStream_t st);

// cluster variant (decode_cluster.cu): 32 clusters x 4 CTAs, mma.sync from fragment-ordered weight streams
bool cluster_supported(int rows, int L, int D, int F, int H, int head_rows, int page_size, int cond_dim, int max_ctx);
size_t cluster_stream_bytes(int L);
size_t cluster_xfix_bytes(int rows, int L);
// false when this device cannot keep all 32 clusters of 4 co-resident (fewer than 128 SMs visible): the caller then
// uses decode_step_persistent instead
bool cluster_launchable(int rows, bool timing);
cudaError_t launch_decode_cluster(const PersistArgs& a, int rows, cudaStream_t st);

cudaError_t launch_embed(const EmbedArgs& a, int rows, cudaStream_t st);
cudaError_t launch_gemv(int epi, bool norm, const GemvArgs& a, cudaStream_t st);
cudaError_t launch_attn(const AttnArgs& a, int nhead, int rows, cudaStream_t st);
cudaError_t launch_cond_project(const float* feats, const float* fc1, const float* fc2, const float* empty, float* out,
                                int rows, int tv, int cin, int C, cudaStream_t st);
cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st);
cudaError_t launch_set_state(StepState* s, int offset, cudaStream_t st);

}  // namespace vaura
