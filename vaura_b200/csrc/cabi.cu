// extern "C" boundary (include/vaura_b200.h): argument checking, workspace carving, launch
// sequencing and CUDA-graph replay of the decode step.  No device memory is allocated here.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/vaura_b200.h"
#include "kernels.h"

using namespace vaura;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// kernels launched (see vaura_launch_count).  One host thread drives a handle (include/vaura_b200.h); the capture flags are
// per thread so that two threads driving two handles do not see each other's capture.
static unsigned long long g_launches = 0;
static thread_local unsigned long long g_capture_nodes = 0;
static thread_local bool g_capturing = false;
#define LAUNCHED(n) do { if (g_capturing) g_capture_nodes += (n); else g_launches += (n); } while (0)

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(VAURA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));   \
  } while (0)

#define CUL(expr) do { CU(expr); LAUNCHED(1); } while (0)

// Everything a captured decode-step graph bakes in: pointers, shapes and sampling parameters.  A generate() call whose key
// equals the cached one replays the instantiated graph instead of capturing, instantiating and re-encoding tensor maps.
// VAURA_DETERMINISTIC=1: the bf16 step kernel adds its split-K partial sums in a fixed order instead of with float atomics,
// so a repeated call returns the same bits (reproducibility)
static int deterministic_mode() { return knobs().deterministic; }

struct GraphKey {
  vaura_generate_params p;
  vaura_kv_cache kv;
  void* workspace;
  int precision, device;
  Knobs knobs;  // a step captured under other knob settings is not replayed
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};

struct vaura_sampler {
  vaura_sampler_dims d;
  vaura_sampler_weights w;
  cudaGraphExec_t graph_exec = nullptr;  // decode-step graph of the last generate() call that captured one
  GraphKey graph_key;                    // valid when graph_exec != nullptr
  unsigned long long graph_nodes = 0;    // kernel nodes of graph_exec
  cudaStream_t capture_stream = nullptr; // capture origin only (the legacy default stream cannot be captured);
                                         // nothing ever executes on it, graphs are launched on the caller's stream
  cudaEvent_t loop_ev[2] = {nullptr, nullptr};  // recorded on the caller's stream around the decode-step launches of the
  int loop_steps = 0;                           // last generate() call (vaura_sampler_last_loop_ms)
};

extern "C" int vaura_version(void) { return 1; }
extern "C" const char* vaura_arch(void) { return "sm_100a"; }
extern "C" const char* vaura_last_error(void) { return g_err; }
extern "C" unsigned long long vaura_launch_count(void) { return g_launches; }

extern "C" int vaura_linear_bf16(const uint16_t* A, const uint16_t* W, float* y, int32_t R, int32_t N, int32_t K,
                                 int32_t block_n, void* stream) {
  refresh_knobs();
  if (!A || !W || !y || R <= 0 || N <= 0 || K <= 0 || (K % 64) || block_n <= 0 || (N % block_n))
    return fail(VAURA_ERR_INVALID, "bad argument");
  LinearTcArgs g{};
  g.A = A; g.lda = K; g.W = W; g.N = N; g.K = K; g.R = R; g.epi = EPI_STORE; g.out_f32 = y; g.ldo = N; g.block_n = block_n;
  g.npos = 1; g.ksplit = 1; g.pdl = 0;
  CUL(launch_linear_tc(g, (cudaStream_t)stream));
  return VAURA_OK;
}

extern "C" int vaura_gemv_bf16w(const uint16_t* W, const float* x, float* y, int32_t N, int32_t K, int32_t R, void* stream) {
  refresh_knobs();
  if (!W || !x || !y || N <= 0 || K <= 0 || R <= 0 || (N & 1) || (K & 7)) return fail(VAURA_ERR_INVALID, "bad argument");
  CU(init_decode_kernels());
  GemvArgs g{};
  g.W = W; g.x = x; g.out = y; g.N = N; g.K = K; g.R = R; g.ldx = K; g.ldo = N; g.npos = 1;
  CUL(launch_gemv(EPI_STORE, false, g, (cudaStream_t)stream));
  return VAURA_OK;
}

extern "C" int vaura_sampler_create(const vaura_sampler_dims* dims, const vaura_sampler_weights* weights,
                                    vaura_sampler** out) {
  refresh_knobs();
  if (!dims || !weights || !out) return fail(VAURA_ERR_INVALID, "null argument");
  const vaura_sampler_dims& d = *dims;
  if (d.nhead <= 0 || d.d_model % d.nhead != 0 || d.d_model / d.nhead != kHeadDim)
    return fail(VAURA_ERR_UNSUPPORTED, "head_dim %d unsupported (kernels are built for %d)",
                d.nhead > 0 ? d.d_model / d.nhead : -1, kHeadDim);
  if (d.d_model % 64 || d.ffn_dim % 64 || d.cond_dim % 4 || (d.d_model - d.cond_dim) % 4)
    return fail(VAURA_ERR_UNSUPPORTED, "d_model/ffn_dim must be multiples of 64");
  if (d.vocab != 1024) return fail(VAURA_ERR_UNSUPPORTED, "vocab %d unsupported (sampling kernel is built for 1024)", d.vocab);
  if (d.num_codebooks < 1 || d.num_codebooks > 16) return fail(VAURA_ERR_UNSUPPORTED, "num_codebooks must be 1..16");
  if (d.block_size > kMaxCtx) return fail(VAURA_ERR_UNSUPPORTED, "block_size > %d", kMaxCtx);
  // every embedding path divides the position by audio_tokens_per_video_frame (llama.py:555-586); the reference raises on
  // the host for a non-positive value
  if (d.audio_tokens_per_video_frame < 1)
    return fail(VAURA_ERR_INVALID, "audio_tokens_per_video_frame = %d must be >= 1", d.audio_tokens_per_video_frame);
  if (d.cond_tokens < 1 || d.cond_dim < 4 || d.cond_in < 1) return fail(VAURA_ERR_INVALID, "bad conditioning shape");
  int dev = 0, major = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(VAURA_ERR_CUDA, "device compute capability %d.x is not sm_100; no fallback path exists", major);
  CU(init_decode_kernels());
  vaura_sampler* s = new (std::nothrow) vaura_sampler();
  if (!s) return fail(VAURA_ERR_INVALID, "out of host memory");
  s->d = d;
  s->w = *weights;
  cudaError_t ce = cudaStreamCreateWithFlags(&s->capture_stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete s; return fail(VAURA_ERR_CUDA, "cudaStreamCreateWithFlags: %s", cudaGetErrorString(ce)); }
  *out = s;
  return VAURA_OK;
}

extern "C" void vaura_sampler_destroy(vaura_sampler* s) {
  if (!s) return;
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
  for (cudaEvent_t e : s->loop_ev) if (e) cudaEventDestroy(e);
  delete s;
}

extern "C" int vaura_sampler_cond_project(vaura_sampler* s, const float* feats, int32_t rows, int32_t tv,
                                          float* rows_out, void* stream) {
  refresh_knobs();
  if (!s || !feats || !rows_out || rows <= 0 || tv <= 0) return fail(VAURA_ERR_INVALID, "bad argument");
  CUL(launch_cond_project(feats, s->w.fc1, s->w.fc2, s->w.empty_video_emb, rows_out, rows, tv, s->d.cond_in,
                         s->d.cond_dim, (cudaStream_t)stream));
  return VAURA_OK;
}

// ---- workspace layout (fp32act) ---------------------------------------------------------------------
struct Workspace {
  StepState* state;
  unsigned long long* timing;
  float *h, *q, *attn, *act, *logits, *attn_part;        // fp32act path
  long long* xfix;                                       // fp32act path, cluster decode kernel (rows <= 2)
  __nv_bfloat16 *xn_b, *q_b, *attn_b, *act_b;            // bf16 path (h and logits stay fp32)
  __nv_bfloat16 *x3, *act3;                              // fp32act path, tensor-core prefill: operands as three bf16 terms
  char* f2;                                              // bf16 path, decode_step_fused2: h_t | q_t | ssq_part | w2_part | w2_cnt
  float* part;                                           // bf16 path, decode_step_fused_bf16: split-K partial sums of wo / w2
  int prefill_terms;                                     // fp32act path, tensor-core prefill: 3 (fp32-equivalent) or 1 (bf16 operands)
  size_t bytes;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static Workspace carve(const vaura_sampler_dims& d, int rows, int max_pos, int precision, void* base) {
  Workspace w{};
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n) {
    char* r = p ? p + off : nullptr;
    off += align_up(n);
    return r;
  };
  const size_t R = (size_t)rows * max_pos;
  w.state = (StepState*)take(sizeof(StepState));
  w.timing = (unsigned long long*)take(16384);  // debug: persistent-kernel phase timestamps (workspace bytes [256, 16640))
  w.h = (float*)take(R * d.d_model * 4);
  w.logits = (float*)take((size_t)rows * d.num_codebooks * d.vocab * 4);
  if (precision == VAURA_PRECISION_BF16) {
    w.xn_b = (__nv_bfloat16*)take(R * d.d_model * 2);
    w.q_b = (__nv_bfloat16*)take(R * d.d_model * 2);
    w.attn_b = (__nv_bfloat16*)take(R * d.d_model * 2);
    w.act_b = (__nv_bfloat16*)take(R * d.ffn_dim * 2);
    w.f2 = (char*)take(fused2_workspace_bytes(d.d_model));
    w.part = (float*)take(fused_part_bytes(rows <= 128 ? rows : 1, d.d_model));
  } else {
    w.q = (float*)take(R * d.d_model * 4);
    w.attn = (float*)take(R * d.d_model * 4);
    w.act = (float*)take(R * d.ffn_dim * 4);
    w.attn_part = (float*)take(persistent_attn_part_bytes(rows, d.nhead));
    w.xfix = (long long*)take(cluster_xfix_bytes(rows <= 2 ? rows : 1, d.num_layers));
    if (R >= 16) {  // multi-position passes (prefill, teacher-forced forward) run their GEMMs on the tensor cores
      w.x3 = (__nv_bfloat16*)take(R * 3 * d.d_model * 2);
      w.act3 = (__nv_bfloat16*)take(R * 3 * d.ffn_dim * 2);
    }
  }
  w.bytes = off;
  return w;
}

// AUTO: the tensor-core path needs at least one 16-row UMMA N/M granule of real work to pay off; below that the
// decode step is a pure weight stream and the fp32-activation path gives bit-stable greedy tokens.
// From 3 rows the fused bf16 step kernel (~1 ms) is faster than the fp32-activation paths (1.2-2.9 ms at 3..15 rows,
// profiles/scripts/rows_sweep.py); a call that samples has no bit-exactness contract, so AUTO takes it there as well.
static int resolve_precision(int precision, int rows, bool sampling) {
  if (precision != VAURA_PRECISION_AUTO) return precision;
  return (rows >= 16 || (rows >= 3 && sampling)) ? VAURA_PRECISION_BF16 : VAURA_PRECISION_FP32ACT;
}

extern "C" size_t vaura_sampler_workspace_bytes(const vaura_sampler* s, int32_t rows, int32_t max_positions,
                                                int32_t precision) {
  if (!s || rows <= 0 || max_positions <= 0) return 0;
  if (precision == VAURA_PRECISION_AUTO) {  // the mode depends on the call (sampling or not): enough for either
    const size_t a = carve(s->d, rows, max_positions, VAURA_PRECISION_BF16, nullptr).bytes;
    const size_t b = carve(s->d, rows, max_positions, VAURA_PRECISION_FP32ACT, nullptr).bytes;
    return a > b ? a : b;
  }
  return carve(s->d, rows, max_positions, precision, nullptr).bytes;
}

static KvView kv_view(const vaura_kv_cache* kv, int nhead) {
  KvView v;
  v.pages = kv->pages;
  v.page_table = kv->page_table;
  v.num_pages = kv->num_pages;
  v.page_size = kv->page_size;
  v.max_pages_per_seq = kv->max_pages_per_seq;
  v.nhead = nhead;
  return v;
}

static int transformer_pass_tc3(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                                const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                                const KvView& kv, float* logits_dst, bool logits_all, cudaStream_t st);
static int transformer_pass_tc1(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                                const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                                const KvView& kv, float* logits_dst, cudaStream_t st);

// One transformer pass over `npos` new positions per sequence row (fp32act path).
//   state != nullptr: positions come from the device-resident loop state (graph replay);
//   otherwise pos0 is the first new position.
//   logits_all: write logits of every position to logits_dst [rows*npos][K*V]; else only the last
//   position of each row to logits_dst [rows][K*V].
static int transformer_pass(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                            const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                            const KvView& kv, float* logits_dst, bool logits_all, cudaStream_t st) {
  const vaura_sampler_dims& d = s->d;
  const vaura_sampler_weights& w = s->w;
  const int R = rows * npos;
  {
    const bool tc = knobs().prefill_tc && npos > 1 && R >= 16 && ws.x3 && d.d_model % 64 == 0 && d.ffn_dim % 64 == 0;
    if (tc && ws.prefill_terms == 1 && !logits_all)
      return transformer_pass_tc1(s, ws, seq, batch, S, cond_rows, rows, npos, pos0, state, kv, logits_dst, st);
    if (tc)
      return transformer_pass_tc3(s, ws, seq, batch, S, cond_rows, rows, npos, pos0, state, kv, logits_dst, logits_all, st);
  }
  EmbedArgs e{};
  e.seq = seq; e.cond_rows = cond_rows; e.tables = w.tok_tables; e.h = ws.h; e.state = state; e.pos0 = pos0;
  e.npos = npos; e.batch = batch; e.K = d.num_codebooks; e.S = S; e.vocab = d.vocab; e.d_model = d.d_model;
  e.cond_dim = d.cond_dim; e.cond_tokens = d.cond_tokens; e.atpvf = d.audio_tokens_per_video_frame;
  CUL(launch_embed(e, R, st));
  const size_t D = d.d_model, F = d.ffn_dim;
  for (int l = 0; l < d.num_layers; ++l) {
    GemvArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.R = R; g.layer = l; g.d_model = d.d_model; g.eps = d.norm_eps;
    g.kv = kv; g.rope = w.rope;
    // attention_norm -> wqkv -> RoPE -> KV append
    g.W = w.wqkv + (size_t)l * 3 * D * D; g.x = ws.h; g.ldx = D; g.norm_w = w.attn_norm + l * D;
    g.out = ws.q; g.ldo = D; g.N = 3 * D; g.K = D;
    CUL(launch_gemv(EPI_QKV, true, g, st));
    AttnArgs a{};
    a.q = ws.q; a.out = ws.attn; a.kv = kv; a.state = state; a.pos0 = pos0; a.npos = npos; a.layer = l;
    a.d_model = d.d_model; a.scale = 1.0f / sqrtf((float)kHeadDim);
    CUL(launch_attn(a, d.nhead, R, st));
    // wo + residual
    g.W = w.wo + (size_t)l * D * D; g.x = ws.attn; g.ldx = D; g.out = ws.h; g.ldo = D; g.N = D; g.K = D;
    CUL(launch_gemv(EPI_RESID, false, g, st));
    // ffn_norm -> w1|w3 -> silu*mul
    g.W = w.w13 + (size_t)l * 2 * F * D; g.x = ws.h; g.ldx = D; g.norm_w = w.ffn_norm + l * D;
    g.out = ws.act; g.ldo = F; g.N = 2 * F; g.K = D;
    CUL(launch_gemv(EPI_SWIGLU, true, g, st));
    // w2 + residual
    g.W = w.w2 + (size_t)l * D * F; g.x = ws.act; g.ldx = F; g.out = ws.h; g.ldo = D; g.N = D; g.K = F;
    CUL(launch_gemv(EPI_RESID, false, g, st));
  }
  GemvArgs g{};
  g.state = state; g.pos0 = pos0; g.npos = npos; g.layer = 0; g.d_model = d.d_model; g.eps = d.norm_eps;
  g.W = w.w_heads; g.norm_w = w.final_norm; g.N = d.num_codebooks * d.vocab; g.K = D; g.ldo = g.N; g.out = logits_dst;
  if (logits_all) { g.x = ws.h; g.ldx = D; g.R = R; g.perm_S = npos; g.perm_V = d.vocab; }
  else { g.x = ws.h + (size_t)(npos - 1) * D; g.ldx = (size_t)npos * D; g.R = rows; }
  CUL(launch_gemv(EPI_STORE, true, g, st));
  return VAURA_OK;
}

// The fp32-activation pass over many positions (prompt prefill of a chunked long clip, teacher-forced forward) with the four
// projections of a layer on the tensor cores: every fp32 A operand is split into three bf16 terms (x = t1 + t2 + t3 to 24
// significant bits, common.cuh: split3) laid side by side along K, the bf16 weights are multiplied with each term
// (LinearTcArgs::w_k re-reads their K blocks) and tcgen05 accumulates in fp32 - the products are the fp32 products, only
// the summation order differs from gemv_kernel.  The weights are read once per pass instead of once per 8 rows
// (launch_gemv_nb: ceil(166 / 8) = 21 passes for a 166-position prompt).  KV cache, q, attention and the residual stream
// stay fp32, so the decode steps that follow continue bit-compatibly.  Every output element has one owner: with one or two row
// tiles (a prompt window) K is split over a cluster of 2 or 4 CTAs whose partial sums meet in the owner's shared memory and are
// added in rank order (gemm_tc_kernel, CK > 1) - deterministic, no float atomics.
static int transformer_pass_tc3(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                                const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                                const KvView& kv, float* logits_dst, bool logits_all, cudaStream_t st) {
  const vaura_sampler_dims& d = s->d;
  const vaura_sampler_weights& w = s->w;
  const int R = rows * npos;
  EmbedArgs e{};
  e.seq = seq; e.cond_rows = cond_rows; e.tables = w.tok_tables; e.h = ws.h; e.state = state; e.pos0 = pos0;
  e.npos = npos; e.batch = batch; e.K = d.num_codebooks; e.S = S; e.vocab = d.vocab; e.d_model = d.d_model;
  e.cond_dim = d.cond_dim; e.cond_tokens = d.cond_tokens; e.atpvf = d.audio_tokens_per_video_frame;
  CUL(launch_embed(e, R, st));
  const size_t D = d.d_model, F = d.ffn_dim;
  // N tiles.  Every CTA streams its whole row tile of A (three terms) once, so the L2 -> SM traffic of a GEMM is
  // (tiles) x (3 x 128 + BN) x K x 2 bytes: wide tiles for the wide matrices (wqkv, w1|w3: measured L2-bound with narrow
  // ones), 64-wide tiles for the two 1536-wide ones so that they still cover 48 SMs.
  auto pick_bn = [&](int N) {
    if (N >= 4096 && N % 256 == 0) return 256;
    if (N >= 2048 && N % 128 == 0) return 128;
    return N % 64 == 0 ? 64 : 32;
  };
  for (int l = 0; l < d.num_layers; ++l) {
    LinearTcArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.R = R; g.layer = l; g.d_model = d.d_model; g.kv = kv; g.rope = w.rope;
    g.ksplit = 0; g.pdl = 0;  // 0 = auto: K split inside clusters when the tiles cover less than half of the SMs (prompt prefill)
    CUL(launch_rmsnorm_split3(ws.h, w.attn_norm + l * D, ws.x3, R, (int)D, D, d.norm_eps, st));
    g.A = ws.x3; g.lda = 3 * D; g.K = 3 * D; g.w_k = D; g.W = w.wqkv + (size_t)l * 3 * D * D; g.N = 3 * D; g.epi = EPI_QKV_F32;
    g.out_f32 = ws.q; g.ldo = D; g.block_n = pick_bn(3 * D);
    CUL(launch_linear_tc(g, st));
    AttnArgs a{};
    a.q = ws.q; a.out = ws.attn; a.out3 = reinterpret_cast<uint16_t*>(ws.x3); a.kv = kv; a.state = state; a.pos0 = pos0;
    a.npos = npos; a.layer = l; a.d_model = d.d_model; a.scale = 1.0f / sqrtf((float)kHeadDim);
    CUL(launch_attn(a, d.nhead, R, st));
    g.A = ws.x3; g.lda = 3 * D; g.K = 3 * D; g.w_k = D; g.W = w.wo + (size_t)l * D * D; g.N = D; g.epi = EPI_RESID;
    g.out_f32 = ws.h; g.ldo = D; g.block_n = pick_bn(D);
    CUL(launch_linear_tc(g, st));
    CUL(launch_rmsnorm_split3(ws.h, w.ffn_norm + l * D, ws.x3, R, (int)D, D, d.norm_eps, st));
    g.A = ws.x3; g.lda = 3 * D; g.K = 3 * D; g.w_k = D; g.W = w.w13 + (size_t)l * 2 * F * D; g.N = 2 * F; g.epi = EPI_SWIGLU_SPLIT3;
    g.out_bf16 = ws.act3; g.ldo = 3 * F; g.aux = (int)F; g.block_n = pick_bn(2 * F);
    CUL(launch_linear_tc(g, st));
    g.A = ws.act3; g.lda = 3 * F; g.K = 3 * F; g.w_k = F; g.W = w.w2 + (size_t)l * D * F; g.N = D; g.epi = EPI_RESID;
    g.out_f32 = ws.h; g.ldo = D; g.block_n = pick_bn(D);
    CUL(launch_linear_tc(g, st));
  }
  if (logits_all) {  // every position feeds the heads: the logits land in the reference layout [rows][K][S][V] (llama.py:504)
    LinearTcArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.d_model = d.d_model; g.ksplit = 1;
    CUL(launch_rmsnorm_split3(ws.h, w.final_norm, ws.x3, R, (int)D, D, d.norm_eps, st));
    g.A = ws.x3; g.lda = 3 * D; g.K = 3 * D; g.w_k = D; g.W = w.w_heads; g.N = d.num_codebooks * d.vocab; g.R = R;
    g.epi = EPI_STORE; g.out_f32 = logits_dst; g.ldo = g.N; g.perm_S = npos; g.perm_V = d.vocab; g.block_n = pick_bn(g.N);
    if (g.N % g.block_n) return fail(VAURA_ERR_UNSUPPORTED, "heads width %d is not a multiple of %d", g.N, g.block_n);
    CUL(launch_linear_tc(g, st));
  } else {  // only the last position of every sequence row feeds the heads: a few rows, the weight-streaming GEMV
    GemvArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.layer = 0; g.d_model = d.d_model; g.eps = d.norm_eps;
    g.W = w.w_heads; g.norm_w = w.final_norm; g.N = d.num_codebooks * d.vocab; g.K = D; g.ldo = g.N; g.out = logits_dst;
    g.x = ws.h + (size_t)(npos - 1) * D; g.ldx = (size_t)npos * D; g.R = rows;
    CUL(launch_gemv(EPI_STORE, true, g, st));
  }
  return VAURA_OK;
}

// Prompt prefill of a call that samples (no bit-exactness contract): the same pass with the GEMM operands
// rounded to bf16 once (one term instead of three: a third of the MMA work and of the activation bytes every CTA ingests).
// Residual stream, q, K/V pages and the attention arithmetic stay fp32, so the fp32-activation decode steps that follow read
// the cache they expect; the split-K residual GEMMs use float reductions (a sampling call is not reproducible bit for bit
// across precision modes anyway).  Logits stay within the bf16 tolerance (1e-2 of max |logit|), tests/test_gpu_prefill.py.
static int transformer_pass_tc1(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                                const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                                const KvView& kv, float* logits_dst, cudaStream_t st) {
  const vaura_sampler_dims& d = s->d;
  const vaura_sampler_weights& w = s->w;
  const int R = rows * npos;
  EmbedArgs e{};
  e.seq = seq; e.cond_rows = cond_rows; e.tables = w.tok_tables; e.h = ws.h; e.state = state; e.pos0 = pos0;
  e.npos = npos; e.batch = batch; e.K = d.num_codebooks; e.S = S; e.vocab = d.vocab; e.d_model = d.d_model;
  e.cond_dim = d.cond_dim; e.cond_tokens = d.cond_tokens; e.atpvf = d.audio_tokens_per_video_frame;
  CUL(launch_embed(e, R, st));
  const size_t D = d.d_model, F = d.ffn_dim;
  const int mt = (R + (R <= 64 ? 63 : 127)) / (R <= 64 ? 64 : 128);
  // N tile and K split so that the tiles of a GEMM cover the SMs about once
  auto tile_n = [&](int N, int want) {
    int bn = 128;
    while (bn > 32 && (N % bn != 0 || mt * (N / bn) < want)) bn >>= 1;
    return N % bn == 0 ? bn : 0;
  };
  auto split_k = [&](int N, int bn, int kblocks) {
    int ks = 148 / (mt * (N / bn));
    if (ks > 6) ks = 6;
    if (ks > kblocks / 4) ks = kblocks / 4;
    return ks < 1 ? 1 : ks;
  };
  const int bn_qkv = tile_n(3 * (int)D, 96), bn_13 = tile_n(2 * (int)F, 96), bn_o = tile_n((int)D, 24);
  if (!bn_qkv || !bn_13 || !bn_o) return fail(VAURA_ERR_UNSUPPORTED, "bf16 prefill: d_model %d / ffn %d do not tile", d.d_model, d.ffn_dim);
  __nv_bfloat16* xn = ws.x3;     // [R][D] here
  __nv_bfloat16* act = ws.act3;  // [R][F] here
  for (int l = 0; l < d.num_layers; ++l) {
    LinearTcArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.R = R; g.layer = l; g.d_model = d.d_model; g.kv = kv; g.rope = w.rope;
    g.ksplit = 1; g.pdl = 0;
    CUL(launch_rmsnorm_bf16(ws.h, w.attn_norm + l * D, xn, R, (int)D, D, d.norm_eps, 0, st));
    g.A = xn; g.lda = D; g.K = D; g.W = w.wqkv + (size_t)l * 3 * D * D; g.N = 3 * D; g.epi = EPI_QKV_F32;
    g.out_f32 = ws.q; g.ldo = D; g.block_n = bn_qkv;
    CUL(launch_linear_tc(g, st));
    AttnArgs a{};
    a.q = ws.q; a.out = ws.attn; a.out3 = reinterpret_cast<uint16_t*>(xn); a.out_terms = 1; a.kv = kv; a.state = state;
    a.pos0 = pos0; a.npos = npos; a.layer = l; a.d_model = d.d_model; a.scale = 1.0f / sqrtf((float)kHeadDim);
    CUL(launch_attn(a, d.nhead, R, st));
    g.A = xn; g.lda = D; g.K = D; g.W = w.wo + (size_t)l * D * D; g.N = D; g.epi = EPI_RESID; g.out_f32 = ws.h; g.ldo = D;
    g.block_n = bn_o; g.ksplit = split_k((int)D, bn_o, (int)D / 64);
    CUL(launch_linear_tc(g, st));
    g.ksplit = 1;
    CUL(launch_rmsnorm_bf16(ws.h, w.ffn_norm + l * D, xn, R, (int)D, D, d.norm_eps, 0, st));
    g.A = xn; g.lda = D; g.K = D; g.W = w.w13 + (size_t)l * 2 * F * D; g.N = 2 * F; g.epi = EPI_SWIGLU;
    g.out_bf16 = act; g.ldo = F; g.block_n = bn_13;
    CUL(launch_linear_tc(g, st));
    g.A = act; g.lda = F; g.K = F; g.W = w.w2 + (size_t)l * D * F; g.N = D; g.epi = EPI_RESID; g.out_f32 = ws.h; g.ldo = D;
    g.block_n = bn_o; g.ksplit = split_k((int)D, bn_o, (int)F / 64);
    CUL(launch_linear_tc(g, st));
  }
  // only the last position of every sequence row feeds the heads: a few rows, the weight-streaming GEMV (fp32 activations)
  GemvArgs g{};
  g.state = state; g.pos0 = pos0; g.npos = npos; g.layer = 0; g.d_model = d.d_model; g.eps = d.norm_eps;
  g.W = w.w_heads; g.norm_w = w.final_norm; g.N = d.num_codebooks * d.vocab; g.K = D; g.ldo = g.N; g.out = logits_dst;
  g.x = ws.h + (size_t)(npos - 1) * D; g.ldx = (size_t)npos * D; g.R = rows;
  CUL(launch_gemv(EPI_STORE, true, g, st));
  return VAURA_OK;
}

// Same pass on the tensor-core path: bf16 activations, tcgen05 GEMMs with fused epilogues.
static int transformer_pass_bf16(const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                                 const float* cond_rows, int rows, int npos, int pos0, const StepState* state,
                                 const KvView& kv, float* logits_dst, bool logits_all, cudaStream_t st,
                                 const SampleArgs* fuse_sample = nullptr, bool* sampled = nullptr) {
  const vaura_sampler_dims& d = s->d;
  const vaura_sampler_weights& w = s->w;
  const int R = rows * npos;
  EmbedArgs e{};
  e.seq = seq; e.cond_rows = cond_rows; e.tables = w.tok_tables; e.h = ws.h; e.state = state; e.pos0 = pos0;
  e.npos = npos; e.batch = batch; e.K = d.num_codebooks; e.S = S; e.vocab = d.vocab; e.d_model = d.d_model;
  e.cond_dim = d.cond_dim; e.cond_tokens = d.cond_tokens; e.atpvf = d.audio_tokens_per_video_frame;
  // rows <= 64, one new position per row (graph-replayed decode step): one cooperative kernel for all layers + heads
  // and, when the caller hands in the sampling arguments, for the embedding and the sampling stage as well
  {
    const bool fused_on = knobs().fused_step, fuse_io_on = knobs().fused_io;
    if (fused_on && npos == 1 && !logits_all && state && fused_step_supported(R, d.d_model, d.ffn_dim, d.num_codebooks * d.vocab)) {
      const bool io = fuse_io_on && fuse_sample && sampled && d.cond_dim % 4 == 0 && (d.d_model - d.cond_dim) % 4 == 0;
      // second design (decode_fused2.cu): swap-AB tiles, K split inside CTA pairs, norms folded into their neighbours.
      // Parity-green but measured slower than decode_step_fused_bf16 (profiles/r02_fused2_timeline.summary.txt: 70 vs 40 us
      // per layer at position 127), so it is opt-in: VAURA_FUSED2=1.
      const bool fused2_on = knobs().fused2;
      int sms2 = 0, dev2 = 0;
      cudaGetDevice(&dev2);
      cudaDeviceGetAttribute(&sms2, cudaDevAttrMultiProcessorCount, dev2);
      if (fused2_on && io && ws.f2 &&
          fused2_supported(R, d.num_layers, d.d_model, d.ffn_dim, d.nhead, d.num_codebooks * d.vocab, sms2 & ~1, kv.page_size,
                           kv.max_pages_per_seq)) {
        Fused2Args fa{};
        const size_t Dm = d.d_model;
        char* fp = ws.f2;
        fa.h_t = (float*)fp; fp += Dm * 64 * 4;
        fa.q_t = (__nv_bfloat16*)fp; fp += Dm * 64 * 2;
        fa.ssq_part = (float*)fp; fp += 64 * 64 * 4;
        fa.w2_part = (float*)fp; fp += (Dm / 64) * 3 * 2 * 32 * 64 * 4;
        fa.w2_cnt = (unsigned*)fp;
        fa.attn_norm = w.attn_norm; fa.ffn_norm = w.ffn_norm; fa.final_norm = w.final_norm; fa.rope = w.rope;
        fa.hb = ws.xn_b; fa.attn = ws.attn_b; fa.act = ws.act_b; fa.logits = logits_dst;
        fa.kv = kv; fa.state = const_cast<StepState*>(state);
        fa.R = R; fa.L = d.num_layers; fa.D = d.d_model; fa.F = d.ffn_dim; fa.H = d.nhead; fa.NH = d.num_codebooks * d.vocab;
        fa.eps = d.norm_eps; fa.scale = 1.0f / sqrtf((float)kHeadDim);
        fa.timing = knobs().phase_timing ? ws.timing : nullptr;
        fa.step_times = ws.timing + 1024;
        fa.timing_cta = knobs().timing_cta;
        fa.flags = knobs().fused2_flags;
        fa.seq = seq; fa.cond_rows = cond_rows; fa.tables = w.tok_tables; fa.batch = batch; fa.Kc = d.num_codebooks; fa.S = S;
        fa.vocab = d.vocab; fa.cond_dim = d.cond_dim; fa.cond_tokens = d.cond_tokens; fa.atpvf = d.audio_tokens_per_video_frame;
        fa.sample = *fuse_sample;
        fa.sample.state = nullptr;
        *sampled = true;
        CUL(launch_decode_fused2(fa, w.wqkv, w.wo, w.w13, w.w2, w.w_heads, st));
        return VAURA_OK;
      }
      if (!io) CUL(launch_embed(e, R, st));
      FusedStepArgs fa{};
      fa.attn_norm = w.attn_norm; fa.ffn_norm = w.ffn_norm; fa.final_norm = w.final_norm; fa.rope = w.rope;
      fa.h = ws.h; fa.xn = ws.xn_b; fa.q = ws.q_b; fa.attn = ws.attn_b; fa.act = ws.act_b; fa.logits = logits_dst;
      fa.kv = kv; fa.state = const_cast<StepState*>(state);
      fa.R = R; fa.L = d.num_layers; fa.D = d.d_model; fa.F = d.ffn_dim; fa.H = d.nhead; fa.NH = d.num_codebooks * d.vocab;
      fa.wo_ksplit = kFusedKsplit; fa.w2_ksplit = kFusedKsplit;
      // split-K sums of wo / w2: float reductions into h (default), or - reproducible mode - one partial slice per K split,
      // added in order by the CTA that normalises the row (measured 11 % slower per step: 1095 vs 990 us at 64 rows)
      fa.part = deterministic_mode() ? ws.part : nullptr;
      fa.eps = d.norm_eps; fa.scale = 1.0f / sqrtf((float)kHeadDim);
      fa.timing = knobs().phase_timing ? ws.timing : nullptr;
      fa.step_times = ws.timing + 1024;
      fa.timing_cta = knobs().timing_cta;
      fa.fuse_io = io ? 1 : 0;
      if (io) {
        fa.seq = seq; fa.cond_rows = cond_rows; fa.tables = w.tok_tables; fa.batch = batch; fa.Kc = d.num_codebooks; fa.S = S;
        fa.vocab = d.vocab; fa.cond_dim = d.cond_dim; fa.cond_tokens = d.cond_tokens; fa.atpvf = d.audio_tokens_per_video_frame;
        fa.sample = *fuse_sample;
        fa.sample.state = nullptr;
        *sampled = true;
      }
      CUL(launch_decode_fused_bf16(fa, w.wqkv, w.wo, w.w13, w.w2, w.w_heads, st));
      return VAURA_OK;
    }
  }
  CUL(launch_embed(e, R, st));
  const size_t D = d.d_model, F = d.ffn_dim;
  // narrow N tiles when there is a single M tile so the weight stream is spread over all SMs
  const bool small = R <= 128;
  // Programmatic dependent launch is OFF by default: measured on B200 it gave no step-time gain, and an explicit
  // griddepcontrol.launch_dependents inside the GEMM made dependents observe stale activations.
  // VAURA_PDL_MODE bits: 1 attribute+wait, 2 GEMM trigger after its wait, 4 weight prefetch before the wait,
  // 8 trigger in the small kernels.
  const int pdl = knobs().pdl_mode;
  const bool splitk = small && !knobs().no_splitk;
  // tuning knobs of the two residual GEMMs (N tile, split-K factor)
  const int wo_bn = knobs().wo_bn, wo_ks = knobs().wo_ksplit, w2_bn = knobs().w2_bn, w2_ks = knobs().w2_ksplit;
  for (int l = 0; l < d.num_layers; ++l) {
    LinearTcArgs g{};
    g.state = state; g.pos0 = pos0; g.npos = npos; g.R = R; g.layer = l; g.d_model = d.d_model; g.kv = kv; g.rope = w.rope;
    g.pdl = pdl;
    CUL(launch_rmsnorm_bf16(ws.h, w.attn_norm + l * D, ws.xn_b, R, D, D, d.norm_eps, pdl, st));
    g.A = ws.xn_b; g.lda = D; g.W = w.wqkv + (size_t)l * 3 * D * D; g.N = 3 * D; g.K = D; g.epi = EPI_QKV;
    g.out_bf16 = ws.q_b; g.block_n = small ? 32 : 128;
    CUL(launch_linear_tc(g, st));
    AttnBf16Args a{};
    a.q = ws.q_b; a.out = ws.attn_b; a.kv = kv; a.state = state; a.pos0 = pos0; a.npos = npos; a.layer = l;
    a.d_model = d.d_model; a.scale = 1.0f / sqrtf((float)kHeadDim); a.pdl = pdl;
    CUL(launch_attn_bf16(a, d.nhead, R, st));
    g.A = ws.attn_b; g.lda = D; g.W = w.wo + (size_t)l * D * D; g.N = D; g.K = D; g.epi = EPI_RESID; g.out_f32 = ws.h;
    g.ldo = D; g.block_n = splitk ? wo_bn : (small ? 16 : 128); g.ksplit = splitk ? wo_ks : 1;
    CUL(launch_linear_tc(g, st));
    g.ksplit = 1;
    CUL(launch_rmsnorm_bf16(ws.h, w.ffn_norm + l * D, ws.xn_b, R, D, D, d.norm_eps, pdl, st));
    g.A = ws.xn_b; g.lda = D; g.W = w.w13 + (size_t)l * 2 * F * D; g.N = 2 * F; g.K = D; g.epi = EPI_SWIGLU;
    g.out_bf16 = ws.act_b; g.ldo = F; g.block_n = small ? 64 : 128;
    CUL(launch_linear_tc(g, st));
    g.A = ws.act_b; g.lda = F; g.W = w.w2 + (size_t)l * D * F; g.N = D; g.K = F; g.epi = EPI_RESID; g.out_f32 = ws.h;
    g.ldo = D; g.block_n = splitk ? w2_bn : (small ? 16 : 128); g.ksplit = splitk ? w2_ks : 1;
    CUL(launch_linear_tc(g, st));
    g.ksplit = 1;
  }
  LinearTcArgs g{};
  g.state = state; g.pos0 = pos0; g.npos = npos; g.d_model = d.d_model; g.pdl = pdl;
  g.W = w.w_heads; g.N = d.num_codebooks * d.vocab; g.K = D; g.epi = EPI_STORE; g.out_f32 = logits_dst; g.ldo = g.N;
  g.block_n = small ? 64 : 128;
  if (logits_all) {
    CUL(launch_rmsnorm_bf16(ws.h, w.final_norm, ws.xn_b, R, D, D, d.norm_eps, pdl, st));
    g.A = ws.xn_b; g.lda = D; g.R = R; g.perm_S = npos; g.perm_V = d.vocab;
  } else {  // only the last position of every sequence row feeds the heads
    CUL(launch_rmsnorm_bf16(ws.h + (size_t)(npos - 1) * D, w.final_norm, ws.xn_b, rows, D, (size_t)npos * D, d.norm_eps, pdl, st));
    g.A = ws.xn_b; g.lda = D; g.R = rows;
  }
  CUL(launch_linear_tc(g, st));
  return VAURA_OK;
}

static int run_pass(int precision, const vaura_sampler* s, const Workspace& ws, const int32_t* seq, int batch, int S,
                    const float* cond_rows, int rows, int npos, int pos0, const StepState* state, const KvView& kv,
                    float* logits_dst, bool logits_all, cudaStream_t st, const SampleArgs* fuse_sample = nullptr,
                    bool* sampled = nullptr) {
  return precision == VAURA_PRECISION_BF16
             ? transformer_pass_bf16(s, ws, seq, batch, S, cond_rows, rows, npos, pos0, state, kv, logits_dst, logits_all, st,
                                     fuse_sample, sampled)
             : transformer_pass(s, ws, seq, batch, S, cond_rows, rows, npos, pos0, state, kv, logits_dst, logits_all, st);
}

static int check_kv(const vaura_sampler* s, const vaura_kv_cache* kv, int want_dtype) {
  if (!kv || !kv->pages || !kv->page_table) return fail(VAURA_ERR_INVALID, "kv cache missing");
  {
    // VAURA_ANY_PAGE=1 (experiment): any power-of-two page >= 16 on the bf16 path
    const bool pow2 = kv->page_size >= 16 && !(kv->page_size & (kv->page_size - 1));
    if (kv->page_size != 16 && kv->page_size != 32 && !(knobs().any_page && pow2))
      return fail(VAURA_ERR_INVALID, "page_size must be 16 or 32");
  }
  if (kv->max_pages_per_seq * kv->page_size < s->d.block_size)
    return fail(VAURA_ERR_INVALID, "page table covers %d positions < block_size %d", kv->max_pages_per_seq * kv->page_size,
                s->d.block_size);
  if (kv->dtype != want_dtype) return fail(VAURA_ERR_INVALID, "kv dtype %d does not match precision mode", kv->dtype);
  return VAURA_OK;
}

extern "C" int vaura_sampler_generate(vaura_sampler* s, const vaura_generate_params* p, const vaura_kv_cache* kv,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  refresh_knobs();
  if (!s || !p || !workspace) return fail(VAURA_ERR_INVALID, "null argument");
  const vaura_sampler_dims& d = s->d;
  const int K = d.num_codebooks, S = p->timesteps + K;
  if (p->batch <= 0 || p->timesteps <= 0 || !p->sequence || !p->cond_rows) return fail(VAURA_ERR_INVALID, "bad params");
  if (p->start_offset < 1 || p->start_offset >= S || p->end_offset > S || p->end_offset <= p->start_offset)
    return fail(VAURA_ERR_INVALID, "bad offsets [%d,%d) for S=%d", p->start_offset, p->end_offset, S);
  // sequence positions >= block_size overflow the RoPE table in the reference too (llama.py:493-497)
  if (S - 1 > d.block_size) return fail(VAURA_ERR_INVALID, "sequence of %d columns exceeds block_size %d", S, d.block_size);
  const int rows = p->batch * (p->use_cfg ? 2 : 1);
  const int precision = resolve_precision(p->precision, rows, p->use_sampling && p->temp > 0.f);
  if (precision != VAURA_PRECISION_FP32ACT && precision != VAURA_PRECISION_BF16)
    return fail(VAURA_ERR_INVALID, "unknown precision mode %d", precision);
  int rc = check_kv(s, kv, precision == VAURA_PRECISION_BF16 ? VAURA_KV_BF16 : VAURA_KV_F32);
  if (rc) return rc;
  const int npre = p->start_offset;  // columns [0, start) are consumed by the first pass (prefill when > 1)
  Workspace ws = carve(d, rows, npre, precision, workspace);
  if (ws.bytes > workspace_bytes) return fail(VAURA_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
  // a call that samples has no bit-exactness contract: its prompt prefill takes bf16 operands (transformer_pass_tc1)
  ws.prefill_terms = (p->use_sampling && p->temp > 0.f && knobs().prefill_bf16) ? 1 : 3;
  cudaStream_t st = (cudaStream_t)stream;
  const KvView kvv = kv_view(kv, d.nhead);
  // CUDA events on the caller's stream around the decode-step launches alone (first pass, memsets and host glue outside):
  // what bench.py divides by the step count for the step kernel's average launch duration
  auto loop_mark = [&](int which, int nsteps_) {
    if (!s->loop_ev[which] && cudaEventCreate(&s->loop_ev[which]) != cudaSuccess) { s->loop_ev[which] = nullptr; return; }
    cudaEventRecord(s->loop_ev[which], st);
    s->loop_steps = nsteps_;
  };

  SampleArgs sa{};
  sa.logits = ws.logits; sa.sequence = p->sequence; sa.logits_out = p->logits_out; sa.clip_ids = p->clip_ids;
  sa.B = p->batch; sa.K = K; sa.V = d.vocab; sa.S = S; sa.T = p->timesteps; sa.use_cfg = p->use_cfg;
  sa.use_sampling = p->use_sampling; sa.top_k = p->top_k; sa.cfg_scale = p->cfg_scale; sa.temp = p->temp;
  sa.top_p = p->top_p; sa.seed_lo = (uint32_t)p->seed; sa.seed_hi = (uint32_t)(p->seed >> 32);
  sa.stream_id = p->stream_id;

  const bool persist = precision == VAURA_PRECISION_FP32ACT && persistent_supported(rows, d.d_model, d.ffn_dim, kv->page_size) &&
                       !knobs().no_persistent;
  // rows <= 2: cluster variant (decode_cluster.cu) when the weight streams were packed.  Without a prompt it also runs
  // the first position (a decode step with an empty KV cache), so the whole clip is one kernel per column.
  const bool phase_timing = knobs().phase_timing;
  const bool use_cluster = persist && s->w.wstream && !knobs().no_cluster &&
                           cluster_supported(rows, d.num_layers, d.d_model, d.ffn_dim, d.nhead, K * d.vocab, kv->page_size,
                                             d.cond_dim, S) &&
                           cluster_launchable(rows, phase_timing);
  // Without a prompt the first pass is the decode step of position 0 with an empty K/V cache: the step kernels run it as
  // their first launch (cluster kernel; graph-replayed bf16 step, fused or not) instead of ~170 separate first-pass kernels
  const bool bf16_first = knobs().bf16_step_first;  // 0: keep the separate first pass on the bf16 path
  const bool cluster_first = npre == 1 && (use_cluster || (bf16_first && !persist && precision == VAURA_PRECISION_BF16));
  int nsteps = p->end_offset - (p->start_offset + 1);
  if (!cluster_first) {
    // first pass: positions [0, start) -> sample column start
    rc = run_pass(precision, s, ws, p->sequence, p->batch, S, p->cond_rows, rows, npre, 0, nullptr, kvv, ws.logits, false, st);
    if (rc) return rc;
    sa.state = nullptr; sa.offset = p->start_offset;
    CUL(launch_sample(sa, st));
    if (nsteps <= 0) return VAURA_OK;
    CUL(launch_set_state(ws.state, p->start_offset + 1, st));
  } else {
    CUL(launch_set_state(ws.state, p->start_offset, st));
    nsteps += 1;
  }

  // decode steps
  if (persist) {
    // rows <= 4: one persistent cooperative kernel per step (weights streamed through an smem ring by TMA)
    PersistArgs pa{};
    const vaura_sampler_weights& w = s->w;
    pa.wqkv = w.wqkv; pa.wo = w.wo; pa.w13 = w.w13; pa.w2 = w.w2; pa.w_heads = w.w_heads; pa.attn_norm = w.attn_norm;
    pa.ffn_norm = w.ffn_norm; pa.final_norm = w.final_norm; pa.tok_tables = w.tok_tables; pa.rope = w.rope;
    pa.seq = p->sequence; pa.cond_rows = p->cond_rows; pa.h = ws.h; pa.q = ws.q; pa.act = ws.act; pa.logits = ws.logits;
    pa.attn_part = ws.attn_part; pa.kv = kvv; pa.state = ws.state; pa.sample = sa; pa.sample.state = nullptr;
    pa.L = d.num_layers; pa.D = d.d_model; pa.F = d.ffn_dim; pa.H = d.nhead; pa.Kc = K; pa.V = d.vocab; pa.S = S;
    pa.batch = p->batch; pa.cond_dim = d.cond_dim; pa.cond_tokens = d.cond_tokens; pa.atpvf = d.audio_tokens_per_video_frame;
    pa.eps = d.norm_eps; pa.scale = 1.0f / sqrtf((float)kHeadDim);
    pa.timing = phase_timing ? ws.timing : nullptr;
    pa.step_times = ws.timing + 1024;  // workspace bytes [256 + 8192, ...): one timestamp per generated column
    pa.timing_cta = knobs().timing_cta;
    if (use_cluster) {
      pa.wstream = w.wstream;
      pa.xfix = ws.xfix;
      pa.prefetch_ahead = knobs().cluster_ring;
      pa.pace_cycles = knobs().cluster_l2_ahead;
      pa.tail_units = knobs().cluster_tail_units;
      CU(cudaMemsetAsync(ws.xfix, 0, cluster_xfix_bytes(rows, d.num_layers), st));
    }
    loop_mark(0, nsteps);
    for (int i = 0; i < nsteps; ++i) {
      if (use_cluster) CUL(launch_decode_cluster(pa, rows, st));
      else CUL(launch_decode_persistent(pa, rows, st));
    }
    loop_mark(1, nsteps);
    return VAURA_OK;
  }
  if (precision == VAURA_PRECISION_BF16 && ws.f2)  // decode_step_fused2: arrival counters of the w2 K thirds start at 0
    CU(cudaMemsetAsync(ws.f2 + fused2_workspace_bytes(d.d_model) - 1024, 0, 1024, st));
  // otherwise: capture one step (reads its position from the device state) and replay it; a call with the same pointers,
  // shapes and sampling parameters as the last one replays the graph instantiated then
  GraphKey key;
  memset(&key, 0, sizeof(key));
  key.p = *p; key.kv = *kv; key.workspace = workspace; key.precision = precision; key.knobs = knobs();
  cudaGetDevice(&key.device);
  if (s->graph_exec && s->graph_key == key) {
    loop_mark(0, nsteps);
    for (int i = 0; i < nsteps; ++i) {
      CU(cudaGraphLaunch(s->graph_exec, st));
      g_launches += s->graph_nodes;
    }
    loop_mark(1, nsteps);
    return VAURA_OK;
  }
  cudaGraph_t graph = nullptr;
  cudaStream_t cs = s->capture_stream;
  CU(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  g_capturing = true;
  g_capture_nodes = 0;
  bool sampled = false;  // the fused bf16 step kernel also samples and advances the loop state
  rc = run_pass(precision, s, ws, p->sequence, p->batch, S, p->cond_rows, rows, 1, 0, ws.state, kvv, ws.logits, false, cs, &sa,
                &sampled);
  if (rc == VAURA_OK && !sampled) {
    sa.state = ws.state;
    cudaError_t e = launch_sample(sa, cs);
    if (e != cudaSuccess) rc = fail(VAURA_ERR_CUDA, "launch_sample: %s", cudaGetErrorString(e));
    LAUNCHED(1);
  }
  g_capturing = false;
  cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return fail(VAURA_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  ce = cudaGraphInstantiate(&s->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return fail(VAURA_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
  s->graph_key = key;
  s->graph_nodes = g_capture_nodes;
  loop_mark(0, nsteps);
  for (int i = 0; i < nsteps; ++i) {
    CU(cudaGraphLaunch(s->graph_exec, st));
    g_launches += s->graph_nodes;
  }
  loop_mark(1, nsteps);
  return VAURA_OK;
}

extern "C" int vaura_sampler_last_loop_ms(vaura_sampler* s, float* ms_out, int32_t* steps_out) {
  if (!s || !ms_out || !steps_out) return fail(VAURA_ERR_INVALID, "null argument");
  if (!s->loop_ev[0] || !s->loop_ev[1] || s->loop_steps <= 0) return fail(VAURA_ERR_INVALID, "no generate() call has run its step loop yet");
  CU(cudaEventSynchronize(s->loop_ev[1]));
  CU(cudaEventElapsedTime(ms_out, s->loop_ev[0], s->loop_ev[1]));
  *steps_out = s->loop_steps;
  return VAURA_OK;
}

extern "C" int vaura_sampler_forward(vaura_sampler* s, const int32_t* sequence, const float* cond_rows, int32_t rows,
                                     int32_t S, float* logits_out, const vaura_kv_cache* kv, int32_t precision,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  refresh_knobs();
  if (!s || !sequence || !cond_rows || !logits_out || !workspace || rows <= 0 || S <= 0)
    return fail(VAURA_ERR_INVALID, "bad argument");
  const vaura_sampler_dims& d = s->d;
  if (S > d.block_size) return fail(VAURA_ERR_INVALID, "S=%d exceeds block_size %d (llama.py:493-497)", S, d.block_size);
  precision = resolve_precision(precision, rows, false);
  if (precision != VAURA_PRECISION_FP32ACT && precision != VAURA_PRECISION_BF16)
    return fail(VAURA_ERR_INVALID, "unknown precision mode %d", precision);
  int rc = check_kv(s, kv, precision == VAURA_PRECISION_BF16 ? VAURA_KV_BF16 : VAURA_KV_F32);
  if (rc) return rc;
  Workspace ws = carve(d, rows, S, precision, workspace);
  if (ws.bytes > workspace_bytes) return fail(VAURA_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
  // the heads GEMV stores straight into the reference layout [rows][K][S][V] (llama.py:504 torch.stack(dim=1))
  return run_pass(precision, s, ws, sequence, rows, S, cond_rows, rows, S, 0, nullptr, kv_view(kv, d.nhead), logits_out, true,
                  (cudaStream_t)stream);
}

extern "C" int vaura_sample_logits(const float* logits, int32_t rows, int32_t K, int32_t V, int32_t use_cfg,
                                   float cfg_scale, int32_t use_sampling, float temp, int32_t top_k, float top_p,
                                   uint64_t seed, const int32_t* clip_ids, int32_t offset, int32_t* tokens_out,
                                   float* probs_out, void* stream) {
  refresh_knobs();
  if (!logits || !tokens_out || rows <= 0 || K <= 0) return fail(VAURA_ERR_INVALID, "bad argument");
  if (V != 1024) return fail(VAURA_ERR_UNSUPPORTED, "vocab %d unsupported (sampling kernel is built for 1024)", V);
  SampleArgs sa{};
  sa.logits = logits; sa.tokens_out = tokens_out; sa.probs_out = probs_out; sa.clip_ids = clip_ids; sa.offset = offset;
  sa.B = rows; sa.K = K; sa.V = V; sa.S = 0; sa.T = 0; sa.use_cfg = use_cfg; sa.use_sampling = use_sampling;
  sa.top_k = top_k; sa.cfg_scale = cfg_scale; sa.temp = temp; sa.top_p = top_p; sa.seed_lo = (uint32_t)seed;
  sa.seed_hi = (uint32_t)(seed >> 32);
  sa.stream_id = 0;
  CUL(launch_sample(sa, (cudaStream_t)stream));
  return VAURA_OK;
}

// ---- codec ------------------------------------------------------------------------------------------
// slot order of the weight blob (written by vaura_b200/weights.py: pack_codec)
//   0 code_tables f16 [Kc][Vc][latent]     1 conv_in W f16 [7][C0][latent]     2 conv_in bias f32
//   per block i (base 3 + 21 i): +0 snake alpha f32 [Cin]  +1 convT W f16 [s][2][Cout][Cin]  +2 convT bias
//       per residual unit j (base +3 + 6 j): +0 alpha1  +1 conv7 W f16 [7][C][C]  +2 bias  +3 alpha2
//                                            +4 conv1 W f16 [1][C][C]  +5 bias
//   tail (base 3 + 21 n): +0 final alpha f32 [Cl]  +1 conv_out W f32 [7][Cl]  +2 conv_out bias f32 [1]
//   last slot: tap-offset tables int32 (built by weights.py): [k7 d1][k7 d3][k7 d9][k1][k7 pad3] then per block
//   [s][2] conv-transpose offsets
struct vaura_codec {
  vaura_codec_dims d;
  const char* blob;
  std::vector<int64_t> off;
  std::vector<int> taps;  // host copy of the tap-offset tables (same order as the device slot)
  bool use_tc = true;     // tcgen05 implicit GEMM where the shape is supported, SIMT kernel otherwise
};

static int conv_dispatch(const vaura_codec* c, const ConvArgs& a, int tap_index, int B, cudaStream_t st) {
  if (c->use_tc && conv_tc_supported(a.Cin, a.Cout, a.ntaps, a.nphase)) {
    CUL(launch_conv_tc(a, c->taps.data() + tap_index, B, st));
  } else {
    CUL(launch_conv_gemm(a, B, st));
  }
  return VAURA_OK;
}

extern "C" int vaura_codec_create(const vaura_codec_dims* dims, const vaura_codec_weights* w, vaura_codec** out) {
  refresh_knobs();
  if (!dims || !w || !out || !w->blob || !w->offsets) return fail(VAURA_ERR_INVALID, "null argument");
  const int want = 3 + 21 * dims->n_blocks + 3 + 1;
  if (w->n_offsets != want) return fail(VAURA_ERR_INVALID, "codec blob has %d slots, expected %d", w->n_offsets, want);
  if (dims->n_blocks < 1 || dims->n_blocks > 8) return fail(VAURA_ERR_INVALID, "n_blocks must be 1..8");
  if ((dims->decoder_dim >> dims->n_blocks) % 16 != 0 || dims->latent_dim % 16 != 0)
    return fail(VAURA_ERR_UNSUPPORTED, "channel counts must be multiples of 16");
  if (dims->n_codebooks > 16) return fail(VAURA_ERR_UNSUPPORTED, "n_codebooks > 16");
  vaura_codec* c = new (std::nothrow) vaura_codec();
  if (!c) return fail(VAURA_ERR_INVALID, "out of host memory");
  c->d = *dims;
  c->blob = (const char*)w->blob;
  c->off.assign(w->offsets, w->offsets + w->n_offsets);
  for (int dil : {1, 3, 9})
    for (int j = 0; j < 7; ++j) c->taps.push_back(j * dil - 3 * dil);
  c->taps.push_back(0);
  for (int j = 0; j < 7; ++j) c->taps.push_back(j - 3);
  for (int i = 0; i < dims->n_blocks; ++i) {
    const int s = dims->rates[i], pad = (s + 1) / 2;
    if (s % 2) { delete c; return fail(VAURA_ERR_UNSUPPORTED, "odd upsampling rate %d", s); }
    for (int r = 0; r < s; ++r) {
      c->taps.push_back(0);
      c->taps.push_back(r + pad >= s ? 1 : -1);
    }
  }
  c->use_tc = !knobs().codec_simt;
  *out = c;
  return VAURA_OK;
}

extern "C" void vaura_codec_destroy(vaura_codec* c) { delete c; }

struct CodecWs {
  __half *z, *x, *a0, *a1, *h;
  size_t bytes;
};

static CodecWs codec_carve(const vaura_codec_dims& d, int B, int T, void* base) {
  CodecWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n) {
    char* r = p ? p + off : nullptr;
    off += align_up(n);
    return (__half*)r;
  };
  size_t maxel = (size_t)T * d.decoder_dim, t = T;
  for (int i = 0; i < d.n_blocks; ++i) {
    t *= d.rates[i];
    const size_t el = t * (size_t)(d.decoder_dim >> (i + 1));
    if (el > maxel) maxel = el;
  }
  w.z = take((size_t)B * T * d.latent_dim * 2);
  w.x = take(B * maxel * 2);
  w.a0 = take(B * maxel * 2);
  w.a1 = take(B * maxel * 2);
  w.h = take(B * maxel * 2);
  w.bytes = off;
  return w;
}

extern "C" size_t vaura_codec_workspace_bytes(const vaura_codec* c, int32_t batch, int32_t frames) {
  if (!c || batch <= 0 || frames <= 0) return 0;
  return codec_carve(c->d, batch, frames, nullptr).bytes;
}

extern "C" int vaura_codec_decode(vaura_codec* c, const int32_t* codes, int32_t B, int32_t T, uint16_t* wav_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  refresh_knobs();
  if (!c || !codes || !wav_out || !workspace || B <= 0 || T <= 0) return fail(VAURA_ERR_INVALID, "bad argument");
  const vaura_codec_dims& d = c->d;
  CodecWs ws = codec_carve(d, B, T, workspace);
  if (ws.bytes > workspace_bytes) return fail(VAURA_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  auto H = [&](int slot) { return (const __half*)(c->blob + c->off[slot]); };
  auto F = [&](int slot) { return (const float*)(c->blob + c->off[slot]); };
  const int tail = 3 + 21 * d.n_blocks;
  const int* taps = (const int*)(c->blob + c->off[tail + 3]);
  const int* taps_k7[3] = {taps, taps + 7, taps + 14};
  const int* taps_k1 = taps + 21;
  const int* taps_in = taps + 22;
  const int* taps_ct = taps + 29;

  CUL(launch_from_codes(codes, H(0), ws.z, B, d.n_codebooks, T, d.codebook_size, d.latent_dim, st));
  // conv_in k7 pad 3: z -> (activated with block 0's snake) a0
  ConvArgs a{};
  a.in = ws.z; a.W = H(1); a.tap_off = taps_in; a.bias = F(2); a.alpha = F(3); a.residual = nullptr; a.out_raw = nullptr;
  a.out_act = ws.a0; a.Tin = T; a.Tq = T; a.Tout = T; a.Cin = d.latent_dim; a.Cout = d.decoder_dim; a.ntaps = 7;
  a.nphase = 1; a.ostride = 1;
  int rc = conv_dispatch(c, a, 22, B, st);
  if (rc) return rc;
  int ct_index = 29;
  __half* act_in = ws.a0;
  __half* act_out = ws.a1;
  __half* act_spare = ws.h;  // fused residual units write their activated output next to the one they read (halo rows)
  const bool fused_ru = knobs().codec_fused_ru;  // 0: conv k7 and conv k1 of a residual unit as two launches
  int t = T;
  for (int i = 0; i < d.n_blocks; ++i) {
    const int base = 3 + 21 * i, s = d.rates[i];
    const int cin = d.decoder_dim >> i, cout = d.decoder_dim >> (i + 1);
    // conv-transpose (polyphase): act_in -> x (raw) and act_out = snake(x, alpha of res unit 0)
    ConvArgs ct{};
    ct.in = act_in; ct.W = H(base + 1); ct.tap_off = taps_ct; ct.bias = F(base + 2); ct.alpha = F(base + 3);
    ct.out_raw = ws.x; ct.out_act = act_out; ct.Tin = t; ct.Tq = t; ct.Tout = t * s; ct.Cin = cin; ct.Cout = cout;
    ct.ntaps = 2; ct.nphase = s; ct.ostride = s;
    if ((rc = conv_dispatch(c, ct, ct_index, B, st))) return rc;
    taps_ct += 2 * s;
    ct_index += 2 * s;
    t *= s;
    const bool fuse = fused_ru && c->use_tc && ru_fused_supported(cout);
    for (int j = 0; j < 3; ++j) {
      const int rb = base + 3 + 6 * j;
      // alpha of whatever consumes the block output next: next res unit, next block's snake, or the final snake
      const float* next_alpha = j < 2 ? F(rb + 6) : (i + 1 < d.n_blocks ? F(3 + 21 * (i + 1)) : F(tail));
      if (fuse) {  // one launch per unit, the k7 output stays in shared memory (gemm_ru_fused_kernel)
        RuArgs r{};
        r.act = act_out; r.x = ws.x; r.W7 = H(rb + 1); r.W1 = H(rb + 4); r.bias7 = F(rb + 2); r.alpha2 = F(rb + 3);
        r.bias1 = F(rb + 5); r.alpha_next = next_alpha; r.out_raw = j < 2 ? ws.x : nullptr; r.out_act = act_spare;
        r.T = t; r.C = cout;
        CUL(launch_ru_fused(r, c->taps.data() + 7 * j, B, st));
        __half* tmp2 = act_out; act_out = act_spare; act_spare = tmp2;
        continue;
      }
      ConvArgs c7{};
      c7.in = act_out; c7.W = H(rb + 1); c7.tap_off = taps_k7[j]; c7.bias = F(rb + 2); c7.alpha = F(rb + 3);
      c7.out_act = act_spare; c7.Tin = t; c7.Tq = t; c7.Tout = t; c7.Cin = cout; c7.Cout = cout; c7.ntaps = 7; c7.nphase = 1;
      c7.ostride = 1;
      if ((rc = conv_dispatch(c, c7, 7 * j, B, st))) return rc;
      ConvArgs c1{};
      c1.in = act_spare; c1.W = H(rb + 4); c1.tap_off = taps_k1; c1.bias = F(rb + 5); c1.alpha = next_alpha; c1.residual = ws.x;
      c1.out_raw = j < 2 ? ws.x : nullptr; c1.out_act = act_out; c1.Tin = t; c1.Tq = t; c1.Tout = t; c1.Cin = cout;
      c1.Cout = cout; c1.ntaps = 1; c1.nphase = 1; c1.ostride = 1;
      if ((rc = conv_dispatch(c, c1, 21, B, st))) return rc;
    }
    __half* tmp = act_in; act_in = act_out; act_out = tmp;
  }
  CUL(launch_conv_out_tanh(act_in, F(tail + 1), F(tail + 2), (__half*)wav_out, B, t, d.decoder_dim >> d.n_blocks, st));
  return VAURA_OK;
}

// ---- codec encode --------------------------------------------------------------------------------------
struct vaura_codec_encoder {
  vaura_codec_dims d;
  int enc_dim, cb_dim;
  const char* blob;
  std::vector<int64_t> off;
  std::vector<int> taps;  // host copy: [k7 d1][k7 d3][k7 d9][k1][k3]
  bool use_tc = true;
};

extern "C" int vaura_codec_encoder_create(const vaura_codec_dims* dims, int32_t encoder_dim, int32_t codebook_dim,
                                          const vaura_codec_weights* w, vaura_codec_encoder** out) {
  refresh_knobs();
  if (!dims || !w || !out || !w->blob || !w->offsets) return fail(VAURA_ERR_INVALID, "null argument");
  if (dims->n_blocks < 1 || dims->n_blocks > 8) return fail(VAURA_ERR_INVALID, "n_blocks must be 1..8");
  const int want = 2 + 21 * dims->n_blocks + 8;
  if (w->n_offsets != want) return fail(VAURA_ERR_INVALID, "codec encoder blob has %d slots, expected %d", w->n_offsets, want);
  if (encoder_dim % 16 || encoder_dim < 16) return fail(VAURA_ERR_UNSUPPORTED, "encoder_dim must be a multiple of 16");
  if (codebook_dim < 1 || codebook_dim > 32) return fail(VAURA_ERR_UNSUPPORTED, "codebook_dim must be 1..32");
  if (dims->latent_dim > 8192 || dims->n_codebooks > 32) return fail(VAURA_ERR_UNSUPPORTED, "latent_dim / n_codebooks too large");
  vaura_codec_encoder* c = new (std::nothrow) vaura_codec_encoder();
  if (!c) return fail(VAURA_ERR_INVALID, "out of host memory");
  c->d = *dims; c->enc_dim = encoder_dim; c->cb_dim = codebook_dim;
  c->blob = (const char*)w->blob;
  c->off.assign(w->offsets, w->offsets + w->n_offsets);
  for (int dil : {1, 3, 9})
    for (int j = 0; j < 7; ++j) c->taps.push_back(j * dil - 3 * dil);
  c->taps.push_back(0);
  for (int j = 0; j < 3; ++j) c->taps.push_back(j - 1);
  c->use_tc = !knobs().codec_simt;
  *out = c;
  return VAURA_OK;
}

extern "C" void vaura_codec_encoder_destroy(vaura_codec_encoder* c) { delete c; }

struct EncWs {
  float* wav_unused;
  __half *x, *x2, *a0, *a1, *h, *z;
  size_t bytes;
};

static EncWs enc_carve(const vaura_codec_encoder* c, int B, int L, void* base) {
  EncWs w{};
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n) {
    char* r = p ? p + off : nullptr;
    off += align_up(n);
    return (__half*)r;
  };
  // activation size per clip: L * C0 in the first block; a block with stride s turns (T, C) into (T / s, 2 C), never larger
  size_t maxel = (size_t)L * c->enc_dim, t = L, ch = c->enc_dim;
  for (int i = 0; i < c->d.n_blocks; ++i) {
    const int s = c->d.rates[c->d.n_blocks - 1 - i];
    t /= s; ch *= 2;
    if (t * ch > maxel) maxel = t * ch;
  }
  w.x = take(B * maxel * 2); w.x2 = take(B * maxel * 2); w.a0 = take(B * maxel * 2); w.a1 = take(B * maxel * 2);
  w.h = take(B * maxel * 2);
  w.z = take((size_t)B * t * c->d.latent_dim * 2);
  w.bytes = off;
  return w;
}

extern "C" size_t vaura_codec_encoder_workspace_bytes(const vaura_codec_encoder* c, int32_t batch, int32_t samples) {
  if (!c || batch <= 0 || samples <= 0) return 0;
  return enc_carve(c, batch, samples, nullptr).bytes;
}

extern "C" int vaura_codec_encode(vaura_codec_encoder* c, const float* wav, int32_t B, int32_t L, int32_t* codes_out,
                                  uint16_t* latent_out, void* workspace, size_t workspace_bytes, void* stream) {
  refresh_knobs();
  if (!c || !wav || !codes_out || !workspace || B <= 0 || L <= 0) return fail(VAURA_ERR_INVALID, "bad argument");
  const vaura_codec_dims& d = c->d;
  int hop = 1;
  for (int i = 0; i < d.n_blocks; ++i) hop *= d.rates[i];
  if (L % hop) return fail(VAURA_ERR_INVALID, "samples %d is not a multiple of the hop length %d (pad first: DAC.preprocess)", L, hop);
  EncWs ws = enc_carve(c, B, L, workspace);
  if (ws.bytes > workspace_bytes) return fail(VAURA_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  auto H = [&](int slot) { return (const __half*)(c->blob + c->off[slot]); };
  auto F = [&](int slot) { return (const float*)(c->blob + c->off[slot]); };
  const int tail = 2 + 21 * d.n_blocks;
  const int* taps = (const int*)(c->blob + c->off[tail + 7]);
  const int* taps_k7[3] = {taps, taps + 7, taps + 14};
  const int* taps_k1 = taps + 21;
  const int* taps_k3 = taps + 22;
  auto conv = [&](const ConvArgs& a, int tap_index) -> int {
    if (c->use_tc && conv_tc_supported(a.Cin, a.Cout, a.ntaps, a.nphase)) {
      CUL(launch_conv_tc(a, c->taps.data() + tap_index, B, st));
    } else {
      CUL(launch_conv_gemm(a, B, st));
    }
    return VAURA_OK;
  };
  int rc;
  int t = L, ch = c->enc_dim;
  __half *x = ws.x, *x2 = ws.x2, *act = ws.a0, *act2 = ws.a1, *spare = ws.h;
  const bool fused_ru = knobs().codec_fused_ru;  // 0: conv k7 and conv k1 of a residual unit as two launches
  CUL(launch_enc_conv_in(wav, F(0), F(1), F(2), x, act, B, L, ch, st));
  for (int i = 0; i < d.n_blocks; ++i) {
    const int base = 2 + 21 * i, s = d.rates[d.n_blocks - 1 - i];  // encoder strides = reversed decoder rates
    const bool fuse = fused_ru && c->use_tc && ru_fused_supported(ch);
    for (int j = 0; j < 3; ++j) {
      const int rb = base + 6 * j;
      if (fuse) {  // one launch per unit (gemm_ru_fused_kernel); the activated output goes to the spare buffer (halo rows)
        RuArgs r{};
        r.act = act; r.x = x; r.W7 = H(rb + 1); r.W1 = H(rb + 4); r.bias7 = F(rb + 2); r.alpha2 = F(rb + 3); r.bias1 = F(rb + 5);
        r.alpha_next = j < 2 ? F(rb + 6) : F(base + 18); r.out_raw = j < 2 ? x : nullptr; r.out_act = spare; r.T = t; r.C = ch;
        CUL(launch_ru_fused(r, c->taps.data() + 7 * j, B, st));
        __half* tmp2 = act; act = spare; spare = tmp2;
        continue;
      }
      ConvArgs c7{};
      c7.in = act; c7.W = H(rb + 1); c7.tap_off = taps_k7[j]; c7.bias = F(rb + 2); c7.alpha = F(rb + 3); c7.out_act = spare;
      c7.Tin = t; c7.Tq = t; c7.Tout = t; c7.Cin = ch; c7.Cout = ch; c7.ntaps = 7; c7.nphase = 1; c7.ostride = 1;
      if ((rc = conv(c7, 7 * j))) return rc;
      ConvArgs c1{};
      c1.in = spare; c1.W = H(rb + 4); c1.tap_off = taps_k1; c1.bias = F(rb + 5);
      c1.alpha = j < 2 ? F(rb + 6) : F(base + 18);  // next residual unit's first Snake, or the block's Snake before the stride
      c1.residual = x; c1.out_raw = j < 2 ? x : nullptr; c1.out_act = act;
      c1.Tin = t; c1.Tq = t; c1.Tout = t; c1.Cin = ch; c1.Cout = ch; c1.ntaps = 1; c1.nphase = 1; c1.ostride = 1;
      if ((rc = conv(c1, 21))) return rc;
    }
    // WNConv1d(C -> 2C, k = 2 s, stride s, pad ceil(s / 2)) as a three-tap convolution over frames of s samples
    if (t % s) return fail(VAURA_ERR_INVALID, "length %d is not a multiple of stride %d", t, s);
    ConvArgs cs{};
    cs.in = act; cs.W = H(base + 19); cs.tap_off = taps_k3; cs.bias = F(base + 20);
    cs.alpha = i + 1 < d.n_blocks ? F(2 + 21 * (i + 1)) : F(tail);  // next block's first Snake, or the final Snake
    cs.out_raw = i + 1 < d.n_blocks ? x2 : nullptr; cs.out_act = act2;
    cs.Tin = t / s; cs.Tq = t / s; cs.Tout = t / s; cs.Cin = s * ch; cs.Cout = 2 * ch; cs.ntaps = 3; cs.nphase = 1; cs.ostride = 1;
    if ((rc = conv(cs, 22))) return rc;
    t /= s; ch *= 2;
    __half* tmp = x; x = x2; x2 = tmp;
    tmp = act; act = act2; act2 = tmp;
  }
  __half* zdst = latent_out ? (__half*)latent_out : ws.z;
  ConvArgs cz{};
  cz.in = act; cz.W = H(tail + 1); cz.tap_off = taps_k3; cz.bias = F(tail + 2); cz.alpha = nullptr; cz.out_raw = zdst;
  cz.Tin = t; cz.Tq = t; cz.Tout = t; cz.Cin = ch; cz.Cout = d.latent_dim; cz.ntaps = 3; cz.nphase = 1; cz.ostride = 1;
  if ((rc = conv(cz, 22))) return rc;
  CUL(launch_rvq_encode(zdst, F(tail + 3), F(tail + 4), F(tail + 5), F(tail + 6), codes_out, B, d.n_codebooks, t,
                        d.codebook_size, d.latent_dim, c->cb_dim, st));
  return VAURA_OK;
}

// ---- Segment-AVCLIP visual tower (avclip.cu) ------------------------------------------------------------
struct vaura_avclip {
  vaura_avclip_dims d;
  const char* blob;
  std::vector<int64_t> off;
};

struct AvclipWs {
  float* x;
  __nv_bfloat16 *xn, *qkv, *att, *hid, *acls, *xn2, *hid2;
  size_t bytes;
};

static AvclipWs avclip_carve(const vaura_avclip_dims& d, int S, void* base) {
  AvclipWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n) {
    char* r = p ? p + off : nullptr;
    off += align_up(n);
    return r;
  };
  const size_t D = d.embed_dim, t = d.frames / d.tubelet, g = d.img_size / d.patch_size, n = g * g, T = 1 + t * n;
  const size_t rows = (size_t)S * t * (n + 1);  // >= S * T: the aggregation sequences have one CLS row per frame
  const size_t Kp = (size_t)d.in_chans * d.tubelet * d.patch_size * d.patch_size;
  size_t hid = (size_t)S * T * d.mlp_ratio * D;
  if ((size_t)S * t * n * Kp > hid) hid = (size_t)S * t * n * Kp;
  w.x = (float*)take((size_t)S * T * D * 4);
  w.xn = (__nv_bfloat16*)take(rows * D * 2);
  w.qkv = (__nv_bfloat16*)take(rows * 3 * D * 2);
  w.att = (__nv_bfloat16*)take((size_t)S * T * D * 2);
  w.hid = (__nv_bfloat16*)take(hid * 2);
  w.acls = (__nv_bfloat16*)take((size_t)S * t * D * 2);
  w.xn2 = (__nv_bfloat16*)take((size_t)S * t * D * 2);
  w.hid2 = (__nv_bfloat16*)take((size_t)S * t * d.mlp_ratio * D * 2);
  w.bytes = off;
  return w;
}

extern "C" int vaura_avclip_create(const vaura_avclip_dims* dims, const vaura_avclip_weights* w, vaura_avclip** out) {
  refresh_knobs();
  if (!dims || !w || !out || !w->blob || !w->offsets) return fail(VAURA_ERR_INVALID, "null argument");
  const vaura_avclip_dims& d = *dims;
  if (d.depth < 1 || d.num_heads < 1 || d.embed_dim != d.num_heads * 64)
    return fail(VAURA_ERR_UNSUPPORTED, "head width %d unsupported (kernels are built for 64)", d.num_heads ? d.embed_dim / d.num_heads : 0);
  if (d.embed_dim % 128 || d.embed_dim > 1024 || (d.embed_dim != 256 && d.embed_dim != 512 && d.embed_dim != 768 && d.embed_dim != 1024))
    return fail(VAURA_ERR_UNSUPPORTED, "embed_dim %d unsupported (256, 512, 768, 1024)", d.embed_dim);
  if (d.tubelet < 1 || d.frames % d.tubelet || d.frames / d.tubelet != 8)
    return fail(VAURA_ERR_UNSUPPORTED, "frames / tubelet must be 8 (TEMPORAL_RESOLUTION of divided_224_16x4)");
  if (d.patch_size % 8 || d.img_size % d.patch_size || (d.in_chans * d.tubelet * d.patch_size * d.patch_size) % 64)
    return fail(VAURA_ERR_UNSUPPORTED, "patch geometry unsupported");
  const int g = d.img_size / d.patch_size;
  if (g * g + 1 > 1024) return fail(VAURA_ERR_UNSUPPORTED, "more than 1023 patches per frame");
  if (d.mlp_ratio < 1) return fail(VAURA_ERR_INVALID, "mlp_ratio");
  const int want = 4 + 18 * d.depth + 15;
  if (w->n_offsets != want) return fail(VAURA_ERR_INVALID, "avclip blob has %d slots, expected %d", w->n_offsets, want);
  vaura_avclip* a = new (std::nothrow) vaura_avclip();
  if (!a) return fail(VAURA_ERR_INVALID, "out of host memory");
  a->d = d;
  a->blob = (const char*)w->blob;
  a->off.assign(w->offsets, w->offsets + w->n_offsets);
  *out = a;
  return VAURA_OK;
}

extern "C" void vaura_avclip_destroy(vaura_avclip* a) { delete a; }

extern "C" size_t vaura_avclip_workspace_bytes(const vaura_avclip* a, int32_t segments) {
  if (!a || segments <= 0) return 0;
  return avclip_carve(a->d, segments, nullptr).bytes;
}

static int avclip_chunk(const vaura_avclip* a, const float* frames, int S, float* feats, const AvclipWs& ws, cudaStream_t st) {
  const vaura_avclip_dims& d = a->d;
  const int D = d.embed_dim, H = d.num_heads, t = d.frames / d.tubelet, g = d.img_size / d.patch_size, n = g * g, T = 1 + t * n;
  const int Kp = d.in_chans * d.tubelet * d.patch_size * d.patch_size, F = d.mlp_ratio * D;
  const float eps = 1e-6f;  // partial(nn.LayerNorm, eps=1e-6) (video_model_builder.py:41), layer_norm_eps=1e-6 (motionformer.py:176)
  auto W = [&](int slot) { return (const void*)(a->blob + a->off[slot]); };
  auto V = [&](int slot) { return (const float*)(a->blob + a->off[slot]); };
  auto linear = [&](const void* A, int lda, int M, int K, int wslot, int N, int mode, int gelu, void* out_bf16, float* out_f32,
                    int ldo) -> int {
    VitLinearArgs l{};
    l.A = A; l.lda = lda; l.M = M; l.K = K; l.W = W(wslot); l.bias = V(wslot + 1); l.N = N; l.mode = mode; l.gelu = gelu;
    l.out_bf16 = out_bf16; l.out_f32 = out_f32; l.ldo = ldo;
    CUL(launch_vit_linear(l, st));
    return VAURA_OK;
  };
  int rc;
  // tubelet embedding + position embeddings (video_model_builder.py:185, :213-245)
  CUL(launch_vit_patchify(frames, ws.hid, S, d.in_chans, d.frames, d.img_size, d.img_size, d.tubelet, d.patch_size, st));
  {
    VitLinearArgs l{};
    l.A = ws.hid; l.lda = Kp; l.M = S * t * n; l.K = Kp; l.W = W(0); l.bias = V(1); l.N = D; l.mode = VIT_PATCH;
    l.out_f32 = ws.x; l.ldo = D; l.pos = V(2); l.rows_in = t * n; l.rows_out = T; l.row_off = 1;
    CUL(launch_vit_linear(l, st));
  }
  CUL(launch_vit_broadcast_row(ws.x, V(3), D, S, (size_t)T, st));
  for (int i = 0; i < d.depth; ++i) {  // DividedSpaceTimeBlock.forward (vit_helper.py:443-472)
    const int b = 4 + 18 * i;
    CUL(launch_vit_layernorm(ws.x, V(b), V(b + 1), ws.xn, S * T, D, eps, st));
    if ((rc = linear(ws.xn, D, S * T, D, b + 2, 3 * D, VIT_STORE_BF16, 0, ws.qkv, nullptr, 3 * D))) return rc;
    CUL(launch_vit_time_attn(ws.qkv, ws.att, S, t, n, H, st));
    CUL(launch_vit_cls_attn(ws.qkv, ws.att, S, T, H, T, st));
    if ((rc = linear(ws.att, D, S * T, D, b + 4, D, VIT_RESID_F32, 0, nullptr, ws.x, D))) return rc;
    CUL(launch_vit_layernorm(ws.x, V(b + 6), V(b + 7), ws.xn, S * T, D, eps, st));
    if ((rc = linear(ws.xn, D, S * T, D, b + 8, 3 * D, VIT_STORE_BF16, 0, ws.qkv, nullptr, 3 * D))) return rc;
    CUL(launch_vit_space_attn(ws.qkv, ws.att, S, t, n, H, st));
    CUL(launch_vit_cls_attn(ws.qkv, ws.att, S, T, H, T, st));
    if ((rc = linear(ws.att, D, S * T, D, b + 10, D, VIT_RESID_F32, 0, nullptr, ws.x, D))) return rc;
    CUL(launch_vit_layernorm(ws.x, V(b + 12), V(b + 13), ws.xn, S * T, D, eps, st));
    if ((rc = linear(ws.xn, D, S * T, D, b + 14, F, VIT_STORE_BF16, 1, ws.hid, nullptr, F))) return rc;
    if ((rc = linear(ws.hid, F, S * T, F, b + 16, D, VIT_RESID_F32, 0, nullptr, ws.x, D))) return rc;
  }
  // feature head (motionformer.py:309-342): final norm on the patch tokens, one aggregation sequence per frame
  const int tb = 4 + 18 * d.depth;
  CUL(launch_vit_final_norm_agg(ws.x, V(tb + 2), V(tb), V(tb + 1), V(tb + 3), V(tb + 4), ws.xn, S, t, n, D, eps, st));
  if ((rc = linear(ws.xn, D, S * t * (n + 1), D, tb + 5, 3 * D, VIT_STORE_BF16, 0, ws.qkv, nullptr, 3 * D))) return rc;
  CUL(launch_vit_cls_attn(ws.qkv, ws.acls, S * t, n + 1, H, 1, st));
  CUL(launch_vit_broadcast_row(feats, V(tb + 2), D, S * t, 1, st));  // residual of the CLS row = the CLS token itself
  if ((rc = linear(ws.acls, D, S * t, D, tb + 7, D, VIT_RESID_F32, 0, nullptr, feats, D))) return rc;
  CUL(launch_vit_layernorm(feats, V(tb + 9), V(tb + 10), ws.xn2, S * t, D, eps, st));
  if ((rc = linear(ws.xn2, D, S * t, D, tb + 11, F, VIT_STORE_BF16, 1, ws.hid2, nullptr, F))) return rc;
  if ((rc = linear(ws.hid2, F, S * t, F, tb + 13, D, VIT_RESID_F32, 0, nullptr, feats, D))) return rc;
  return VAURA_OK;
}

extern "C" int vaura_avclip_forward(vaura_avclip* a, const float* frames, int32_t segments, float* features_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  refresh_knobs();
  if (!a || !frames || !features_out || !workspace || segments <= 0) return fail(VAURA_ERR_INVALID, "bad argument");
  const vaura_avclip_dims& d = a->d;
  int chunk = segments;
  while (chunk > 1 && avclip_carve(d, chunk, nullptr).bytes > workspace_bytes) chunk = (chunk + 1) / 2;
  if (avclip_carve(d, chunk, nullptr).bytes > workspace_bytes)
    return fail(VAURA_ERR_WORKSPACE, "workspace %zu < %zu bytes (one segment)", workspace_bytes, avclip_carve(d, 1, nullptr).bytes);
  const AvclipWs ws = avclip_carve(d, chunk, workspace);
  const size_t seg_in = (size_t)d.in_chans * d.frames * d.img_size * d.img_size;
  const size_t seg_out = (size_t)(d.frames / d.tubelet) * d.embed_dim;
  for (int s0 = 0; s0 < segments; s0 += chunk) {
    const int ns = segments - s0 < chunk ? segments - s0 : chunk;
    int rc = avclip_chunk(a, frames + (size_t)s0 * seg_in, ns, features_out + (size_t)s0 * seg_out, ws, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return VAURA_OK;
}
