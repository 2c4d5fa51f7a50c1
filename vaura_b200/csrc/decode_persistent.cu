// Persistent decode-step kernel for the HBM-bound small-batch regime (rows <= 4, fp32 activations).
//
// One CTA per SM runs the WHOLE decode step: embedding -> 24 x {RMSNorm+QKV+RoPE+KV append | paged attention |
// wo+residual | RMSNorm+w1|w3+SiLU*mul | w2+residual} -> final norm + 9 heads -> CFG/sampling/write-back.
// Every CTA owns a fixed contiguous block of output rows of every weight matrix, so its weight bytes for the
// step are one fixed sequence of contiguous chunks, streamed with cp.async.bulk (TMA 1-D) into two ~96 KB
// shared-memory slots (double buffer) whose arrival is tracked by one mbarrier each.  A slot holds one "group" of up
// to 16 work items (one row pair per warp; pairs with K > 2048 are split along K over two warps).  All 16 warps
// compute; when a group is finished (CTA-wide named barrier) thread 0 immediately issues the copy of the group after
// next into the slot that just became free, so the copy of group g+1 is always in flight while group g is consumed
// and during grid barriers / attention / staging.  The warps read weights and activations from shared memory only;
// phases are separated by a device-wide barrier (red.release + ld.acquire spin per CTA).
//
// Replaces the same reference lines as decode_fp32.cu (llama.py:445-517 for one position) plus the sampling
// stage of sampling.cu; arithmetic (fp32 accumulate order per row aside) is identical to gemv_kernel/attn_kernel.
#include <cstdlib>

#include "sampling.cuh"

namespace vaura {

namespace {

constexpr int NW = 16;                     // warps per CTA (512 threads -> 128 registers per thread); all of them compute,
                                           // thread 0 also issues the weight copies
constexpr int kConsumers = NW * 32;        // 512
constexpr int kThreadsP = kConsumers;
constexpr int kMaxSplit = 4;               // a row pair with K > 1536 is split over up to 4 warps along K
constexpr int ATT_STRIDE = 100;            // floats per attention partial: m, l, pad, pad, o[96]

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mb_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kSpinTimeoutNs = 2000000000ull;  // 2 s: a protocol bug traps (-> CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
  if (mb_try_wait(bar, parity)) return;
  const uint64_t t0 = now_ns();
  while (!mb_try_wait(bar, parity))
    if (now_ns() - t0 > kSpinTimeoutNs) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// device-wide barrier among the consumer threads of all CTAs (all CTAs are co-resident: cooperative launch).
// arrive = red.release (cumulative over the CTA's writes ordered by bar.sync), wait = ld.acquire spin.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  consumer_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    if ((int)(ld_acquire(counter) - target) < 0) {
      const uint64_t t0 = now_ns();
      while ((int)(ld_acquire(counter) - target) < 0)
        if (now_ns() - t0 > kSpinTimeoutNs) __trap();
    }
  }
  consumer_sync();
}

// explicit shared-memory loads: through generic pointers the compiler emits LD.E (generic), whose latency to the
// shared window is several times that of LDS and sat on the critical path of every row pair
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// dot products of one row pair (rows at shared addresses w0, w1; K bf16 each) over chunks [c0, c1) with NB
// activation rows (xs_addr: permuted fp32 rows in shared memory).  Loads of a 3-chunk window are issued before its FMAs.
template <int NB>
__device__ __forceinline__ void pair_dot(uint32_t w0, uint32_t w1, uint32_t xs_addr, int K, int c0, int c1, int lane,
                                         float (&acc0)[NB], float (&acc1)[NB]) {
#pragma unroll
  for (int r = 0; r < NB; ++r) acc0[r] = acc1[r] = 0.f;
  for (int base = c0 + lane; base < c1; base += 96) {
    uint4 wa[3], wb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = base + 32 * i;
      if (c < c1) { wa[i] = lds_u4(w0 + 16 * c); wb[i] = lds_u4(w1 + 16 * c); }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      float4 xl[3], xh[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = base + 32 * i;
        if (c < c1) {
          xl[i] = lds_f4(xs_addr + 4 * (r * K + 4 * c));
          xh[i] = lds_f4(xs_addr + 4 * (r * K + (K >> 1) + 4 * c));
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = base + 32 * i;
        if (c < c1) {
          float p0 = bf16_lo(wa[i].x) * xl[i].x, p1 = bf16_lo(wb[i].x) * xl[i].x;
          p0 = fmaf(bf16_hi(wa[i].x), xl[i].y, p0); p1 = fmaf(bf16_hi(wb[i].x), xl[i].y, p1);
          p0 = fmaf(bf16_lo(wa[i].y), xl[i].z, p0); p1 = fmaf(bf16_lo(wb[i].y), xl[i].z, p1);
          p0 = fmaf(bf16_hi(wa[i].y), xl[i].w, p0); p1 = fmaf(bf16_hi(wb[i].y), xl[i].w, p1);
          p0 = fmaf(bf16_lo(wa[i].z), xh[i].x, p0); p1 = fmaf(bf16_lo(wb[i].z), xh[i].x, p1);
          p0 = fmaf(bf16_hi(wa[i].z), xh[i].y, p0); p1 = fmaf(bf16_hi(wb[i].z), xh[i].y, p1);
          p0 = fmaf(bf16_lo(wa[i].w), xh[i].z, p0); p1 = fmaf(bf16_lo(wb[i].w), xh[i].z, p1);
          p0 = fmaf(bf16_hi(wa[i].w), xh[i].w, p0); p1 = fmaf(bf16_hi(wb[i].w), xh[i].w, p1);
          acc0[r] += p0; acc1[r] += p1;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) { acc0[r] = warp_sum(acc0[r]); acc1[r] = warp_sum(acc1[r]); }
}

__device__ __forceinline__ void pair_range(int pairs, int cta, int G, int& p0, int& p1) {
  p0 = (int)(((unsigned)pairs * (unsigned)cta) / (unsigned)G);  // pairs * G < 2^31 for every supported shape
  p1 = (int)(((unsigned)pairs * (unsigned)(cta + 1)) / (unsigned)G);
}

// Grouping of a CTA's row pairs of one phase: `splits` warps share a pair along K, a group holds at most
// NW / splits pairs and at most slot_cap bytes, and the pairs are spread evenly over the groups.
struct GroupPlan {
  int splits, ngroups, base, rem;  // group g holds base + (g < rem) pairs
  __device__ GroupPlan(int K, int npairs, int slot_cap) {
    splits = (K + 2047) / 2048;
    if (splits > kMaxSplit) splits = kMaxSplit;
    int ppg = NW / splits;
    const int cap = slot_cap / (4 * K);
    if (ppg > cap) ppg = cap;
    ngroups = (npairs + ppg - 1) / ppg;
    base = ngroups ? npairs / ngroups : 0;
    rem = ngroups ? npairs % ngroups : 0;
  }
};

// The CTA's weight groups of one decode step in consumption order.
struct WeightSched {
  int cta, G, l, ph, g, pp, open;
  int gp_ngroups, gp_base, gp_rem;
  const uint16_t* W;
  int K;
  __device__ void init(int cta_, int G_) {
    cta = cta_; G = G_; l = 0; ph = -1; g = 0; pp = 0; open = 0; gp_ngroups = 0; gp_base = 0; gp_rem = 0; W = nullptr; K = 0;
  }
  __device__ bool advance_phase(const PersistArgs& a) {
    ++ph;
    if (l < a.L && ph == 4) { ph = 0; ++l; }
    if (l > a.L || (l == a.L && ph > 0)) return false;
    const int D = a.D, F = a.F;
    int pairs;
    if (l == a.L) { W = a.w_heads; K = D; pairs = a.Kc * a.V / 2; }
    else if (ph == 0) { W = a.wqkv + (size_t)l * 3 * D * D; K = D; pairs = 3 * D / 2; }
    else if (ph == 1) { W = a.wo + (size_t)l * D * D; K = D; pairs = D / 2; }
    else if (ph == 2) { W = a.w13 + (size_t)l * 2 * F * D; K = D; pairs = F; }
    else { W = a.w2 + (size_t)l * D * F; K = F; pairs = D / 2; }
    int p0, p1;
    pair_range(pairs, cta, G, p0, p1);
    const GroupPlan gp(K, p1 - p0, a.slot_cap);
    gp_ngroups = gp.ngroups; gp_base = gp.base; gp_rem = gp.rem;
    g = 0;
    pp = p0;
    return true;
  }
  __device__ bool next(const PersistArgs& a, const uint16_t*& ptr, uint32_t& bytes) {
    while (!open || g >= gp_ngroups) {
      if (!advance_phase(a)) return false;
      open = 1;
    }
    const int n = gp_base + (g < gp_rem ? 1 : 0);
    ptr = W + (size_t)pp * 2 * K;
    bytes = (uint32_t)n * 4u * (uint32_t)K;
    pp += n;
    ++g;
    return true;
  }
};

constexpr int HOWN = 32;       // floats per row of this CTA's own slice of the residual stream
constexpr int ATT_SCR = 4096;  // floats of attention scratch

__device__ __forceinline__ int perm_idx(int k, int K) { return ((k & 4) ? (K >> 1) : 0) + ((k >> 3) << 2) + (k & 3); }

struct Phase {
  const uint16_t* W;  // [2*pairs][K] bf16
  int K, pairs, epi;
};


}  // namespace


// CTA-wide context at the start of dynamic shared memory.  The phase functions are __noinline__ (the kernel's code
// must stay inside the instruction cache) and read everything they need from here with LDS: with 227 KB of the SM
// carved out as shared memory there is almost no L1 left, so a by-reference kernel-parameter struct or a local-
// memory context would turn every field access into an L2 round trip.
struct SmemCtx {
  PersistArgs a;
  WeightSched sched;   // load cursor, touched by thread 0 only
  unsigned issued;     // groups whose copy has been issued
  int slot_cap, slots_off, xs_off, scr_off, hown_off, rope_off, red_off, part_off, page_off, bar_off;
  int cta, G, p, own0;
};
constexpr int kCtxBytes = 1024;  // PersistArgs + schedule cursor + layout offsets
static_assert(sizeof(SmemCtx) <= kCtxBytes, "SmemCtx must fit its reserved shared-memory block");

#define VAURA_SMEM_VIEW()                                                                     \
  extern __shared__ __align__(128) uint8_t smem[];                                            \
  const SmemCtx& sc = *reinterpret_cast<const SmemCtx*>(smem);                                \
  const PersistArgs& a = sc.a;                                                                \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                              \
  float* xs = reinterpret_cast<float*>(smem + sc.xs_off);                                     \
  float* red = reinterpret_cast<float*>(smem + sc.red_off);                                   \
  float* part_s = reinterpret_cast<float*>(smem + sc.part_off);                               \
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + sc.bar_off);                            \
  (void)a; (void)tid; (void)warp; (void)lane; (void)xs; (void)red; (void)part_s; (void)full;

// thread 0: issue the bulk copy of the next group of the schedule into slot (issued & 1)
__device__ __forceinline__ void issue_next_group() {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemCtx& w = *reinterpret_cast<SmemCtx*>(smem);
  const uint16_t* ptr;
  uint32_t bytes;
  if (!w.sched.next(w.a, ptr, bytes)) return;
  const int slot = w.issued & 1;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + w.bar_off);
  mb_expect_tx(&full[slot], bytes);
  bulk_g2s(smem + w.slots_off + (size_t)slot * w.slot_cap, ptr, bytes, &full[slot]);
  ++w.issued;
}

// xs[r] = (x * rsqrt(mean(x^2)+eps)) * w, permuted.  Source: global h (through L2) or, for layer 0, the embedding
// rows sitting unpermuted in xs (read into registers before the permuted overwrite; syncs in between).
template <int NB>
__device__ __noinline__ void stage_norm(const float* src_smem, const float* src_global, const float* w, int K) {
  VAURA_SMEM_VIEW();
  const float eps = a.eps;
  for (int r = 0; r < NB; ++r) {
    float ss = 0.f;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = tid < (K >> 2);
    if (live) {
      v = src_smem ? *reinterpret_cast<const float4*>(src_smem + (size_t)r * K + 4 * tid)
                   : __ldcg(reinterpret_cast<const float4*>(src_global + (size_t)r * K) + tid);
      ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    consumer_sync();
    if (lane == 0) red[warp] = ss;
    consumer_sync();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red[i];
    const float rs = rsqrtf(tot / (float)K + eps);
    if (live) {
      const float4 g = *reinterpret_cast<const float4*>(w + 4 * tid);
      *reinterpret_cast<float4*>(xs + (size_t)r * K + perm_idx(4 * tid, K)) =
          make_float4(v.x * rs * g.x, v.y * rs * g.y, v.z * rs * g.z, v.w * rs * g.w);
    }
  }
  consumer_sync();
}

// fused epilogue of one finished row pair (rows nn, nn+1), executed by lane r for activation row r
template <int NB>
__device__ __forceinline__ void pair_epilogue(int epi, int layer, int pairs, int pair, int r, float y0, float y1) {
  VAURA_SMEM_VIEW();
  const int D = a.D, nn = 2 * pair;
  const float* rope_s = reinterpret_cast<const float*>(smem + sc.rope_off);
  const int* page_s = reinterpret_cast<const int*>(smem + sc.page_off);
  if (epi == EPI_STORE) {
    *reinterpret_cast<float2*>(a.logits + (size_t)r * (2 * pairs) + nn) = make_float2(y0, y1);
  } else if (epi == EPI_RESID) {  // this CTA owns h[nn], h[nn+1]: the running value lives in smem
    float* ho = reinterpret_cast<float*>(smem + sc.hown_off) + r * HOWN + (nn - 2 * sc.own0);
    const float v0 = ho[0] + y0, v1 = ho[1] + y1;
    ho[0] = v0; ho[1] = v1;
    *reinterpret_cast<float2*>(a.h + (size_t)r * D + nn) = make_float2(v0, v1);
  } else if (epi == EPI_SWIGLU) {
    a.act[(size_t)r * a.F + pair] = y0 / (1.f + expf(-y0)) * y1;
  } else {  // EPI_QKV: RoPE (llama.py:633-650) + KV append
    const int sec = nn / D, within = nn % D, hd = within / kHeadDim, e = within % kHeadDim;
    float o0 = y0, o1 = y1;
    if (sec != 2) {
      const float cs = rope_s[e], sn = rope_s[e + 1];
      o0 = y0 * cs - y1 * sn;
      o1 = y1 * cs + y0 * sn;
    }
    if (sec == 0) {
      *reinterpret_cast<float2*>(a.q + (size_t)r * D + within) = make_float2(o0, o1);
    } else {
      const size_t row = ((((size_t)(layer * 2 + (sec - 1)) * a.kv.num_pages + page_s[r]) * a.kv.nhead + hd) *
                              a.kv.page_size + (sc.p % a.kv.page_size)) * kHeadDim;
      *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.kv.pages) + row + e) = make_float2(o0, o1);
    }
  }
}

// GEMV phase over this CTA's row pairs; weights come from the double-buffered smem slots, activations from xs
template <int NB>
__device__ __noinline__ unsigned gemv_phase(int K, int pairs, int epi, int layer, unsigned grp_ctr) {
  VAURA_SMEM_VIEW();
  const int pair_bytes = 4 * K, K8 = K >> 3;
  int p0, p1;
  pair_range(pairs, sc.cta, sc.G, p0, p1);
  const GroupPlan gp(K, p1 - p0, sc.slot_cap);
  int pp = p0;
  for (int g = 0; g < gp.ngroups; ++g) {
    const int n = gp.base + (g < gp.rem ? 1 : 0);
    const int slot = grp_ctr & 1;
    const uint32_t par = (grp_ctr >> 1) & 1;
    if (tid == 0) mb_wait(&full[slot], par);
    consumer_sync();
    const uint8_t* sb = smem + sc.slots_off + (size_t)slot * sc.slot_cap;
    const int item = warp;
    const bool active = item < n * gp.splits;
    const int j = item / gp.splits, seg = item % gp.splits;
    float acc0[NB], acc1[NB];
    if (active) {
      const uint32_t w0 = s_u32(sb) + (uint32_t)(j * pair_bytes);
      const int c0 = (K8 * seg) / gp.splits, c1 = (K8 * (seg + 1)) / gp.splits;
      pair_dot<NB>(w0, w0 + 2 * K, s_u32(xs), K, c0, c1, lane, acc0, acc1);
      if (gp.splits == 1) {
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (lane == r) pair_epilogue<NB>(epi, layer, pairs, pp + j, r, acc0[r], acc1[r]);
      } else {
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (lane == r) {
            part_s[(item * NB + r) * 2] = acc0[r];
            part_s[(item * NB + r) * 2 + 1] = acc1[r];
          }
      }
    }
    if (gp.splits > 1) {
      consumer_sync();
      if (active && seg == 0 && lane < NB) {
        float y0 = 0.f, y1 = 0.f;
        for (int sgm = 0; sgm < gp.splits; ++sgm) {
          y0 += part_s[((item + sgm) * NB + lane) * 2];
          y1 += part_s[((item + sgm) * NB + lane) * 2 + 1];
        }
        pair_epilogue<NB>(epi, layer, pairs, pp + j, lane, y0, y1);
      }
    }
    consumer_sync();  // every warp is done reading the slot (and part_s): refill it with the group after next
    if (tid == 0) issue_next_group();
    pp += n;
    ++grp_ctr;
  }
  return grp_ctr;
}

template <int NB>
__global__ void __launch_bounds__(kThreadsP, 1) decode_step_persistent(const PersistArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kmax = a.D > a.F ? a.D : a.F;
  // layout: [SmemCtx 1 KB][2 weight slots][xs][(attention scratch)][hown][rope][red][split-K partials][pages][mbarriers]
  const bool scr_alias = (size_t)NB * Kmax >= ATT_SCR;  // scratch aliases xs (dead between the QKV phase and the combine)
  const int slots_off = kCtxBytes;
  const int xs_off = slots_off + 2 * a.slot_cap;
  const int scr_off = scr_alias ? xs_off : xs_off + NB * Kmax * 4;
  const int hown_off = xs_off + NB * Kmax * 4 + (scr_alias ? 0 : ATT_SCR * 4);
  const int rope_off = hown_off + NB * HOWN * 4;
  const int red_off = rope_off + kHeadDim * 4;
  const int part_off = red_off + 64 * 4;
  const int page_off = part_off + 128 * 4;
  const int bar_off = page_off + 8 * 4;
  float* xs = reinterpret_cast<float*>(smem + xs_off);
  float* scr = reinterpret_cast<float*>(smem + scr_off);
  float* hown = reinterpret_cast<float*>(smem + hown_off);
  float* rope_s = reinterpret_cast<float*>(smem + rope_off);
  float* part_s = reinterpret_cast<float*>(smem + part_off);
  int* page_s = reinterpret_cast<int*>(smem + page_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + bar_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int D = a.D, F = a.F;
  const int offset = a.state->offset;
  const unsigned epoch = a.state->epoch;
  const int p = offset - 1;  // position fed by this step
  const unsigned nbar = (unsigned)(a.L * 5 + 1);
  unsigned bar_i = 0;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) mb_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    SmemCtx& w = *reinterpret_cast<SmemCtx*>(smem);
    w.a = a;
    w.slot_cap = a.slot_cap; w.slots_off = slots_off; w.xs_off = xs_off; w.scr_off = scr_off; w.hown_off = hown_off;
    w.rope_off = rope_off; w.red_off = red_off; w.part_off = part_off; w.page_off = page_off; w.bar_off = bar_off;
    w.cta = cta; w.G = G; w.p = p;
    int o0_, o1_;
    pair_range(a.D / 2, cta, G, o0_, o1_);
    w.own0 = o0_;
    w.sched.init(cta, G);
    w.issued = 0;
    issue_next_group();  // both slots start filling before the embedding / first norm
    issue_next_group();
  }
  __syncthreads();

  const int qkv_pairs = 3 * D / 2, d_pairs = D / 2, f_pairs = F, head_pairs = a.Kc * a.V / 2;

  int own0, own1;  // this CTA's slice of the residual stream (features [2*own0, 2*own1))
  pair_range(d_pairs, cta, G, own0, own1);

  // optional phase timestamps of CTA 0 (profiles/persist_timing.py; a.timing == nullptr in production)
  int stamp_i = 0;
  auto stamp = [&]() {
    if (a.timing && cta == 0 && tid == 0) a.timing[stamp_i] = now_ns();
    ++stamp_i;
  };
  stamp();

  // per-step constants of the QKV epilogue: RoPE row and KV page of position p
  if (tid < kHeadDim) rope_s[tid] = a.rope[(size_t)p * kHeadDim + tid];
  if (tid < NB) page_s[tid] = a.kv.page_table[tid * a.kv.max_pages_per_seq + p / a.kv.page_size];

  unsigned grp_ctr = 0;

  // ---- embedding (llama.py:455-472): every CTA builds the full rows in smem; owners keep/publish their slice ----
  {
    const int C = a.cond_dim, TD = D - C;
    int vrow = p / a.atpvf;
    if (vrow > a.cond_tokens) vrow = a.cond_tokens;
    for (int r = 0; r < NB; ++r) {
      const int bt = r % a.batch;
      for (int i = tid; i < D; i += kConsumers) {
        float v;
        if (i < C) {
          v = a.cond_rows[((size_t)r * (a.cond_tokens + 1) + vrow) * C + i];
        } else {
          v = 0.f;
          for (int k = 0; k < a.Kc; ++k) {
            const int tok = a.seq[((size_t)bt * a.Kc + k) * a.S + p];
            v += a.tok_tables[((size_t)k * (a.V + 1) + tok) * TD + (i - C)];
          }
        }
        xs[(size_t)r * D + i] = v;
        if (i >= 2 * own0 && i < 2 * own1) {
          hown[r * HOWN + (i - 2 * own0)] = v;
          a.h[(size_t)r * D + i] = v;
        }
      }
    }
    consumer_sync();
  }

  const int npages = p / a.kv.page_size + 1;
  for (int l = 0; l < a.L; ++l) {
    // ---------------- P1: attention_norm + wqkv + RoPE + KV append ----------------
    if (l == 0) stage_norm<NB>(xs, nullptr, a.attn_norm, D);
    else stage_norm<NB>(nullptr, a.h, a.attn_norm + (size_t)l * D, D);
    stamp();
    grp_ctr = gemv_phase<NB>(D, qkv_pairs, EPI_QKV, l, grp_ctr);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- P2: paged attention partials, one CTA per (row, head, page) ----------------
    {
      const int units = NB * a.H * npages;
      const float* kvp = reinterpret_cast<const float*>(a.kv.pages);
      float* part_s = scr;                 // [768]  per-float4 partial dot products
      float* wv = scr + 768;               // [32][96] probability-weighted V rows
      float* ev = scr + 768 + 3072;        // [32]
      float* mq = ev + 32;                 // [96]
      for (int u = cta; u < units; u += G) {
        const int g = u % npages, hh = (u / npages) % a.H, r = u / (npages * a.H);
        const int pos0 = g * a.kv.page_size;
        const int nvalid = min(a.kv.page_size, p + 1 - pos0);
        // a page holds 32 positions x 96 dims contiguously per head: 768 float4 of K and of V; all loads up front
        const float4* k4 = reinterpret_cast<const float4*>(kvp + a.kv.row(l, 0, r, pos0, hh));
        const float4* v4 = reinterpret_cast<const float4*>(kvp + a.kv.row(l, 1, r, pos0, hh));
        const int i0 = tid, i1 = tid + kConsumers;
        const bool ok0 = (i0 / 24) < nvalid, ok1 = i1 < 768 && (i1 / 24) < nvalid;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ka = ok0 ? __ldcg(k4 + i0) : z, kb = ok1 ? __ldcg(k4 + i1) : z;
        const float4 va = ok0 ? __ldcg(v4 + i0) : z, vb = ok1 ? __ldcg(v4 + i1) : z;
        if (tid < kHeadDim) mq[tid] = __ldcg(a.q + (size_t)r * D + hh * kHeadDim + tid);
        consumer_sync();
        {
          const float4 qa = *reinterpret_cast<const float4*>(mq + (i0 % 24) * 4);
          part_s[i0] = ka.x * qa.x + ka.y * qa.y + ka.z * qa.z + ka.w * qa.w;
          if (i1 < 768) {
            const float4 qb = *reinterpret_cast<const float4*>(mq + (i1 % 24) * 4);
            part_s[i1] = kb.x * qb.x + kb.y * qb.y + kb.z * qb.z + kb.w * qb.w;
          }
        }
        consumer_sync();
        if (warp == 0) {  // warp-level softmax over the page
          float s = 0.f;
#pragma unroll 8
          for (int c = 0; c < 24; ++c) s += part_s[lane * 24 + c];
          const bool valid = lane < nvalid;
          s = valid ? s * a.scale : -INFINITY;
          const float m = warp_max(s);
          const float e = valid ? expf(s - m) : 0.f;
          const float lsum = warp_sum(e);
          ev[lane] = e;
          if (lane == 0) {
            float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * ATT_STRIDE;
            part[0] = m;
            part[1] = lsum;
          }
        }
        consumer_sync();
        {
          const float ea = ev[i0 / 24];
          *reinterpret_cast<float4*>(wv + 4 * i0) = make_float4(va.x * ea, va.y * ea, va.z * ea, va.w * ea);
          if (i1 < 768) {
            const float eb = ev[i1 / 24];
            *reinterpret_cast<float4*>(wv + 4 * i1) = make_float4(vb.x * eb, vb.y * eb, vb.z * eb, vb.w * eb);
          }
        }
        consumer_sync();
        if (tid < kHeadDim) {
          float o = 0.f;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) o += wv[j * kHeadDim + tid];
          a.attn_part[(size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * ATT_STRIDE + 4 + tid] = o;
        }
        consumer_sync();
      }
    }
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- P3: combine partials -> xs, wo + residual ----------------
    for (int r = 0; r < NB; ++r) {
      // per (head, page) weight exp(m - M) / sum_g exp(m_g - M) l_g, 8 lanes per head
      float* wgt = part_s;  // [H*8] (H <= 16)
      if (tid < a.H * (kMaxCtx / 32)) {
        const int hh = tid >> 3, g = tid & 7;
        const float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * ATT_STRIDE;
        const float m = g < npages ? __ldcg(part) : -INFINITY;
        const float lg = g < npages ? __ldcg(part + 1) : 0.f;
        float M = m;
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 4));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 1));
        const float e = g < npages ? expf(m - M) : 0.f;
        float den = e * lg;
        den += __shfl_xor_sync(0xffffffffu, den, 4);
        den += __shfl_xor_sync(0xffffffffu, den, 2);
        den += __shfl_xor_sync(0xffffffffu, den, 1);
        wgt[tid] = e / den;
      }
      consumer_sync();
      for (int i = tid; i < D; i += kConsumers) {
        const int hh = i / kHeadDim, dd = i % kHeadDim;
        const float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32)) * ATT_STRIDE + 4 + dd;
        float o = 0.f;
#pragma unroll
        for (int g = 0; g < kMaxCtx / 32; ++g)
          if (g < npages) o = fmaf(wgt[hh * 8 + g], __ldcg(part + g * ATT_STRIDE), o);
        xs[(size_t)r * D + perm_idx(i, D)] = o;
      }
      consumer_sync();
    }
    consumer_sync();
    stamp();
    grp_ctr = gemv_phase<NB>(D, d_pairs, EPI_RESID, l, grp_ctr);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- P4: ffn_norm + w1|w3 + SiLU*mul ----------------
    stage_norm<NB>(nullptr, a.h, a.ffn_norm + (size_t)l * D, D);
    stamp();
    grp_ctr = gemv_phase<NB>(D, f_pairs, EPI_SWIGLU, l, grp_ctr);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- P5: w2 + residual ----------------
    for (int r = 0; r < NB; ++r)
      for (int i = tid; i < (F >> 2); i += kConsumers) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.act + (size_t)r * F) + i);
        *reinterpret_cast<float4*>(xs + (size_t)r * F + perm_idx(4 * i, F)) = v;
      }
    consumer_sync();
    stamp();
    grp_ctr = gemv_phase<NB>(F, d_pairs, EPI_RESID, l, grp_ctr);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();
  }

  // ---------------- final norm + heads ----------------
  stage_norm<NB>(nullptr, a.h, a.final_norm, D);
  stamp();
  grp_ctr = gemv_phase<NB>(D, head_pairs, EPI_STORE, 0, grp_ctr);
  stamp();
  grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
  stamp();

  // ---------------- CFG / sampling / mask-fix / write-back: one warp per (clip, codebook) ----------------
  {
    const int nrows = a.sample.B * a.sample.K;
    const SampleArgs& sa = reinterpret_cast<const SmemCtx*>(smem)->a.sample;
    for (int u = cta + G * warp; u < nrows; u += G * NW) sample_row(sa, u / sa.K, u % sa.K, lane, offset);
  }
  stamp();
  if (cta == 0 && tid == 0) {  // every CTA read offset/epoch before its first barrier arrival
    a.state->offset = offset + 1;
    a.state->epoch = epoch + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t persist_smem(int NB, int D, int F, int& slot_cap) {
  const int Kmax = D > F ? D : F;
  const bool alias = (size_t)NB * Kmax >= (size_t)ATT_SCR;
  const size_t fixed = kCtxBytes + (size_t)NB * Kmax * 4 + (alias ? 0 : ATT_SCR * 4) + NB * HOWN * 4 + kHeadDim * 4 + 64 * 4 +
                       128 * 4 + 8 * 4 + 4 * 8 + 256;
  const size_t budget = 227 * 1024;
  size_t cap = (budget - fixed) / 2;
  const size_t want = (size_t)NW * 4 * 1536;  // 16 row pairs of K=1536
  if (cap > want) cap = want;
  cap &= ~(size_t)1023;
  slot_cap = (int)cap;
  return fixed + 2 * cap;
}

bool persistent_supported(int rows, int D, int F, int page_size) {
  if (rows != 1 && rows != 2 && rows != 4) return false;
  if (page_size != 32) return false;
  if (D % 16 || F % 16 || D > 4096 || F > 1536 * kMaxSplit) return false;
  if (D / 4 > kConsumers) return false;  // stage_norm covers a row with one float4 per thread
  if (2 * ((D / 2 + 131) / 132) > HOWN) return false;  // own slice of h per CTA (>= 132 SMs assumed)
  return true;
}

size_t persistent_attn_part_bytes(int rows, int H) { return (size_t)rows * H * (kMaxCtx / 32) * ATT_STRIDE * sizeof(float); }

template <int NB>
static cudaError_t launch_persist_t(PersistArgs& a, cudaStream_t st) {
  int slot_cap = 0;
  const size_t smem = persist_smem(NB, a.D, a.F, slot_cap);
  if (slot_cap < 4 * a.D || slot_cap < 4 * a.F) return cudaErrorInvalidValue;  // a slot must hold one row pair
  a.slot_cap = slot_cap;
  a.prefetch_ahead = knobs().persist_prefetch;  // groups of L2 prefetch ahead of the smem fill (measured: no gain, off)
  static int grid = 0;
  if (!grid) {
    cudaError_t e = cudaFuncSetAttribute(decode_step_persistent<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_step_persistent<NB>, kThreadsP, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    grid = sms;
  }
  void* args[] = {(void*)&a};
  return cudaLaunchCooperativeKernel((const void*)decode_step_persistent<NB>, dim3(grid), dim3(kThreadsP), args, smem, st);
}

cudaError_t launch_decode_persistent(PersistArgs& a, int rows, cudaStream_t st) {
  switch (rows) {
    case 1: return launch_persist_t<1>(a, st);
    case 2: return launch_persist_t<2>(a, st);
    case 4: return launch_persist_t<4>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vaura
