// Persistent decode-step kernel for the HBM-bound small-batch regime (rows <= 4, fp32 activations).
//
// One CTA per SM runs the WHOLE decode step: embedding -> 24 x {RMSNorm+QKV+RoPE+KV append | paged attention |
// wo+residual | RMSNorm+w1|w3+SiLU*mul | w2+residual} -> final norm + 9 heads -> CFG/sampling/write-back.
// Every CTA owns a fixed contiguous block of output rows of every weight matrix, so its weight bytes for the
// step are one fixed sequence of contiguous chunks.  A dedicated producer warp streams that sequence with
// cp.async.bulk (TMA 1-D) into a ring of 18 KB shared-memory slots guarded by full/empty mbarriers; it never
// waits for the dependency chain, so HBM keeps streaming while the 16 consumer warps sit in a grid barrier or in
// the attention phase (the ring holds ~4.5 us of traffic per SM).  Consumers read weights and activations from
// shared memory only; phases are separated by a device-wide barrier (one atomic + acquire spin per CTA).
//
// Replaces the same reference lines as decode_fp32.cu (llama.py:445-517 for one position) plus the sampling
// stage of sampling.cu; arithmetic (fp32 accumulate order per row aside) is identical to gemv_kernel/attn_kernel.
#include <cstdlib>

#include "sampling.cuh"

namespace vaura {

namespace {

constexpr int NW = 15;                     // consumer warps (+1 producer = 16 warps: 128 registers per thread)
constexpr int kConsumers = NW * 32;        // 480
constexpr int kThreadsP = kConsumers + 32;  // + producer warp
constexpr int SLOT_BYTES = 18 * 1024;      // 3 row pairs of K=1536, or 1 row pair of K=4096
constexpr int ATT_STRIDE = 100;            // floats per attention partial: m, l, pad, pad, o[96]

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
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

// dot products of one row pair (rows w0, w1 of K bf16 in smem) with NB activation rows (xs, permuted, smem).
// All weight loads of a 192-chunk window are issued before the first FMA; every chunk is its own FMA chain.
template <int NB>
__device__ __forceinline__ void pair_dot(const uint4* __restrict__ w0, const uint4* __restrict__ w1,
                                         const float* __restrict__ xs, int K, int lane, float (&acc0)[NB], float (&acc1)[NB]) {
  const int K8 = K >> 3, halfK = K >> 1;
#pragma unroll
  for (int r = 0; r < NB; ++r) acc0[r] = acc1[r] = 0.f;
  for (int base = 0; base < K8; base += 192) {
    uint4 wa[6], wb[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = base + lane + 32 * i;
      if (c < K8) { wa[i] = w0[c]; wb[i] = w1[c]; }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = base + lane + 32 * i;
      if (c < K8) {
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          const float4 xl = *reinterpret_cast<const float4*>(xs + (size_t)r * K + 4 * c);
          const float4 xh = *reinterpret_cast<const float4*>(xs + (size_t)r * K + halfK + 4 * c);
          float p0 = bf16_lo(wa[i].x) * xl.x, p1 = bf16_lo(wb[i].x) * xl.x;
          p0 = fmaf(bf16_hi(wa[i].x), xl.y, p0); p1 = fmaf(bf16_hi(wb[i].x), xl.y, p1);
          p0 = fmaf(bf16_lo(wa[i].y), xl.z, p0); p1 = fmaf(bf16_lo(wb[i].y), xl.z, p1);
          p0 = fmaf(bf16_hi(wa[i].y), xl.w, p0); p1 = fmaf(bf16_hi(wb[i].y), xl.w, p1);
          p0 = fmaf(bf16_lo(wa[i].z), xh.x, p0); p1 = fmaf(bf16_lo(wb[i].z), xh.x, p1);
          p0 = fmaf(bf16_hi(wa[i].z), xh.y, p0); p1 = fmaf(bf16_hi(wb[i].z), xh.y, p1);
          p0 = fmaf(bf16_lo(wa[i].w), xh.z, p0); p1 = fmaf(bf16_lo(wb[i].w), xh.z, p1);
          p0 = fmaf(bf16_hi(wa[i].w), xh.w, p0); p1 = fmaf(bf16_hi(wb[i].w), xh.w, p1);
          acc0[r] += p0; acc1[r] += p1;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) { acc0[r] = warp_sum(acc0[r]); acc1[r] = warp_sum(acc1[r]); }
}

constexpr int HOWN = 32;       // floats per row of this CTA's own slice of the residual stream
constexpr int ATT_SCR = 4096;  // floats of attention scratch

__device__ __forceinline__ int perm_idx(int k, int K) { return ((k & 4) ? (K >> 1) : 0) + ((k >> 3) << 2) + (k & 3); }

struct Phase {
  const uint16_t* W;  // [2*pairs][K] bf16
  int K, pairs, epi;
};

__device__ __forceinline__ void pair_range(int pairs, int cta, int G, int& p0, int& p1) {
  p0 = (int)(((long long)pairs * cta) / G);
  p1 = (int)(((long long)pairs * (cta + 1)) / G);
}

}  // namespace


// Per-thread view of the CTA's shared-memory layout and ring position (kept small: it lives in local memory across
// the __noinline__ phase functions, which keep the kernel's code footprint inside the instruction cache).
struct Ctx {
  uint8_t* slots;
  float *xs, *scr, *hown, *rope_s, *red;
  int* page_s;
  uint64_t *full, *empty;
  unsigned slot_ctr, pair_ctr;
  int nslots, tid, warp, lane, cta, G, p, own0;
};

// xs[r] = (x * rsqrt(mean(x^2)+eps)) * w, permuted.  Source: global h (through L2) or, for layer 0, the embedding
// rows sitting unpermuted in xs (read into registers before the permuted overwrite; syncs in between).
template <int NB>
__device__ __noinline__ void stage_norm(const Ctx& c, const float* src_smem, const float* src_global, const float* w, int K,
                                        float eps) {
  const int tid = c.tid, lane = c.lane, warp = c.warp;
  float* xs = c.xs;
  float* red = c.red;
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

// GEMV phase over this CTA's row pairs; weights come from the smem ring, activations from xs
template <int NB>
__device__ __noinline__ void gemv_phase(const PersistArgs& a, Ctx& c, int K, int pairs, int epi, int layer) {
  const int warp = c.warp, lane = c.lane, p = c.p, D = a.D, F = a.F;
  const int pair_bytes = 4 * K, pps = SLOT_BYTES / pair_bytes;
  unsigned slot_ctr = c.slot_ctr, pair_ctr = c.pair_ctr;
  const int nslots = c.nslots;
  const float* xs = c.xs;
  int p0, p1;
  pair_range(pairs, c.cta, c.G, p0, p1);
  for (int pp = p0; pp < p1; pp += pps) {
    const int n = min(pps, p1 - pp);
    const int s = slot_ctr % nslots;
    const uint32_t par = (slot_ctr / nslots) & 1;
    mb_wait(&c.full[s], par);
    const uint8_t* sb = c.slots + (size_t)s * SLOT_BYTES;
    for (int j = 0; j < n; ++j) {
      if ((int)((pair_ctr + j) % NW) != warp) continue;
      const int pair = pp + j, nn = 2 * pair;
      const uint4* w0 = reinterpret_cast<const uint4*>(sb + (size_t)j * pair_bytes);
      float acc0[NB], acc1[NB];
      pair_dot<NB>(w0, w0 + (K >> 3), xs, K, lane, acc0, acc1);
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        if (lane != r) continue;
        const float y0 = acc0[r], y1 = acc1[r];
        if (epi == EPI_STORE) {
          *reinterpret_cast<float2*>(a.logits + (size_t)r * (2 * pairs) + nn) = make_float2(y0, y1);
        } else if (epi == EPI_RESID) {  // this CTA owns h[nn], h[nn+1]: running value lives in smem
          float* ho = c.hown + r * HOWN + (nn - 2 * c.own0);
          const float v0 = ho[0] + y0, v1 = ho[1] + y1;
          ho[0] = v0; ho[1] = v1;
          *reinterpret_cast<float2*>(a.h + (size_t)r * D + nn) = make_float2(v0, v1);
        } else if (epi == EPI_SWIGLU) {
          a.act[(size_t)r * F + pair] = y0 / (1.f + expf(-y0)) * y1;
        } else {  // EPI_QKV: RoPE (llama.py:633-650) + KV append
          const int sec = nn / D, within = nn % D, hd = within / kHeadDim, e = within % kHeadDim;
          float o0 = y0, o1 = y1;
          if (sec != 2) {
            const float cs = c.rope_s[e], sn = c.rope_s[e + 1];
            o0 = y0 * cs - y1 * sn;
            o1 = y1 * cs + y0 * sn;
          }
          if (sec == 0) {
            *reinterpret_cast<float2*>(a.q + (size_t)r * D + within) = make_float2(o0, o1);
          } else {
            const size_t row = ((((size_t)(layer * 2 + (sec - 1)) * a.kv.num_pages + c.page_s[r]) * a.kv.nhead + hd) *
                                    a.kv.page_size + (p % a.kv.page_size)) * kHeadDim;
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.kv.pages) + row + e) = make_float2(o0, o1);
          }
        }
      }
    }
    pair_ctr += n;
    __syncwarp();
    if (lane == 0) mb_arrive(&c.empty[s]);
    ++slot_ctr;
  }
  c.slot_ctr = slot_ctr;
  c.pair_ctr = pair_ctr;
}

template <int NB>
__global__ void __launch_bounds__(kThreadsP, 1) decode_step_persistent(const PersistArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kmax = a.D > a.F ? a.D : a.F;
  uint8_t* slots = smem;
  float* xs = reinterpret_cast<float*>(smem + (size_t)a.nslots * SLOT_BYTES);  // [NB][Kmax] permuted activations
  float* scr = xs + (size_t)NB * Kmax;                                          // attention scratch
  float* hown = scr + ATT_SCR;                                                  // [NB][HOWN] own slice of h
  float* rope_s = hown + NB * HOWN;                                             // [96] cos,sin of position p
  float* red = rope_s + kHeadDim;                                               // [64]
  int* page_s = reinterpret_cast<int*>(red + 64);                               // [8] KV page of position p per row
  uint64_t* full = reinterpret_cast<uint64_t*>(page_s + 8);
  uint64_t* empty = full + a.nslots;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int D = a.D, F = a.F;
  const int offset = a.state->offset;
  const unsigned epoch = a.state->epoch;
  const int p = offset - 1;  // position fed by this step
  const unsigned nbar = (unsigned)(a.L * 5 + 1);
  unsigned bar_i = 0;

  if (tid == 0) {
    for (int s = 0; s < a.nslots; ++s) {
      mb_init(&full[s], 1);
      mb_init(&empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int qkv_pairs = 3 * D / 2, d_pairs = D / 2, f_pairs = F, head_pairs = a.Kc * a.V / 2;

  // ============================== producer warp ==============================
  if (warp == NW) {
    if (lane == 0) {
      unsigned slot_ctr = 0;
      for (int l = 0; l <= a.L; ++l) {
        const int nph = l < a.L ? 4 : 1;
        for (int ph = 0; ph < nph; ++ph) {
          Phase P;
          if (l == a.L) P = {a.w_heads, D, head_pairs, EPI_STORE};
          else if (ph == 0) P = {a.wqkv + (size_t)l * 3 * D * D, D, qkv_pairs, EPI_QKV};
          else if (ph == 1) P = {a.wo + (size_t)l * D * D, D, d_pairs, EPI_RESID};
          else if (ph == 2) P = {a.w13 + (size_t)l * 2 * F * D, D, f_pairs, EPI_SWIGLU};
          else P = {a.w2 + (size_t)l * D * F, F, d_pairs, EPI_RESID};
          const int pair_bytes = 4 * P.K, pps = SLOT_BYTES / pair_bytes;
          int p0, p1;
          pair_range(P.pairs, cta, G, p0, p1);
          for (int pp = p0; pp < p1; pp += pps) {
            const int n = min(pps, p1 - pp);
            const int s = slot_ctr % a.nslots;
            const uint32_t par = (slot_ctr / a.nslots) & 1;
            mb_wait(&empty[s], par ^ 1);
            mb_expect_tx(&full[s], (uint32_t)(n * pair_bytes));
            bulk_g2s(slots + (size_t)s * SLOT_BYTES, P.W + (size_t)pp * 2 * P.K, (uint32_t)(n * pair_bytes), &full[s]);
            ++slot_ctr;
          }
        }
      }
    }
    return;
  }

  // ============================== consumer warps ==============================
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

  Ctx c;
  c.slots = slots; c.xs = xs; c.scr = scr; c.hown = hown; c.rope_s = rope_s; c.red = red; c.page_s = page_s;
  c.full = full; c.empty = empty; c.slot_ctr = 0; c.pair_ctr = 0; c.nslots = a.nslots; c.tid = tid; c.warp = warp;
  c.lane = lane; c.cta = cta; c.G = G; c.p = p; c.own0 = own0;

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
    if (l == 0) stage_norm<NB>(c, xs, nullptr, a.attn_norm, D, a.eps);
    else stage_norm<NB>(c, nullptr, a.h, a.attn_norm + (size_t)l * D, D, a.eps);
    stamp();
    gemv_phase<NB>(a, c, D, qkv_pairs, EPI_QKV, l);
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
      for (int i = tid; i < D; i += kConsumers) {
        const int hh = i / kHeadDim, dd = i % kHeadDim;
        const float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32)) * ATT_STRIDE;
        float mg[kMaxCtx / 32], M = -INFINITY;
#pragma unroll
        for (int g = 0; g < kMaxCtx / 32; ++g) {
          mg[g] = g < npages ? __ldcg(part + g * ATT_STRIDE) : -INFINITY;
          M = fmaxf(M, mg[g]);
        }
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int g = 0; g < kMaxCtx / 32; ++g) {
          if (g < npages) {
            const float wgt = expf(mg[g] - M);
            num = fmaf(wgt, __ldcg(part + g * ATT_STRIDE + 4 + dd), num);
            den = fmaf(wgt, __ldcg(part + g * ATT_STRIDE + 1), den);
          }
        }
        xs[(size_t)r * D + perm_idx(i, D)] = num / den;
      }
    }
    consumer_sync();
    stamp();
    gemv_phase<NB>(a, c, D, d_pairs, EPI_RESID, l);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- P4: ffn_norm + w1|w3 + SiLU*mul ----------------
    stage_norm<NB>(c, nullptr, a.h, a.ffn_norm + (size_t)l * D, D, a.eps);
    stamp();
    gemv_phase<NB>(a, c, D, f_pairs, EPI_SWIGLU, l);
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
    gemv_phase<NB>(a, c, F, d_pairs, EPI_RESID, l);
    stamp();
    grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();
  }

  // ---------------- final norm + heads ----------------
  stage_norm<NB>(c, nullptr, a.h, a.final_norm, D, a.eps);
  stamp();
  gemv_phase<NB>(a, c, D, head_pairs, EPI_STORE, 0);
  stamp();
  grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
  stamp();

  // ---------------- CFG / sampling / mask-fix / write-back: one warp per (clip, codebook) ----------------
  {
    const int nrows = a.sample.B * a.sample.K;
    for (int u = cta + G * warp; u < nrows; u += G * NW) sample_row(a.sample, u / a.sample.K, u % a.sample.K, lane, offset);
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
static size_t persist_smem(int NB, int D, int F, int& nslots) {
  const int Kmax = D > F ? D : F;
  const size_t fixed = (size_t)NB * Kmax * 4 + ATT_SCR * 4 + NB * HOWN * 4 + kHeadDim * 4 + 64 * 4 + 8 * 4 + 2 * 16 * 8 + 128;
  const size_t budget = 227 * 1024;
  nslots = (int)((budget - fixed) / SLOT_BYTES);
  if (nslots > 16) nslots = 16;
  return fixed + (size_t)nslots * SLOT_BYTES;
}

bool persistent_supported(int rows, int D, int F, int page_size) {
  if (rows != 1 && rows != 2 && rows != 4) return false;
  if (page_size != 32) return false;
  if (D % 16 || F % 16 || 4 * D > SLOT_BYTES || 4 * F > SLOT_BYTES) return false;
  if (D / 4 > kConsumers) return false;  // stage_norm covers a row with one float4 per thread
  if (2 * ((D / 2 + 131) / 132) > HOWN) return false;  // own slice of h per CTA (>= 132 SMs assumed)
  return true;
}

size_t persistent_attn_part_bytes(int rows, int H) { return (size_t)rows * H * (kMaxCtx / 32) * ATT_STRIDE * sizeof(float); }

template <int NB>
static cudaError_t launch_persist_t(PersistArgs& a, cudaStream_t st) {
  int nslots = 0;
  const size_t smem = persist_smem(NB, a.D, a.F, nslots);
  if (nslots < 3) return cudaErrorInvalidValue;
  a.nslots = nslots;
  {
    const char* e = getenv("VAURA_PERSIST_INFLIGHT");
    a.max_inflight = e ? atoi(e) : 4;
    if (a.max_inflight < 1) a.max_inflight = 1;
    if (a.max_inflight > nslots) a.max_inflight = nslots;
  }
  static int grid = 0;
  if (!grid) {
    cudaError_t e = cudaFuncSetAttribute(decode_step_persistent<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
