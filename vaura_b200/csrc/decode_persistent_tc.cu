// Persistent decode-step kernel, tensor-core variant (rows <= 2, fp32-equivalent activations).
//
// Same phase structure as decode_persistent.cu (one cooperative CTA per SM runs the whole step; device-wide
// barriers between the five phases of a layer), but the weight-row dot products run on tcgen05 instead of the CUDA
// cores, which were issue-bound (~340 instructions per row pair):
//   * weights stream through TMA *tensor* loads (cp.async.bulk.tensor.3d, 128B swizzle) straight into the K-major
//     layout the UMMA shared-memory descriptor reads: per 64-wide K block a box of this CTA's rows (<= 64) of the
//     matrix; a dedicated producer thread keeps a 16-slot ring full and runs ahead across phase boundaries, so HBM
//     keeps streaming while the consumers sit in grid barriers, attention or staging;
//   * the activation vector is the B operand with N = 8 columns: each fp32 activation is split into three bf16
//     terms x = x1 + x2 + x3 (exact to ~2^-24), one column each, so bf16 x bf16 products accumulated in fp32 give the
//     fp32-activation result (up to summation order) and greedy token parity with the fp32 reference is preserved;
//   * one thread issues tcgen05.mma (M = 64 weight rows, N = 8, K = 16), accumulating in 8 TMEM columns;
//   * four warps read the accumulator (tcgen05.ld), add the three split columns and run the same fused epilogues
//     (RoPE + KV append, residual, SiLU*mul, logits) as the SIMT variant.
// Reference lines replaced: llama.py:445-517 for one position, as decode_fp32.cu / decode_persistent.cu.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "sampling.cuh"

namespace vaura {

namespace {

constexpr int NC = 14;              // consumer warps: staging, attention, epilogues, sampling
constexpr int kCons = NC * 32;      // 448
constexpr int kThreadsT = 512;      // + warp 14 (TMA producer) + warp 15 (TMEM owner, MMA issuer)
constexpr int kSlot = 8192;         // one ring slot: 64 rows x 128 B (one or several K blocks of this CTA's rows)
constexpr int kNumSlots = 16;
constexpr int kNB = 8;              // UMMA N: 3 split terms x up to 2 activation rows, padded to 8
constexpr int kAttScr = 4096;       // floats of attention scratch (also holds the embedding rows at kernel start)
constexpr int kHOwn = 32;
constexpr int kAttStride = 100;
constexpr uint64_t kTimeoutNs = 2000000000ull;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t t_now() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbi(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mb_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arr(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(su32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbw(uint64_t* b, uint32_t parity) {
  if (mb_try(b, parity)) return;
  const uint64_t t0 = t_now();
  while (!mb_try(b, parity))
    if (t_now() - t0 > kTimeoutNs) __trap();  // protocol bug -> CUDA error instead of a hung GPU
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   su32(dst)),
               "l"(map), "r"(su32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, 128B-swizzled operand tile: SBO = 8 rows * 128 B, version 1, layout SWIZZLE_128B
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNB >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory"); }
__device__ __forceinline__ unsigned ldacq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gbar(unsigned* counter, unsigned target) {
  csync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    if ((int)(ldacq(counter) - target) < 0) {
      const uint64_t t0 = t_now();
      while ((int)(ldacq(counter) - target) < 0)
        if (t_now() - t0 > kTimeoutNs) __trap();
    }
  }
  csync();
}
__device__ __forceinline__ void prange(int pairs, int cta, int G, int& p0, int& p1) {
  p0 = (int)(((unsigned)pairs * (unsigned)cta) / (unsigned)G);
  p1 = (int)(((unsigned)pairs * (unsigned)(cta + 1)) / (unsigned)G);
}

// geometry of one GEMV phase for this CTA
struct PhaseGeo {
  int K, pairs, kblocks, box_rows, tile_bytes, pack, containers;
  __device__ __forceinline__ PhaseGeo(int K_, int pairs_, int G) {
    K = K_;
    pairs = pairs_;
    kblocks = K >> 6;
    box_rows = 2 * ((pairs + G - 1) / G);  // must equal the box height of the tensor map (host: tc_box_rows)
    const int rows8 = (box_rows + 7) & ~7;
    tile_bytes = rows8 * 128;
    pack = 64 / rows8;
    containers = (kblocks + pack - 1) / pack;
  }
};

// CTA-wide context at the start of dynamic shared memory (read with LDS by the __noinline__ helpers)
struct TcCtx {
  PersistArgs a;
  int ring_off, xb_off, scr_off, hown_off, rope_off, red_off, wgt_off, page_off, bar_off;
  int cta, G, p, own0;
  uint32_t tmem;
};
constexpr int kTcCtxBytes = 1024;
static_assert(sizeof(TcCtx) <= kTcCtxBytes, "TcCtx must fit its reserved block");

// dynamic shared memory, rounded up to 1024 B in the shared window (the 128B swizzle of TMA / UMMA works on address bits)
#define TC_SMEM_BASE()                                                     \
  extern __shared__ uint8_t smem_raw[];                                    \
  uint8_t* smem = smem_raw + ((1024u - (su32(smem_raw) & 1023u)) & 1023u);

#define TC_VIEW()                                                          \
  TC_SMEM_BASE()                                                           \
  const TcCtx& sc = *reinterpret_cast<const TcCtx*>(smem);                 \
  const PersistArgs& a = sc.a;                                             \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;           \
  (void)a; (void)tid; (void)warp; (void)lane;

// write one float4 (elements k..k+3 of activation row r) as three bf16 split terms into the B operand tile
__device__ __forceinline__ void write_split(uint8_t* xb, int r, int k, float4 v) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  uint32_t w[3][2];
#pragma unroll
  for (int e = 0; e < 4; e += 2) {
    __nv_bfloat16 t[3][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float x = f[e + u];
      t[0][u] = __float2bfloat16_rn(x);
      x -= __bfloat162float(t[0][u]);
      t[1][u] = __float2bfloat16_rn(x);
      x -= __bfloat162float(t[1][u]);
      t[2][u] = __float2bfloat16_rn(x);
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
      w[s][e >> 1] = (uint32_t)__bfloat16_as_ushort(t[s][0]) | ((uint32_t)__bfloat16_as_ushort(t[s][1]) << 16);
  }
  const int kb = k >> 6, kl = k & 63, chunk = kl >> 3, within = kl & 7;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int n = 3 * r + s;
    uint8_t* dst = xb + kb * 1024 + n * 128 + ((chunk ^ n) << 4) + within * 2;
    *reinterpret_cast<uint2*>(dst) = make_uint2(w[s][0], w[s][1]);
  }
}

// RMSNorm (llama.py:147-158) of NB rows -> split bf16 B operand.  Source: global h (L2) or smem rows (layer 0).
template <int NB>
__device__ __noinline__ void stage_norm_tc(const float* src_smem, const float* src_global, const float* w, int K) {
  TC_VIEW();
  float* red = reinterpret_cast<float*>(smem + sc.red_off);
  uint8_t* xb = smem + sc.xb_off;
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
    csync();
    if (lane == 0) red[warp] = ss;
    csync();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) tot += red[i];
    const float rs = rsqrtf(tot / (float)K + a.eps);
    if (live) {
      const float4 g = *reinterpret_cast<const float4*>(w + 4 * tid);
      write_split(xb, r, 4 * tid, make_float4(v.x * rs * g.x, v.y * rs * g.y, v.z * rs * g.z, v.w * rs * g.w));
    }
  }
}

// consumers: publish the staged B operand to the MMA thread
__device__ __forceinline__ void publish_x(uint64_t* xready) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to UMMA
  csync();
  if (threadIdx.x == 0) mb_arr(xready);
}

// fused epilogue of one finished row pair (same arithmetic as decode_persistent.cu: pair_epilogue)
__device__ __forceinline__ void pair_epi(int epi, int layer, int pairs, int pair, int r, float y0, float y1) {
  TC_VIEW();
  const int D = a.D, nn = 2 * pair;
  if (epi == EPI_STORE) {
    *reinterpret_cast<float2*>(a.logits + (size_t)r * (2 * pairs) + nn) = make_float2(y0, y1);
  } else if (epi == EPI_RESID) {
    float* ho = reinterpret_cast<float*>(smem + sc.hown_off) + r * kHOwn + (nn - 2 * sc.own0);
    const float v0 = ho[0] + y0, v1 = ho[1] + y1;
    ho[0] = v0; ho[1] = v1;
    *reinterpret_cast<float2*>(a.h + (size_t)r * D + nn) = make_float2(v0, v1);
  } else if (epi == EPI_SWIGLU) {
    a.act[(size_t)r * a.F + pair] = y0 / (1.f + expf(-y0)) * y1;
  } else {
    const float* rope_s = reinterpret_cast<const float*>(smem + sc.rope_off);
    const int* page_s = reinterpret_cast<const int*>(smem + sc.page_off);
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

// consumer warps 0-3: read the accumulator of the finished phase and run the epilogue
template <int NB>
__device__ __noinline__ void phase_epilogue(int K, int pairs, int epi, int layer, uint32_t done_parity) {
  TC_VIEW();
  if (warp >= 4) return;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sc.bar_off);
  uint64_t* done = bars + 2 * kNumSlots + 1;
  mbw(done, done_parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float c[8];
  tmem_ld8(sc.tmem + ((uint32_t)(warp * 32) << 16), c);  // M=64: rows 16w..16w+15 live in lanes 0..15 of quadrant w
  int p0, p1;
  prange(pairs, sc.cta, sc.G, p0, p1);
  const int row = warp * 16 + lane;
  const bool valid = lane < 16 && row < 2 * (p1 - p0);
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const float y = c[3 * r] + c[3 * r + 1] + c[3 * r + 2];
    const float yo = __shfl_xor_sync(0xffffffffu, y, 1);  // partner row of the pair
    if (valid && !(lane & 1)) pair_epi(epi, layer, pairs, p0 + (row >> 1), r, y, yo);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  (void)K;
}

}  // namespace

template <int NB>
__global__ void __launch_bounds__(kThreadsT, 1)
decode_step_persistent_tc(const PersistArgs a, const __grid_constant__ CUtensorMap tm_qkv,
                          const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_w13,
                          const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_heads) {
  static_assert(3 * NB <= kNB, "split terms must fit the UMMA N");
  TC_SMEM_BASE()
  const int Kmax = a.D > a.F ? a.D : a.F;
  // layout: [ctx 1 KB][ring 16 x 8 KB + 8 KB guard][B operand Kmax/64 KB][scratch 16 KB][hown][rope][red][wgt][page][barriers]
  const int ring_off = kTcCtxBytes;
  const int xb_off = ring_off + (kNumSlots + 1) * kSlot;
  const int scr_off = xb_off + (Kmax >> 6) * 1024;
  const int hown_off = scr_off + kAttScr * 4;
  const int rope_off = hown_off + NB * kHOwn * 4;
  const int red_off = rope_off + kHeadDim * 4;
  const int wgt_off = red_off + 64 * 4;
  const int page_off = wgt_off + 128 * 4;
  const int bar_off = page_off + 8 * 4;
  uint8_t* ring = smem + ring_off;
  uint8_t* xb = smem + xb_off;
  float* scr = reinterpret_cast<float*>(smem + scr_off);
  float* hown = reinterpret_cast<float*>(smem + hown_off);
  float* rope_s = reinterpret_cast<float*>(smem + rope_off);
  float* wgt = reinterpret_cast<float*>(smem + wgt_off);
  int* page_s = reinterpret_cast<int*>(smem + page_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + bar_off);
  uint64_t* empty = full + kNumSlots;
  uint64_t* xready = empty + kNumSlots;
  uint64_t* done = xready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int D = a.D, F = a.F;
  const int offset = a.state->offset;
  const unsigned epoch = a.state->epoch;
  const int p = offset - 1;
  const unsigned nbar = (unsigned)(a.L * 5 + 1);
  unsigned bar_i = 0;
  const int qkv_pairs = 3 * D / 2, d_pairs = D / 2, f_pairs = F, head_pairs = a.Kc * a.V / 2;

  if (tid == 0) {
    for (int s = 0; s < kNumSlots; ++s) { mbi(&full[s], 1); mbi(&empty[s], 1); }
    mbi(xready, 1);
    mbi(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    TcCtx& w = *reinterpret_cast<TcCtx*>(smem);
    w.a = a;
    w.ring_off = ring_off; w.xb_off = xb_off; w.scr_off = scr_off; w.hown_off = hown_off; w.rope_off = rope_off;
    w.red_off = red_off; w.wgt_off = wgt_off; w.page_off = page_off; w.bar_off = bar_off;
    w.cta = cta; w.G = G; w.p = p;
    int o0, o1;
    prange(d_pairs, cta, G, o0, o1);
    w.own0 = o0;
  }
  // rows 3*NB..7 of the B operand are never written: they must be finite (zero)
  for (int i = tid; i < (Kmax >> 6) * 1024 / 16; i += kThreadsT) reinterpret_cast<uint4*>(xb)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 15) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) reinterpret_cast<TcCtx*>(smem)->tmem = tmem;

  // ============================== TMA producer: warp 14, one thread ==============================
  if (warp == 14) {
    if (lane == 0) {
      unsigned ctr = 0;
      for (int l = 0; l <= a.L; ++l) {
        const int nph = l < a.L ? 4 : 1;
        for (int ph = 0; ph < nph; ++ph) {
          const CUtensorMap* tm = l == a.L ? &tm_heads : (ph == 0 ? &tm_qkv : ph == 1 ? &tm_wo : ph == 2 ? &tm_w13 : &tm_w2);
          const PhaseGeo geo(l == a.L ? D : (ph == 3 ? F : D),
                             l == a.L ? head_pairs : (ph == 0 ? qkv_pairs : ph == 2 ? f_pairs : d_pairs), G);
          int p0, p1;
          prange(geo.pairs, cta, G, p0, p1);
          const int row0 = 2 * p0, layer = l == a.L ? 0 : l;
          for (int c = 0; c < geo.containers; ++c) {
            const int s = ctr % kNumSlots;
            const uint32_t par = (ctr / kNumSlots) & 1;
            mbw(&empty[s], par ^ 1);
            const int nkb = min(geo.pack, geo.kblocks - c * geo.pack);
            mb_tx(&full[s], (uint32_t)(nkb * geo.box_rows * 128));
            for (int j = 0; j < nkb; ++j)
              tma3(ring + (size_t)s * kSlot + j * geo.tile_bytes, tm, &full[s], 0, row0, layer * geo.kblocks + c * geo.pack + j);
            ++ctr;
          }
        }
      }
    }
  } else if (warp == 15) {
    // ============================== MMA issuer: warp 15, one thread ==============================
    if (lane == 0) {
      unsigned ctr = 0, phase_idx = 0;
      for (int l = 0; l <= a.L; ++l) {
        const int nph = l < a.L ? 4 : 1;
        for (int ph = 0; ph < nph; ++ph) {
          const PhaseGeo geo(l == a.L ? D : (ph == 3 ? F : D),
                             l == a.L ? head_pairs : (ph == 0 ? qkv_pairs : ph == 2 ? f_pairs : d_pairs), G);
          mbw(xready, phase_idx & 1);  // B operand of this phase is staged (and the previous epilogue has read TMEM)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int c = 0; c < geo.containers; ++c) {
            const int s = ctr % kNumSlots;
            const uint32_t par = (ctr / kNumSlots) & 1;
            mbw(&full[s], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int nkb = min(geo.pack, geo.kblocks - c * geo.pack);
            for (int j = 0; j < nkb; ++j) {
              const int kb = c * geo.pack + j;
              const uint64_t ad = sdesc(su32(ring + (size_t)s * kSlot + j * geo.tile_bytes));
              const uint64_t bd = sdesc(su32(xb + kb * 1024));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma(tmem, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) != 0);
            }
            commit(&empty[s]);
            ++ctr;
          }
          commit(done);
          ++phase_idx;
        }
      }
    }
  } else {
    // ============================== consumer warps 0-13 ==============================
    int own0, own1;
    prange(d_pairs, cta, G, own0, own1);
    int stamp_i = 0;
    auto stamp = [&]() {
      if (a.timing && cta == 0 && tid == 0) a.timing[stamp_i] = t_now();
      ++stamp_i;
    };
    stamp();
    unsigned phase_idx = 0;
    if (tid < kHeadDim) rope_s[tid] = a.rope[(size_t)p * kHeadDim + tid];
    if (tid < NB) page_s[tid] = a.kv.page_table[tid * a.kv.max_pages_per_seq + p / a.kv.page_size];

    // ---- embedding (llama.py:455-472): full rows in smem scratch; owners keep/publish their slice of h ----
    {
      const int C = a.cond_dim, TD = D - C;
      int vrow = p / a.atpvf;
      if (vrow > a.cond_tokens) vrow = a.cond_tokens;
      for (int r = 0; r < NB; ++r) {
        const int bt = r % a.batch;
        for (int i = tid; i < D; i += kCons) {
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
          scr[(size_t)r * D + i] = v;
          if (i >= 2 * own0 && i < 2 * own1) {
            hown[r * kHOwn + (i - 2 * own0)] = v;
            a.h[(size_t)r * D + i] = v;
          }
        }
      }
      csync();
    }

    const int npages = p / a.kv.page_size + 1;
    for (int l = 0; l < a.L; ++l) {
      // ---------------- P1: attention_norm + wqkv + RoPE + KV append ----------------
      if (l == 0) stage_norm_tc<NB>(scr, nullptr, a.attn_norm, D);
      else stage_norm_tc<NB>(nullptr, a.h, a.attn_norm + (size_t)l * D, D);
      publish_x(xready);
      stamp();
      phase_epilogue<NB>(D, qkv_pairs, EPI_QKV, l, phase_idx++ & 1);
      stamp();
      gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();

      // ---------------- P2: paged attention partials, one CTA per (row, head, page) ----------------
      {
        const int units = NB * a.H * npages;
        const float* kvp = reinterpret_cast<const float*>(a.kv.pages);
        float* part_s = scr;
        float* wv = scr + 768;
        float* ev = scr + 768 + 3072;
        float* mq = ev + 32;
        for (int u = cta; u < units; u += G) {
          const int g = u % npages, hh = (u / npages) % a.H, r = u / (npages * a.H);
          const int pos0 = g * a.kv.page_size;
          const int nvalid = min(a.kv.page_size, p + 1 - pos0);
          const float4* k4 = reinterpret_cast<const float4*>(kvp + a.kv.row(l, 0, r, pos0, hh));
          const float4* v4 = reinterpret_cast<const float4*>(kvp + a.kv.row(l, 1, r, pos0, hh));
          const int i0 = tid, i1 = tid + kCons;
          const bool ok0 = (i0 / 24) < nvalid, ok1 = i1 < 768 && (i1 / 24) < nvalid;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 ka = ok0 ? __ldcg(k4 + i0) : z, kb = ok1 ? __ldcg(k4 + i1) : z;
          const float4 va = ok0 ? __ldcg(v4 + i0) : z, vb = ok1 ? __ldcg(v4 + i1) : z;
          if (tid < kHeadDim) mq[tid] = __ldcg(a.q + (size_t)r * D + hh * kHeadDim + tid);
          csync();
          {
            const float4 qa = *reinterpret_cast<const float4*>(mq + (i0 % 24) * 4);
            part_s[i0] = ka.x * qa.x + ka.y * qa.y + ka.z * qa.z + ka.w * qa.w;
            if (i1 < 768) {
              const float4 qb = *reinterpret_cast<const float4*>(mq + (i1 % 24) * 4);
              part_s[i1] = kb.x * qb.x + kb.y * qb.y + kb.z * qb.z + kb.w * qb.w;
            }
          }
          csync();
          if (warp == 0) {
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
              float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * kAttStride;
              part[0] = m;
              part[1] = lsum;
            }
          }
          csync();
          {
            const float ea = ev[i0 / 24];
            *reinterpret_cast<float4*>(wv + 4 * i0) = make_float4(va.x * ea, va.y * ea, va.z * ea, va.w * ea);
            if (i1 < 768) {
              const float eb = ev[i1 / 24];
              *reinterpret_cast<float4*>(wv + 4 * i1) = make_float4(vb.x * eb, vb.y * eb, vb.z * eb, vb.w * eb);
            }
          }
          csync();
          if (tid < kHeadDim) {
            float o = 0.f;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) o += wv[j * kHeadDim + tid];
            a.attn_part[(size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * kAttStride + 4 + tid] = o;
          }
          csync();
        }
      }
      stamp();
      gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();

      // ---------------- P3: combine partials -> B operand, wo + residual ----------------
      for (int r = 0; r < NB; ++r) {
        if (tid < a.H * (kMaxCtx / 32)) {
          const int hh = tid >> 3, g = tid & 7;
          const float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32) + g) * kAttStride;
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
        csync();
        for (int i4 = tid; i4 < (D >> 2); i4 += kCons) {
          const int i = 4 * i4, hh = i / kHeadDim, dd = i % kHeadDim;  // 96 % 4 == 0: a float4 stays inside one head
          const float* part = a.attn_part + (size_t)((r * a.H + hh) * (kMaxCtx / 32)) * kAttStride + 4 + dd;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < kMaxCtx / 32; ++g)
            if (g < npages) {
              const float wg = wgt[hh * 8 + g];
              const float4 pv = __ldcg(reinterpret_cast<const float4*>(part + g * kAttStride));
              o.x = fmaf(wg, pv.x, o.x); o.y = fmaf(wg, pv.y, o.y); o.z = fmaf(wg, pv.z, o.z); o.w = fmaf(wg, pv.w, o.w);
            }
          write_split(xb, r, i, o);
        }
        csync();
      }
      publish_x(xready);
      stamp();
      phase_epilogue<NB>(D, d_pairs, EPI_RESID, l, phase_idx++ & 1);
      stamp();
      gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();

      // ---------------- P4: ffn_norm + w1|w3 + SiLU*mul ----------------
      stage_norm_tc<NB>(nullptr, a.h, a.ffn_norm + (size_t)l * D, D);
      publish_x(xready);
      stamp();
      phase_epilogue<NB>(D, f_pairs, EPI_SWIGLU, l, phase_idx++ & 1);
      stamp();
      gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();

      // ---------------- P5: w2 + residual ----------------
      for (int r = 0; r < NB; ++r)
        for (int i = tid; i < (F >> 2); i += kCons)
          write_split(xb, r, 4 * i, __ldcg(reinterpret_cast<const float4*>(a.act + (size_t)r * F) + i));
      publish_x(xready);
      stamp();
      phase_epilogue<NB>(F, d_pairs, EPI_RESID, l, phase_idx++ & 1);
      stamp();
      gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();
    }

    // ---------------- final norm + heads ----------------
    stage_norm_tc<NB>(nullptr, a.h, a.final_norm, D);
    publish_x(xready);
    stamp();
    phase_epilogue<NB>(D, head_pairs, EPI_STORE, 0, phase_idx++ & 1);
    stamp();
    gbar(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    stamp();

    // ---------------- CFG / sampling / mask-fix / write-back ----------------
    {
      const SampleArgs& sa = reinterpret_cast<const TcCtx*>(smem)->a.sample;
      const int nrows = sa.B * sa.K;
      for (int u = cta + G * warp; u < nrows; u += G * NC) sample_row(sa, u / sa.K, u % sa.K, lane, offset);
    }
    stamp();
    if (cta == 0 && tid == 0) {
      a.state->offset = offset + 1;
      a.state->epoch = epoch + 1;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 15) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tc_box_rows(int pairs, int G) { return 2 * ((pairs + G - 1) / G); }

// K-block-major weights X_t[l][kb][n][64]: dims {64, N, (K/64)*L}; a box = this CTA's rows of one K block, contiguous
static bool weight_map(CUtensorMap* m, const void* base, int K, int rows, int layers, int box_rows) {
  static EncodeTiledFnT fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return false;
    fn = reinterpret_cast<EncodeTiledFnT>(p);
  }
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)layers * (K / 64)};
  cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t tc_smem(int NB, int D, int F) {
  const int Kmax = D > F ? D : F;
  return (size_t)kTcCtxBytes + (kNumSlots + 1) * kSlot + (size_t)(Kmax >> 6) * 1024 + kAttScr * 4 + NB * kHOwn * 4 + kHeadDim * 4 +
         64 * 4 + 128 * 4 + 8 * 4 + (2 * kNumSlots + 2) * 8 + 16 + 1024 /* alignment slack */;
}

bool persistent_tc_supported(int rows, int D, int F, int page_size, int head_pairs, int f_pairs, int sms) {
  if (rows != 1 && rows != 2) return false;
  if (page_size != 32 || D % 64 || F % 64 || D / 4 > kCons) return false;
  if ((size_t)rows * D > (size_t)kAttScr) return false;                      // embedding rows live in the scratch
  if (tc_box_rows(head_pairs, sms) > 64 || tc_box_rows(f_pairs, sms) > 64) return false;  // UMMA M = 64 rows per CTA
  if (2 * ((D / 2 + sms - 1) / sms) > kHOwn) return false;
  return tc_smem(rows, D, F) <= 227 * 1024;
}

template <int NB>
static cudaError_t launch_tc_t(PersistArgs& a, cudaStream_t st) {
  static int grid = 0;
  static CUtensorMap maps[5];
  static const void* key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const size_t smem = tc_smem(NB, a.D, a.F);
  if (!grid) {
    cudaError_t e = cudaFuncSetAttribute(decode_step_persistent_tc<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_step_persistent_tc<NB>, kThreadsT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    grid = sms;
  }
  const void* now[5] = {a.wqkv_t, a.wo_t, a.w13_t, a.w2_t, a.w_heads_t};
  if (memcmp(key, now, sizeof(key)) != 0) {  // (re)encode the weight tensor maps when the weights change
    const int D = a.D, F = a.F;
    bool ok = weight_map(&maps[0], a.wqkv_t, D, 3 * D, a.L, tc_box_rows(3 * D / 2, grid)) &&
              weight_map(&maps[1], a.wo_t, D, D, a.L, tc_box_rows(D / 2, grid)) &&
              weight_map(&maps[2], a.w13_t, D, 2 * F, a.L, tc_box_rows(F, grid)) &&
              weight_map(&maps[3], a.w2_t, F, D, a.L, tc_box_rows(D / 2, grid)) &&
              weight_map(&maps[4], a.w_heads_t, D, a.Kc * a.V, 1, tc_box_rows(a.Kc * a.V / 2, grid));
    if (!ok) return cudaErrorUnknown;
    memcpy(key, now, sizeof(key));
  }
  void* args[] = {(void*)&a, (void*)&maps[0], (void*)&maps[1], (void*)&maps[2], (void*)&maps[3], (void*)&maps[4]};
  return cudaLaunchCooperativeKernel((const void*)decode_step_persistent_tc<NB>, dim3(grid), dim3(kThreadsT), args, smem, st);
}

cudaError_t launch_decode_persistent_tc(PersistArgs& a, int rows, cudaStream_t st) {
  switch (rows) {
    case 1: return launch_tc_t<1>(a, st);
    case 2: return launch_tc_t<2>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vaura
