// Cluster-persistent decode-step kernel for the HBM-bound small-batch regime (rows <= 2, fp32-exact activations).
//
// 32 clusters x 4 CTAs (one CTA per SM, 128 SMs; a B200 co-schedules at most 15 clusters of 8 but 33 of 4, see
// profiles/probes/cluster_occ.cu) run the WHOLE decode step.  Attention head h is served by the two clusters 2h, 2h+1:
// both compute q|k|v and the attention of head h (same bytes, read from HBM once and from L2 the second time) so that
// no attention data ever crosses a cluster; everything else is partitioned without overlap:
//
//   layer l:  x -> [RMSNorm . wqkv rows of head h, K split over the 4 CTAs] -> DSMEM reduce-scatter + RoPE + KV append
//               + all-gather -> attention of head h (positions split over the CTAs, partials combined through DSMEM)
//               -> wo[768s + 192r .. +192, head h]  -> 64-bit fixed-point red.add into the residual    | grid barrier
//             x -> [RMSNorm . w1|w3 of 32 hidden units per CTA] -> SiLU*mul -> DSMEM all-gather of the cluster's 128
//               -> w2[384r .. +384, units of the cluster] -> fixed-point red.add                         | grid barrier
//
// Synchronisation.  Measured on B200 (profiles/probes/dsmem_latency.cu): while cp.async.bulk copies are in flight, any
// release/acquire barrier - DSMEM mbarrier arrive, barrier.cluster, or a global red.release/ld.acquire counter - costs
// ~1.5 us instead of 0.35 us (the fence waits for the outstanding copies), whereas fence-free mechanisms keep their idle
// latency.  So this kernel has NO barrier between the phases of a step:
//   * cluster exchanges use st.async (remote shared-memory store that completes transaction bytes on the destination's
//     mbarrier; 0.24 us with or without streaming); two mbarriers alternate because a peer can be one exchange ahead;
//   * the residual stream is self-validating: every phase output is a fresh [rows][1536] buffer of int64 words holding
//     (2^-32 fixed-point sum << 6) + number of contributions, accumulated with red.global.add.u64.  A reader spins on
//     its own four words until the count says every cluster has contributed.  Integer adds commute, so the value does
//     not depend on arrival order (bit-reproducible).  The 48 buffers of a step are cleared after the step's only
//     device-wide barrier (logits complete -> sampling).
//
// Weights: a producer thread copies the CTA's weights with cp.async.bulk into a ring of 12 KB shared-memory slots; a
// slot holds two 16x16 bf16 tiles for each of the 12 compute warps, stored in the register order of the
// mma.sync.m16n8k16 A fragment (one conflict-free LDS.128 per tile per lane).  The bytes come from two strictly
// sequential streams (vaura_b200/weights.py: pack_cluster_stream): the q|k|v stream shared by the two clusters of a
// head and the CTA's private stream.  Slots are grouped into units of 3 or 4 with one full/empty mbarrier pair each
// (13 units = 45 slots per layer); the producer runs ahead of the consumers through barriers, attention and staging.
//
// Arithmetic: activations stay fp32-exact.  An fp32 value is split into three bf16 terms (hi, mid, lo; 8+8+8 mantissa
// bits) that occupy three of the eight B-operand columns of mma.sync.m16n8k16 (two sequence rows use six); products of
// bf16 values are exact in fp32 and the tensor core accumulates in fp32, so the result equals an fp32 dot product up
// to summation order.  The three column sums are added in the epilogue.
//
// Replaces the same reference lines as decode_persistent.cu (llama.py:445-517 for one position, vaura_model.py:775-827).
#include <cstdlib>
#include <type_traits>

#include "sampling.cuh"

namespace vaura {

namespace {

constexpr int CW = 12;                     // compute warps
constexpr int kCT = CW * 32;               // 384 compute threads
constexpr int kThreadsC = kCT + 32;        // + one producer warp
constexpr int CL = 4;                      // CTAs per cluster
constexpr int NCL = 32;                    // clusters; head = cluster / 2
constexpr int SLOT = 12288;                // ring slot: 12 warps x 2 tiles x 512 B
constexpr int DM = 1536, FF = 4096, NHEAD = 16;
constexpr int QROWS = 3 * kHeadDim;        // 288 q|k|v rows of one head (18 row tiles); K slice per CTA = 384 (24 k-tiles)
constexpr int QOWN = QROWS / CL;           // 72 rows reduced by each CTA
constexpr int HROWS = 288;                 // heads rows per cluster (9 * 1024 / 32)
constexpr int HOWN = HROWS / CL;           // 72
constexpr int HU = FF / NCL;               // 128 hidden units per cluster
constexpr int HUC = HU / CL;               // 32 hidden units per CTA
constexpr int W2_ROWS = DM / CL;           // 384 wo / w2 rows per CTA
constexpr int SH_SLOTS = 24;               // shared stream per layer: 18 q|k|v + 6 wo slots (two k halves of 3)
constexpr int PRIV_SLOTS = 24, HEAD_SLOTS = 18;  // private stream per layer: 16 w1|w3 + 8 w2 slots; heads at the end
constexpr int HEAD_UNITS = 6;
constexpr int MAXIT = 5;                   // attention items (positions) per warp: 4 CTAs x 12 warps x 5 = 240 old positions
constexpr int WP_STRIDE = 100;             // floats per attention partial: m, l, pad, pad, o[96]
constexpr float kFixScale = 4294967296.0f; // residual stream fixed point: 2^-32
constexpr float kFixInv = 2.3283064365386963e-10f;

template <int NB>
struct Lay {
  static constexpr int nslot = NB == 1 ? 17 : 15;
  static constexpr int upl = 12 + NB;                       // units per layer: 6 qkv + NB wo (3 slots each) + 4 w13 + 2 w2 (4 slots each)
  static constexpr int bkt = 96 * NB;                       // bytes of one k-tile of B fragments: 3 NB columns x 4 lanes x 8 B
  static constexpr int ring = 0;
  static constexpr int bx = ring + nslot * SLOT;            // B fragments of the normed residual [96 k-tiles][3 NB cols][4][2] u32
  static constexpr int bs = bx + 96 * bkt;                  // B fragments of attn out (6 k-tiles) / hidden (8 k-tiles)
  static constexpr int ra = bs + 8 * bkt;                   // alias group A:
  static constexpr int halves = ra;                         //   qkv / heads K halves of this CTA [2][NB][288] f32
  static constexpr int qrecv = ra + 2 * QROWS * NB * 4;     //   all-to-all target [4 src][NB][288] f32 (remote-written)
  static constexpr int ra_bytes = (2 + CL) * QROWS * NB * 4;//   also: attention partials of the 12 warps [12][NB][100], w13 partials
  static constexpr int rb = ra + ra_bytes;                  // attention partials of the 4 CTAs [4][NB][100] (remote-written);
  static constexpr int rb_bytes = CL * NB * WP_STRIDE * 4;  //   also the heads reduce-scatter target [4 src][NB][72]
  static constexpr int qkv = rb + rb_bytes;                 // [NB][288] f32: q|k|v of the head after RoPE
  static constexpr int hrecv = qkv + NB * QROWS * 4;        // [NB][128] f32 (remote-written all-gather)
  static constexpr int xown = hrecv + NB * HU * 4;          // [NB][384] f32: this CTA's rows of the phase input
  static constexpr int red = xown + NB * W2_ROWS * 4;       // [12][NB] f32 sums of squares, [16 + 12*NB] score maxima
  static constexpr int rope = red + 256;                    // [96] f32
  static constexpr int bars = rope + kHeadDim * 4;          // full[upl], empty[upl], xbar[2]
  static constexpr int sargs = bars + 32 * 8;               // SampleArgs copy
  static constexpr int kofft = sargs + 256;                 // [12 warps][MAXIT] int: K/V row offsets of a warp's positions
  static constexpr int total = kofft + CW * MAXIT * 4;
  static_assert(CW * NB * WP_STRIDE * 4 <= ra_bytes, "attention warp partials fit alias group A");
  static_assert(CW * 64 * NB * 4 <= ra_bytes, "w13 partials fit alias group A");
  static_assert(sizeof(SampleArgs) <= 256, "SampleArgs copy");
  static_assert(total <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mb_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin waits poll without reading %globaltimer (a read costs on the order of a microsecond and would quantise every
// wait); the timeout that turns a protocol bug into a trap (-> CUDA error) instead of a hung GPU counts SM cycles.
constexpr long long kSpinTimeoutCycles = 4000000000ll;  // ~2 s
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
  if (mb_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mb_try_wait(bar, parity))
    if (clock64() - t0 > kSpinTimeoutCycles) __trap();
}
// Warp-uniform variants for the producer warp: every lane executes the statement with identical operands and one
// elected lane issues.  Inside `if (lane == 0)` the compiler wraps every copy instruction in an ELECT + R2UR loop
// (~0.13 us per copy of the issuing thread's time, profiles/probes/tma_rate.cu).
__device__ __forceinline__ void mb_expect_tx_e(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_e(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t}" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2_e(const void* src, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}" ::"l"(src), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// device-wide barrier among the compute threads of all CTAs (co-resident: one CTA per SM, checked at launch)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  consumer_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    if ((int)(ld_acquire(counter) - target) < 0) {
      const long long t0 = clock64();
      while ((int)(ld_acquire(counter) - target) < 0)
        if (clock64() - t0 > kSpinTimeoutCycles) __trap();
    }
  }
  consumer_sync();
}
// fence-free arrival / wait on the device-wide counter (probe: 0.45 us with or without streaming).  It only tells that
// every CTA has ISSUED its residual adds; completeness of the data is checked on the words themselves.
__device__ __forceinline__ void grid_arrive_relaxed(unsigned* counter) {
  consumer_sync();
  if (threadIdx.x == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void grid_wait_relaxed(unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if ((int)(v - target) < 0) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (clock64() - t0 > kSpinTimeoutCycles) __trap();
      } while ((int)(v - target) < 0);
    }
  }
  consumer_sync();
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void hw_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// remote shared-memory store of 16 bytes that completes 16 transaction bytes on the mbarrier `rbar` of the destination
// CTA (both addresses from mapa): data and completion travel together, no fence involved
__device__ __forceinline__ void st_async_f4(uint32_t raddr, float4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(raddr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar)
               : "memory");
}
// issued where written (asm volatile): the compiler must not sink the K/V prefetch to its first use
__device__ __forceinline__ float4 ldg_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_async_f1(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(raddr), "f"(v), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint4& a, const uint2& b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}

// residual word = (2^-32 fixed-point value << 6) + contribution count
constexpr int kCntBits = 6;
__device__ __forceinline__ long long f2fix(float v) { return __float2ll_rn(v * kFixScale) * (1ll << kCntBits) + 1; }
__device__ __forceinline__ float fix2f(long long w) { return __ll2float_rn(w >> kCntBits) * kFixInv; }
__device__ __forceinline__ void ld_relaxed_x4(const long long* p, long long (&w)[4]) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory");
}
__device__ __forceinline__ void red_add_fix(long long* p, long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// fp32 -> (hi, mid, lo) bf16 bit patterns with hi + mid + lo == v exactly (round-to-nearest residuals)
__device__ __forceinline__ void split3(float v, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  const __nv_bfloat16 l = __float2bfloat16_rn(r2);
  hi = __bfloat16_as_ushort(h); mid = __bfloat16_as_ushort(m); lo = __bfloat16_as_ushort(l);
}
// B-fragment address (bytes from the fragment array base) of element (k, column n): mma.m16n8k16 .col B operand,
// lane = n*4 + (k%8)/2, register (k%16)/8, half k%2.  Only the 3 NB live columns are stored (12 NB lanes per k-tile);
// the other lanes feed zeros.
template <int NB>
__device__ __forceinline__ uint32_t bfrag_off(int k, int n) {
  const int kk = k & 15;
  return (uint32_t)(((k >> 4) * 12 * NB + n * 4 + ((kk & 7) >> 1)) * 8 + (kk >> 3) * 4 + (kk & 1) * 2);
}
// stage the pair (v0, v1) = elements (k, k+1), k even, of sequence row b into B fragments (three split columns)
template <int NB>
__device__ __forceinline__ void stage_pair(uint8_t* base, int k, int b, float v0, float v1) {
  uint32_t h0, m0, l0, h1, m1, l1;
  split3(v0, h0, m0, l0);
  split3(v1, h1, m1, l1);
  *reinterpret_cast<uint32_t*>(base + bfrag_off<NB>(k, 3 * b + 0)) = h0 | (h1 << 16);
  *reinterpret_cast<uint32_t*>(base + bfrag_off<NB>(k, 3 * b + 1)) = m0 | (m1 << 16);
  *reinterpret_cast<uint32_t*>(base + bfrag_off<NB>(k, 3 * b + 2)) = l0 | (l1 << 16);
}
template <int NB>
__device__ __forceinline__ void stage_one(uint8_t* base, int k, int b, float v) {
  uint32_t h, m, l;
  split3(v, h, m, l);
  *reinterpret_cast<uint16_t*>(base + bfrag_off<NB>(k, 3 * b + 0)) = (uint16_t)h;
  *reinterpret_cast<uint16_t*>(base + bfrag_off<NB>(k, 3 * b + 1)) = (uint16_t)m;
  *reinterpret_cast<uint16_t*>(base + bfrag_off<NB>(k, 3 * b + 2)) = (uint16_t)l;
}

// accumulator fragment (rows g, g+8; columns 2t, 2t+1) -> out[half] = sum of the three split columns of sequence row
// tq (columns 3tq..3tq+2), valid in the lanes tq < NB: row 0 = cols 0,1 (lane tq 0) + col 2 (lane tq 1, c[even]);
// row 1 = col 3 (lane tq 1, c[odd]) + cols 4,5 (lane tq 2).  One shuffle per half.
template <int NB>
__device__ __forceinline__ void quad_reduce(const float (&c)[4], int tq, float (&out)[2]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const float ca = c[2 * hf], cb = c[2 * hf + 1];
    const float send = tq == 1 ? ca : ca + cb;
    const float recv = __shfl_down_sync(0xffffffffu, send, 1);
    out[hf] = (tq == 0 ? ca + cb : cb) + recv;
  }
}

}  // namespace

// TM: phase timestamps of one CTA (profiles/cluster_timing.py); the production instance carries none of that code
template <int NB, bool TM>
__global__ void __launch_bounds__(kThreadsC, 1) decode_step_cluster(const __grid_constant__ PersistArgs a) {
  using LY = Lay<NB>;
  constexpr int NSLOT = LY::nslot;
  constexpr int UPL = LY::upl;          // units per layer: 6 qkv + NB wo + 4 w13 + 2 w2
  constexpr int U_WO = 6, U_W13 = 6 + NB, U_W2 = 10 + NB;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int rank = (int)cluster_rank();
  const int cl = cta / CL;                // cluster
  const int head = cl >> 1, sh = cl & 1;  // attention head; which of the head's two clusters
  // The attention block (q|k|v, attention, wo) works on ONE sequence row per cluster: with two rows the two clusters of
  // a head take one row each (and the full wo K range), with one row both compute the row and split the wo K range.
  const int ab = NB == 1 ? 0 : sh;
  const uint32_t sbase = s_u32(smem);
  const uint32_t full0 = sbase + LY::bars, empty0 = full0 + UPL * 8, xbar = empty0 + UPL * 8;
  const int L = a.L;

  if (tid == 0) {
    for (int i = 0; i < UPL; ++i) { mb_init(full0 + 8 * i, 1); mb_init(empty0 + 8 * i, CW); }
    mb_init(xbar, 1);
    mb_init(xbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *reinterpret_cast<SampleArgs*>(smem + LY::sargs) = a.sample;
  }
  __syncthreads();
  hw_cluster_sync();  // every CTA of the cluster has initialised its barriers before any remote arrive / store

  if (warp == CW) {
    // ------------------------------- producer: the CTA's two weight streams -------------------------------
    // (all 32 lanes walk the loop with identical values; one elected lane issues each copy)
    {
      uint64_t pol_first, pol_last;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
      // shared stream of (head, rank): read by this CTA and by its twin in the other cluster of the head
      const uint8_t* shared = a.wstream + (size_t)(head * CL + rank) * L * SH_SLOTS * SLOT;
      const uint8_t* priv = a.wstream + (size_t)NHEAD * CL * L * SH_SLOTS * SLOT +
                            (size_t)cta * ((size_t)L * PRIV_SLOTS + HEAD_SLOTS) * SLOT;
      const int nq_layers = UPL * L, NQ = nq_layers + HEAD_UNITS;
      // unit q of the step: barrier index, slots, source, L2 policy
      auto unit = [&](int q, int& i, int& n, const uint8_t*& src, bool& is_shared) {
        if (q >= nq_layers) { i = q - nq_layers; n = 3; src = priv + ((size_t)L * PRIV_SLOTS + 3 * i) * SLOT; is_shared = false; return; }
        const int l = q / UPL;
        i = q % UPL;
        if (i < U_WO) { n = 3; src = shared + ((size_t)l * SH_SLOTS + 3 * i) * SLOT; is_shared = true; }
        else if (i < U_W13) { n = 3; src = shared + ((size_t)l * SH_SLOTS + 18 + 3 * (NB == 1 ? sh : i - U_WO)) * SLOT; is_shared = true; }
        else if (i < U_W2) { n = 4; src = priv + ((size_t)l * PRIV_SLOTS + 4 * (i - U_W13)) * SLOT; is_shared = false; }
        else { n = 4; src = priv + ((size_t)l * PRIV_SLOTS + 16 + 4 * (i - U_W2)) * SLOT; is_shared = false; }
      };
      int rel = 0;       // units whose release by the 12 compute warps has been observed (in order)
      int freed = 0;     // ring slots of those units
      int issued = 0;    // ring slots handed to the copy engine so far
      int wslot = 0;     // ring position of the next slot
      const int ring_limit = a.prefetch_ahead >= 4 && a.prefetch_ahead < NSLOT ? a.prefetch_ahead : NSLOT;  // experiment knob
      // L2 prefetch cursor: runs `ahead` units in front of the shared-memory fill so that a refill after a burst comes
      // from L2 (measured: 1 unit is best, deeper prefetch slows the residual exchange)
      const int ahead = a.pace_cycles >= 0 ? a.pace_cycles : 1;
      int pq = 0;
      for (int q = 0; q < NQ; ++q) {
        int i, n;
        const uint8_t* src;
        bool is_shared;
        unit(q, i, n, src, is_shared);
        while (pq < NQ && pq <= q + ahead) {
          int pi, pn;
          const uint8_t* psrc;
          bool ps;
          unit(pq, pi, pn, psrc, ps);
          if (pq > q) bulk_prefetch_l2_e(psrc, (uint32_t)pn * SLOT);
          ++pq;
        }
        while (issued + n - freed > ring_limit) {
          int ri, rn;
          const uint8_t* rsrc;
          bool rs;
          unit(rel, ri, rn, rsrc, rs);
          mb_wait(empty0 + 8 * ri, (uint32_t)(rel < nq_layers ? rel / UPL : L) & 1u);
          freed += rn;
          ++rel;
        }
        mb_expect_tx_e(full0 + 8 * i, (uint32_t)n * SLOT);
        {  // the unit's slots are contiguous in the stream and in the ring up to the wrap: one copy, two at the wrap
          const int n1 = min(n, NSLOT - wslot);
          bulk_g2s_e(sbase + LY::ring + wslot * SLOT, src, (uint32_t)n1 * SLOT, full0 + 8 * i, is_shared ? pol_last : pol_first);
          if (n1 < n)
            bulk_g2s_e(sbase + LY::ring, src + (size_t)n1 * SLOT, (uint32_t)(n - n1) * SLOT, full0 + 8 * i,
                       is_shared ? pol_last : pol_first);
          wslot = wslot + n >= NSLOT ? wslot + n - NSLOT : wslot + n;
        }
        issued += n;
      }
      // HBM idles from here to the end of the step (device barrier, sampling, launch gap): pull the head of the next
      // step's streams into L2 meanwhile
      for (int q = 0; q < a.tail_units && q < nq_layers; ++q) {
        int i, n;
        const uint8_t* src;
        bool is_shared;
        unit(q, i, n, src, is_shared);
        bulk_prefetch_l2_e(src, (uint32_t)n * SLOT);
      }
    }
  } else {
    // ------------------------------------------- compute warps -------------------------------------------
    const int gq = lane >> 2, tq = lane & 3;
    const int offset = a.state->offset;
    const unsigned epoch = a.state->epoch;
    const int p = offset - 1;  // position fed by this step
    if (cta == 0 && tid == 0 && a.step_times) a.step_times[offset] = now_ns();  // step-to-step latency (bench.py: p50)
    const unsigned nbar = (unsigned)(2 * L + 1);
    unsigned bar_i = 0;
    uint32_t xc = 0;           // cluster exchanges done so far: exchange xc uses xbar[xc & 1], parity (xc >> 1) & 1
    int rslot = 0;             // ring position of the next slot to consume
    float* red = reinterpret_cast<float*>(smem + LY::red);
    float* rope_s = reinterpret_cast<float*>(smem + LY::rope);
    float* xown = reinterpret_cast<float*>(smem + LY::xown);
    float* qkv_s = reinterpret_cast<float*>(smem + LY::qkv);
    float* halves = reinterpret_cast<float*>(smem + LY::halves);
    const uint32_t aring = sbase + LY::ring + warp * 1024 + lane * 16;
    const uint32_t abx = sbase + LY::bx + lane * 8, abs_ = sbase + LY::bs + lane * 8;
    // B fragments of R sequence rows: 3R live columns = 12R lanes, 96R bytes per k-tile
    auto ldb = [&](uint32_t base, int kt, int R) { return lane < 12 * R ? lds_u2(base + kt * 96 * R) : make_uint2(0u, 0u); };
    // address of tile t of the s-th slot after the ring cursor
    auto tile_addr = [&](int s, int t) {
      int sl = rslot + s;
      if (sl >= NSLOT) sl -= NSLOT;
      return aring + sl * SLOT + t * 512;
    };
    auto advance = [&](int n) { rslot += n; if (rslot >= NSLOT) rslot -= NSLOT; };
    auto release = [&](int i) { __syncwarp(); if (lane == 0) mb_arrive(empty0 + 8 * i); };
    // exchange protocol: every CTA of the cluster receives `bytes` in total through st.async
    auto xarm = [&](uint32_t bytes) { if (tid == 0) mb_expect_tx(xbar + 8 * (xc & 1), bytes); };
    auto xpush = [&](uint32_t local_addr, int dst, float4 v) {
      st_async_f4(mapa_u32(local_addr, dst), v, mapa_u32(xbar + 8 * (xc & 1), dst));
    };
    auto xwait = [&]() { mb_wait(xbar + 8 * (xc & 1), (xc >> 1) & 1u); ++xc; };

    int stamp_i = 0;
    auto stamp = [&]() {
      if (TM) {
        if (cta == a.timing_cta && tid == 0) a.timing[stamp_i] = now_ns();
        ++stamp_i;
      }
    };
    stamp();
    bool dbg = false;  // fine-grained stamps of one layer (profiles/cluster_timing.py)
    auto dstamp = [&](int k) { if (TM && dbg) a.timing[256 + k] = now_ns(); };

    if (tid < kHeadDim) rope_s[tid] = a.rope[(size_t)p * kHeadDim + tid];

    // attention work split: old positions [0, p) in four chunks, one per CTA; warp w takes items w, w+12, ...
    // row offsets live in shared memory (one table per warp), validity in a bit mask: registers are the scarce resource
    const int chunk = (p + CL - 1) / CL;
    const int j0 = rank * chunk, j1 = min(p, j0 + chunk);
    int* kofft = reinterpret_cast<int*>(smem + LY::kofft) + warp * MAXIT;
    unsigned kvmask = 0;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = j0 + warp + CW * it;
      if (j < j1) kvmask |= 1u << it;
      if (lane == 0) {
        const int page = j < j1 ? a.kv.page_table[ab * a.kv.max_pages_per_seq + j / a.kv.page_size] : 0;
        kofft[it] = ((page * a.kv.nhead + head) * a.kv.page_size + (j % a.kv.page_size)) * kHeadDim;
      }
    }
    __syncwarp();
    auto kvalid = [&](int it) { return ((kvmask >> it) & 1u) && lane < 24; };
    auto koff = [&](int it) { return kofft[it] + 4 * lane; };
    const size_t kv_half = (size_t)a.kv.num_pages * a.kv.nhead * a.kv.page_size * kHeadDim;  // floats of K (or V) per layer
    const bool new_warp = rank == CL - 1 && warp == CW - 1;  // takes the position written by this step
    const float* kvbase = reinterpret_cast<const float*>(a.kv.pages);

    // ---- staging: x -> B fragments of x * norm_w (three bf16 split columns per row), sum of squares -> rstd ----
    // every thread owns 4 consecutive features; `own0`: first of the 384 rows this CTA carries into the residual add
    float rstd[NB];
    auto stage_rows = [&](auto load4, const float* norm_w, auto rows_tag, int row0) {  // rows [row0, row0 + R)
      constexpr int R = decltype(rows_tag)::value;
      float ss[R];
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(norm_w) + tid);
#pragma unroll
      for (int b = 0; b < R; ++b) {
        const float4 v = load4(row0 + b);
        dstamp(43);
        ss[b] = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        if (4 * tid >= W2_ROWS * rank && 4 * tid < W2_ROWS * (rank + 1))
          *reinterpret_cast<float4*>(xown + (row0 + b) * W2_ROWS + (4 * tid - W2_ROWS * rank)) = v;
        stage_pair<R>(smem + LY::bx, 4 * tid, b, v.x * g4.x, v.y * g4.y);
        stage_pair<R>(smem + LY::bx, 4 * tid + 2, b, v.z * g4.z, v.w * g4.w);
        ss[b] = warp_sum(ss[b]);
        if (lane == 0) red[warp * NB + b] = ss[b];
      }
      consumer_sync();
#pragma unroll
      for (int b = 0; b < R; ++b) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < CW; ++w) tot += red[w * NB + b];
        rstd[R == NB ? b : 0] = rsqrtf(tot / (float)DM + a.eps);
      }
    };
    // phase output n (1..2L) lives in residual buffer n; a word is complete when its count reaches `expect`
    auto xbuf = [&](int n) { return a.xfix + (size_t)n * NB * DM; };
    auto load_fix = [&](const long long* buf, int expect) {
      return [=](int b) {
        const long long* src = buf + (size_t)b * DM + 4 * tid;
        long long w[4];
        ld_relaxed_x4(src, w);
        if (((w[0] & w[1] & w[2] & w[3]) & 63) != expect || ((w[0] | w[1] | w[2] | w[3]) & 63) != expect) {
          const long long t0 = clock64();
          do {
            ld_relaxed_x4(src, w);
            if (clock64() - t0 > kSpinTimeoutCycles) __trap();
          } while (((w[0] & w[1] & w[2] & w[3]) & 63) != expect || ((w[0] | w[1] | w[2] | w[3]) & 63) != expect);
        }
        return make_float4(fix2f(w[0]), fix2f(w[1]), fix2f(w[2]), fix2f(w[3]));
      };
    };
    // embedding (llama.py:455-472): conditioning row | sum of the 9 folded token tables
    auto load_embed = [&](int b) {
      const int C = a.cond_dim, TD = DM - C, i = 4 * tid;
      if (i < C) {
        int vrow = p / a.atpvf;
        if (vrow > a.cond_tokens) vrow = a.cond_tokens;
        return __ldg(reinterpret_cast<const float4*>(a.cond_rows + ((size_t)b * (a.cond_tokens + 1) + vrow) * C + i));
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int bt = b % a.batch;
      for (int k = 0; k < a.Kc; ++k) {
        const int tok = a.seq[((size_t)bt * a.Kc + k) * a.S + p];
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.tok_tables + ((size_t)k * (a.V + 1) + tok) * TD + (i - C)));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      return v;
    };
    // output rows [384 rank + 16 rt, +16) of a residual phase, R sequence rows starting at row0: add the partial (and,
    // where `carry`, the phase input)
    auto resid_add = [&](const float (&acc)[4], long long* dst, int rt, bool carry, auto rows_tag, int row0) {
      constexpr int R = decltype(rows_tag)::value;
      float o[2];
      quad_reduce<R>(acc, tq, o);
      if (tq < R) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int rl = 16 * rt + gq + 8 * hf;
          float v = o[hf];
          if (carry) v += xown[(row0 + tq) * W2_ROWS + rl];
          red_add_fix(dst + (size_t)(row0 + tq) * DM + W2_ROWS * rank + rl, f2fix(v));
        }
      }
    };
    // K-split GEMV of 288 rows (18 row tiles) x this CTA's 384 features for R sequence rows: 6 units of 3 slots; warp
    // (rh, kq) owns row tiles 3rh..3rh+2 and k-tiles 12kq..12kq+11.  Leaves halves[kq][b][row]
    auto ksplit_288 = [&](int use, auto rows_tag) {
      constexpr int R = decltype(rows_tag)::value;
      const int rh = warp >> 1, kq = warp & 1;
#pragma unroll
      for (int rt = 0; rt < 3; ++rt) {
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hk = 0; hk < 2; ++hk) {
          const int u = 2 * rt + hk;
          uint2 bq[6];
#pragma unroll
          for (int kk = 0; kk < 6; ++kk) bq[kk] = ldb(abx, 24 * rank + 12 * kq + 6 * hk + kk, R);
          dstamp(2 * u);
          mb_wait(full0 + 8 * u, (uint32_t)use & 1u);
          dstamp(2 * u + 1);
          uint4 A[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) A[j] = lds_u4(tile_addr(j >> 1, j & 1));
#pragma unroll
          for (int j = 0; j < 6; j += 2) { mma16816(acc0, A[j], bq[j]); mma16816(acc1, A[j + 1], bq[j + 1]); }
          release(u);
          advance(3);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc0[e] += acc1[e];
        float o[2];
        quad_reduce<R>(acc0, tq, o);
        if (tq < R) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) halves[(kq * R + tq) * QROWS + 16 * (3 * rh + rt) + gq + 8 * hf] = o[hf];
        }
      }
      dstamp(12);
    };
    using One = std::integral_constant<int, 1>;
    using All = std::integral_constant<int, NB>;

    for (int l = 0; l < L; ++l) {
      const uint32_t par = (uint32_t)l & 1u;
      if (TM) dbg = cta == a.timing_cta && tid == 0 && l == L / 2;
      long long* x_mid = xbuf(2 * l + 1);
      long long* x_out = xbuf(2 * l + 2);

      // ---- K rows of this head's old positions (sequence row ab): issued now, consumed after the QKV exchange; the V
      //      rows are only pulled into L2 for now (three 128-byte lines per row) ----
      float4 kreg[MAXIT], vreg[MAXIT];
      const float* kl = kvbase + (size_t)l * 2 * kv_half;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        kreg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid(it)) {
          kreg[it] = ldg_cg_f4(kl + koff(it));
          if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(kl + kv_half + koff(it)));
        }
      }

      // ================= attention block, sequence row ab: RMSNorm . wqkv rows of the head, K split over the 4 CTAs =================
      if (l == 0) stage_rows(load_embed, a.attn_norm, One{}, ab);
      else stage_rows(load_fix(xbuf(2 * l), NCL), a.attn_norm + (size_t)l * DM, One{}, ab);
      const float rstd_a = rstd[0];
      stamp();
      ksplit_288(l, One{});
      // V rows: issued after the QKV MMAs (register budget), hidden behind the exchange and the score pass
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) vreg[it] = kvalid(it) ? ldg_cg_f4(kl + kv_half + koff(it)) : make_float4(0.f, 0.f, 0.f, 0.f);
      {
        // all-to-all: every CTA receives the K-slice sums of all four CTAs, [src][288]
        xarm(CL * QROWS * 4);
        consumer_sync();
        dstamp(13);
        if (tid < QROWS / 4) {
          const float4 h0 = *reinterpret_cast<const float4*>(halves + 4 * tid);
          const float4 h1 = *reinterpret_cast<const float4*>(halves + QROWS + 4 * tid);
          const float4 v = make_float4(h0.x + h1.x, h0.y + h1.y, h0.z + h1.z, h0.w + h1.w);
          const uint32_t local = sbase + LY::qrecv + (rank * QROWS + 4 * tid) * 4;
#pragma unroll
          for (int s = 0; s < CL; ++s) xpush(local, s, v);
        }
        dstamp(14);
        xwait();
        dstamp(15);
        // every CTA: sum the 4 K slices, apply rstd and RoPE (llama.py:633-650); rows [72 rank, +72) append K/V
        const float* qr = reinterpret_cast<const float*>(smem + LY::qrecv);
        if (tid < QROWS) {  // 288 = 9 whole warps
          const int i = tid;
          float y = 0.f;
#pragma unroll
          for (int s = 0; s < CL; ++s) y += qr[s * QROWS + i];
          y *= rstd_a;
          const float other = __shfl_xor_sync(0xffffffffu, y, 1);  // partner of the RoPE pair (rows 2m, 2m+1)
          const int sec = i / kHeadDim, d = i % kHeadDim;
          float o = y;
          if (sec != 2) {
            const float cs = rope_s[d & ~1], sn = rope_s[(d & ~1) + 1];
            o = (d & 1) ? y * cs + other * sn : y * cs - other * sn;
          }
          // one sequence row: the twin cluster computes the same values, cluster 2h appends; two rows: each its own
          if (sec != 0 && (NB > 1 || sh == 0) && i / QOWN == rank) {
            const int page = a.kv.page_table[ab * a.kv.max_pages_per_seq + p / a.kv.page_size];
            float* dstp = reinterpret_cast<float*>(a.kv.pages) + (size_t)l * 2 * kv_half + (size_t)(sec - 1) * kv_half +
                          ((size_t)(page * a.kv.nhead + head) * a.kv.page_size + (p % a.kv.page_size)) * kHeadDim + d;
            *dstp = o;
          }
          qkv_s[i] = o;
        }
        dstamp(16);
        consumer_sync();
        dstamp(17);
      }
      stamp();

      // ================= attention of the head, sequence row ab: positions split over the CTAs =================
      {
        float* wp = reinterpret_cast<float*>(smem + LY::ra);  // [12][100]
        const float4 q4 = lane < 24 ? *reinterpret_cast<const float4*>(qkv_s + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        float sc[MAXIT + 1];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
          const float4 k4 = kreg[it];
          sc[it] = k4.x * q4.x + k4.y * q4.y + k4.z * q4.z + k4.w * q4.w;
        }
        float4 vnew = make_float4(0.f, 0.f, 0.f, 0.f);
        sc[MAXIT] = 0.f;
        if (new_warp && lane < 24) {
          const float4 kn = *reinterpret_cast<const float4*>(qkv_s + kHeadDim + 4 * lane);
          vnew = *reinterpret_cast<const float4*>(qkv_s + 2 * kHeadDim + 4 * lane);
          sc[MAXIT] = kn.x * q4.x + kn.y * q4.y + kn.z * q4.z + kn.w * q4.w;
        }
        // six independent butterfly reductions, interleaved
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int it = 0; it <= MAXIT; ++it) sc[it] += __shfl_xor_sync(0xffffffffu, sc[it], o);
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) sc[it] = ((kvmask >> it) & 1u) ? sc[it] * a.scale : -INFINITY;
        sc[MAXIT] = new_warp ? sc[MAXIT] * a.scale : -INFINITY;
        float m = sc[MAXIT];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) m = fmaxf(m, sc[it]);
        if (lane == 0) red[16 + warp] = m;
        consumer_sync();
        float M = red[16];
#pragma unroll
        for (int w = 1; w < CW; ++w) M = fmaxf(M, red[16 + w]);  // max over the CTA's positions (-inf if it has none)
        float lsum = 0.f;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (M > -INFINITY) {
#pragma unroll
          for (int it = 0; it <= MAXIT; ++it) {
            const float e = sc[it] > -INFINITY ? __expf(sc[it] - M) : 0.f;
            const float4 v4 = it < MAXIT ? vreg[it < MAXIT ? it : 0] : vnew;
            lsum += e;
            o.x = fmaf(e, v4.x, o.x); o.y = fmaf(e, v4.y, o.y); o.z = fmaf(e, v4.z, o.z); o.w = fmaf(e, v4.w, o.w);
          }
        }
        {
          float* w = wp + warp * WP_STRIDE;
          if (lane == 0) { w[0] = M; w[1] = lsum; }
          if (lane < 24) *reinterpret_cast<float4*>(w + 4 + 4 * lane) = o;
        }
        xarm(CL * WP_STRIDE * 4);
        dstamp(18);
        consumer_sync();
        // CTA partial (M, l, -, -, o[96]) -> every CTA of the cluster, 25 float4
        if (tid < 25) {
          const int c4 = tid;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < CW; ++w) {
            const float4 v = *reinterpret_cast<const float4*>(wp + w * WP_STRIDE + 4 * c4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          if (c4 == 0) { acc.x = wp[0]; acc.z = 0.f; acc.w = 0.f; }  // every warp stored the same M
          const uint32_t local = sbase + LY::rb + (rank * WP_STRIDE + 4 * c4) * 4;
#pragma unroll
          for (int s = 0; s < CL; ++s) xpush(local, s, acc);
        }
        dstamp(19);
        xwait();
        dstamp(20);
        if (tid < kHeadDim) {
          const int d = tid;
          const float* ar = reinterpret_cast<const float*>(smem + LY::rb);
          float Mx = -INFINITY;
#pragma unroll
          for (int s = 0; s < CL; ++s) Mx = fmaxf(Mx, ar[s * WP_STRIDE]);
          float den = 0.f, O = 0.f;
#pragma unroll
          for (int s = 0; s < CL; ++s) {
            const float ms = ar[s * WP_STRIDE];
            const float f = ms > -INFINITY ? expf(ms - Mx) : 0.f;
            den = fmaf(f, ar[s * WP_STRIDE + 1], den);
            O = fmaf(f, ar[s * WP_STRIDE + 4 + d], O);
          }
          stage_one<1>(smem + LY::bs, d, 0, O / den);
        }
        consumer_sync();
      }
      stamp();

      // ==== wo[384 rank .. +384, head] for row ab; warp w owns row tiles 2w, 2w+1; one row: this cluster's half of the
      //      head's 96 features (k-tiles 3sh..3sh+2), two rows: all six k-tiles in two units; residual into x_mid ====
      {
        float acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int kh = NB == 1 ? sh : u;  // k half held by this unit
          dstamp(21);
          mb_wait(full0 + 8 * (U_WO + u), par);
          dstamp(22);
          uint4 A[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) A[j] = lds_u4(tile_addr(j >> 1, j & 1));
#pragma unroll
          for (int j = 0; j < 6; ++j) mma16816(acc[j / 3], A[j], ldb(abs_, 3 * kh + j % 3, 1));
          release(U_WO + u);
          advance(3);
        }
        // the phase input is carried by one cluster per sequence row: cluster 0 (one row), the head-0 clusters (two rows)
        const bool carry = NB == 1 ? cl == 0 : head == 0;
        resid_add(acc[0], x_mid, 2 * warp, carry, One{}, ab);
        resid_add(acc[1], x_mid, 2 * warp + 1, carry, One{}, ab);
      }
      grid_arrive_relaxed(&a.state->barrier);
      stamp();
      dstamp(41);
      grid_wait_relaxed(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      dstamp(42);

      // ============ MLP, all rows: RMSNorm . w1|w3 of this CTA's 32 hidden units (K split over the 12 warps) ============
      stage_rows(load_fix(x_mid, NB == 1 ? NCL : NHEAD), a.ffn_norm + (size_t)l * DM, All{}, 0);
      stamp();
      {
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint2 b0 = ldb(abx, 8 * warp + 2 * u, NB), b1 = ldb(abx, 8 * warp + 2 * u + 1, NB);
          dstamp(24 + 2 * u);
          mb_wait(full0 + 8 * (U_W13 + u), par);
          dstamp(25 + 2 * u);
          uint4 A[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) A[j] = lds_u4(tile_addr(j >> 1, j & 1));
#pragma unroll
          for (int j = 0; j < 8; ++j) mma16816(acc[j & 3], A[j], (j >> 2) ? b1 : b0);
          release(U_W13 + u);
          advance(4);
        }
        float* part = reinterpret_cast<float*>(smem + LY::ra);  // [12][NB][64]
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float o[2];
          quad_reduce<NB>(acc[r], tq, o);
          if (tq < NB) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) part[(warp * NB + tq) * 64 + 16 * r + gq + 8 * hf] = o[hf];
          }
        }
        xarm(CL * NB * HUC * 4);
        dstamp(32);
        consumer_sync();
        dstamp(33);
        if (tid < HUC * NB) {  // hidden unit u of this CTA: rows 16R+g (w1) and 16R+8+g (w3), R = u / 8, g = u % 8
          const int b = tid / HUC, u = tid % HUC;
          const int r1 = 16 * (u >> 3) + (u & 7);
          float y1 = 0.f, y3 = 0.f;
#pragma unroll
          for (int w = 0; w < CW; ++w) { y1 += part[(w * NB + b) * 64 + r1]; y3 += part[(w * NB + b) * 64 + r1 + 8]; }
          const float rs = b ? rstd[NB - 1] : rstd[0];  // no runtime-indexed register array
          y1 *= rs; y3 *= rs;
          const float hv = y1 / (1.f + expf(-y1)) * y3;
          const uint32_t local = sbase + LY::hrecv + (b * HU + rank * HUC + u) * 4;
#pragma unroll
          for (int s = 0; s < CL; ++s) st_async_f1(mapa_u32(local, s), hv, mapa_u32(xbar + 8 * (xc & 1), s));
        }
        dstamp(34);
        xwait();
        dstamp(35);
        if (tid < HU * NB / 2) {
          const int b = tid / (HU / 2), k = 2 * (tid % (HU / 2));
          const float2 v = *reinterpret_cast<const float2*>(smem + LY::hrecv + (b * HU + k) * 4);
          stage_pair<NB>(smem + LY::bs, k, b, v.x, v.y);
        }
        consumer_sync();
      }
      stamp();
      // ==== w2[384 rank .. +384, hidden units of the cluster]; warp w owns row tiles 2w, 2w+1; residual into x_out ====
      {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
          dstamp(36 + 2 * u);
          mb_wait(full0 + 8 * (U_W2 + u), par);
          dstamp(37 + 2 * u);
          uint4 A[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) A[j] = lds_u4(tile_addr(j >> 1, j & 1));
#pragma unroll
          for (int j = 0; j < 8; j += 2) { mma16816(acc0, A[j], ldb(abs_, j, NB)); mma16816(acc1, A[j + 1], ldb(abs_, j + 1, NB)); }
          release(U_W2 + u);
          advance(4);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc0[e] += acc1[e];
          resid_add(acc0, x_out, 2 * warp + u, cl == 0, All{}, 0);
        }
      }
      grid_arrive_relaxed(&a.state->barrier);
      stamp();
      grid_wait_relaxed(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
    }

    dbg = false;
    // ============ final norm + heads: the cluster owns rows [288 cl, +288), K split over its CTAs ============
    {
      stage_rows(load_fix(xbuf(2 * L), NCL), a.final_norm, All{}, 0);
      ksplit_288(L, All{});
      // reduce-scatter: rows [72 dst, +72) of every sequence row go to CTA dst, [src][b][72]
      xarm(CL * NB * HOWN * 4);
      consumer_sync();
      if (tid < NB * QROWS / 4) {
        const int b = (4 * tid) / QROWS, i = (4 * tid) % QROWS;
        const float4 h0 = *reinterpret_cast<const float4*>(halves + 4 * tid);
        const float4 h1 = *reinterpret_cast<const float4*>(halves + NB * QROWS + 4 * tid);
        const float4 v = make_float4(h0.x + h1.x, h0.y + h1.y, h0.z + h1.z, h0.w + h1.w);
        xpush(sbase + LY::rb + ((rank * NB + b) * HOWN + i % HOWN) * 4, i / HOWN, v);
      }
      xwait();
      if (tid < HOWN * NB) {
        const int b = tid / HOWN, ii = tid % HOWN;
        const float* rr = reinterpret_cast<const float*>(smem + LY::rb);
        float y = 0.f;
#pragma unroll
        for (int s = 0; s < CL; ++s) y += rr[(s * NB + b) * HOWN + ii];
        a.logits[(size_t)b * (a.Kc * a.V) + cl * HROWS + rank * HOWN + ii] = y * (b ? rstd[NB - 1] : rstd[0]);
      }
      stamp();
      // the step's only device-wide barrier: logits complete, every CTA is done reading the residual buffers
      grid_barrier(&a.state->barrier, (epoch * nbar + (++bar_i)) * (unsigned)G);
      stamp();
    }
    // clear the residual buffers of this step for the next one (the last reader is past the barrier)
    {
      constexpr int per_cta = NB * DM / (CL * NCL);
      for (int i = tid; i < 2 * L * per_cta; i += kCT)
        a.xfix[(size_t)NB * DM + (size_t)(i / per_cta) * NB * DM + cta * per_cta + i % per_cta] = 0;
    }

    // ============ CFG / sampling / mask-fix / write-back: one warp per (clip, codebook) ============
    {
      const SampleArgs& sa = *reinterpret_cast<const SampleArgs*>(smem + LY::sargs);
      const int nrows = sa.B * sa.K;
      for (int u = cta + G * warp; u < nrows; u += G * CW) sample_row(sa, u / sa.K, u % sa.K, lane, offset);
    }
    stamp();
    if (cta == 0 && tid == 0) {  // every CTA read offset/epoch before its first barrier arrival
      a.state->offset = offset + 1;
      a.state->epoch = epoch + 1;
    }
  }
  __syncwarp();
  hw_cluster_sync();  // no CTA exits while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t cluster_stream_bytes(int L) {
  return ((size_t)NHEAD * CL * L * SH_SLOTS + (size_t)NCL * CL * ((size_t)L * PRIV_SLOTS + HEAD_SLOTS)) * SLOT;
}

bool cluster_supported(int rows, int L, int D, int F, int H, int head_rows, int page_size, int cond_dim, int max_ctx) {
  if (rows != 1 && rows != 2) return false;
  if (D != DM || F != FF || H != NHEAD || head_rows != NCL * HROWS) return false;
  if (page_size != 32 || L < 1) return false;
  if (cond_dim % 4 || cond_dim <= 0 || cond_dim >= DM) return false;
  if ((max_ctx + CL - 1) / CL > MAXIT * CW) return false;  // attention items per warp
  return true;
}

size_t cluster_xfix_bytes(int rows, int L) { return (size_t)(2 * L + 1) * rows * DM * sizeof(long long); }

// One-time set-up of an instantiation: opt into the large dynamic shared memory window and check that all 32 clusters
// of 4 CTAs can be co-resident (the kernel's device-wide barriers need that). Returns 0 when the device cannot hold
// them (fewer SMs visible than 128, MIG slice, MPS partition), else 1 (cooperative + cluster) or 2 (cluster only).
template <int NB, bool TM>
static int cluster_mode(cudaError_t* err) {
  static int mode = -1;  // -1 = not probed yet
  if (mode >= 0) { *err = mode ? cudaSuccess : cudaErrorCooperativeLaunchTooLarge; return mode; }
  constexpr int smem = Lay<NB>::total;
  cudaError_t e = cudaFuncSetAttribute(decode_step_cluster<NB, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { *err = e; (void)cudaGetLastError(); return mode = 0; }
  cudaLaunchConfig_t qc{};
  qc.gridDim = dim3(CL * NCL); qc.blockDim = dim3(kThreadsC); qc.dynamicSmemBytes = smem;
  cudaLaunchAttribute qa[1];
  qa[0].id = cudaLaunchAttributeClusterDimension;
  qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
  qc.attrs = qa; qc.numAttrs = 1;
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, decode_step_cluster<NB, TM>, &qc);
  if (e != cudaSuccess) { *err = e; (void)cudaGetLastError(); return mode = 0; }
  if (nclusters < NCL) { *err = cudaErrorCooperativeLaunchTooLarge; return mode = 0; }
  // VAURA_CLUSTER_NOCOOP=1: skip the cooperative attribute (Nsight Compute's kernel replay rejects cooperative
  // cluster launches); co-residency is already established by the occupancy query above
  *err = cudaSuccess;
  return mode = knobs().cluster_nocoop ? 2 : 1;
}

template <int NB, bool TM>
static cudaError_t launch_cluster_t(const PersistArgs& a, cudaStream_t st) {
  constexpr int smem = Lay<NB>::total;
  static int mode = 0;
  if (!mode) {
    cudaError_t e;
    mode = cluster_mode<NB, TM>(&e);
    if (!mode) return e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL * NCL); cfg.blockDim = dim3(kThreadsC); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = mode == 1 ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_step_cluster<NB, TM>, a);
  if (e != cudaSuccess && mode == 1) {
    // cooperative + cluster launch rejected by this driver: co-residency is still guaranteed by the occupancy query
    // above (32 clusters of one CTA per SM on an otherwise idle device), so launch with the cluster attribute alone
    (void)cudaGetLastError();
    mode = 2;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, decode_step_cluster<NB, TM>, a);
  }
  return e;
}

bool cluster_launchable(int rows, bool timing) {
  cudaError_t e;
  switch (rows) {
    case 1: return (timing ? cluster_mode<1, true>(&e) : cluster_mode<1, false>(&e)) != 0;
    case 2: return (timing ? cluster_mode<2, true>(&e) : cluster_mode<2, false>(&e)) != 0;
  }
  return false;
}

cudaError_t launch_decode_cluster(const PersistArgs& a, int rows, cudaStream_t st) {
  const bool tm = a.timing != nullptr;
  switch (rows) {
    case 1: return tm ? launch_cluster_t<1, true>(a, st) : launch_cluster_t<1, false>(a, st);
    case 2: return tm ? launch_cluster_t<2, true>(a, st) : launch_cluster_t<2, false>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vaura
