// tcgen05 / TMEM / TMA GEMM for the dense contractions of the path (sm_100a only):
//   D[m][n] = sum_tap sum_k A[m + off(tap)][k] * B[tap][n][k]        (both operands K-major, fp32 accumulate)
// A = activations (rows = sequence rows, prefill positions, or codec time steps), B = weights.
//   - batch >= 16 decode and prefill: QKV / wo / w1|w3 / w2 / heads projections of llama.py:228,:259,:176-177,:504
//     with RoPE+KV-append, residual, SiLU*mul epilogues (bf16 operands)
//   - codec: every WNConv1d / WNConvTranspose1d of the DAC decoder as an implicit GEMM (fp16 operands): the
//     7 taps (or the 2 taps of a polyphase of the transposed conv) are extra K iterations whose A tile is the
//     same TMA box shifted in time; out-of-range rows are zero-filled by TMA, which is the conv padding.
//
// One CTA computes one 128 x BLOCK_N tile:  warp 0 = TMA producer, warp 1 = TMEM alloc + single-thread
// tcgen05.mma issue, warps 2-9 = epilogue (tcgen05.ld 32 lanes each -> registers -> fused epilogue -> global).
// smem ring of STAGES {A 128xBLOCK_K, B BLOCK_NxBLOCK_K} tiles in the 128B (or 64B) swizzled K-major layout
// that TMA writes and the UMMA shared-memory descriptor reads; full/empty mbarriers; accumulator in TMEM.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sampling.cuh"

namespace vaura {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error at the next API call) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 2000000000ull) __trap();  // 2 s
  }
}
// warp-uniform variants (see umma_bf16_f16_elect): every lane executes, one elected lane issues
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_elect(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// plain (1-D) bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds_u4(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Warp-uniform variants: EVERY lane of a converged warp executes the statement with identical operands and one elected
// lane issues.  Inside `if (lane == 0)` the compiler cannot prove uniformity and wraps each tcgen05.mma / bulk copy in
// an ELECT + R2UR.BROADCAST loop (~80 cycles per instruction, measured 41 ns per MMA); with uniform control flow the
// operands are plain uniform registers.
__device__ __forceinline__ void umma_bf16_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// same, the shared-memory descriptors given by their low words (the 14-bit address field); the high word of a
// 128B-swizzled K-major descriptor is a constant
__device__ __forceinline__ void umma_bf16_f16_elect_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "n"(0x40004040)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  // load + wait in ONE asm statement so no consumer of r[] can be scheduled before the wait
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// thread-block cluster helpers (split-K inside a cluster, partial sums exchanged through distributed shared memory)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on the peer's mbarrier `rbar`
__device__ __forceinline__ void st_async_f4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(raddr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar)
               : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major tile whose rows are one swizzle atom wide
// (BLOCK_K * 2 bytes == swizzle bytes): SBO = 8 rows * swizzle bytes, LBO unused, version 1 (sm_100).
template <int SWIZZLE_BYTES>
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2 : (SWIZZLE_BYTES == 64 ? 4 : 6);
  constexpr uint64_t sbo = (8 * SWIZZLE_BYTES) >> 4;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: fp32 accumulate, A/B both K-major, fmt 0 = f16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kTileM = 128;
constexpr int kGemmThreads = 64 + 256;  // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant)

struct TcShape {
  int ntaps, nphase, kblocks;  // K iterations = ntaps * kblocks (per phase)
  int batch;
  int ksplit;                  // split-K factor (linear layers with an accumulating epilogue); grid.z = batch*nphase*ksplit
  int pdl;                     // launched with programmatic stream serialization: weights may be fetched before the
                               // prerequisite grid has finished, everything else waits on griddepcontrol
  int bwrap;                   // > 0: operand B has only this many K blocks; block kb of the K loop reads block kb % bwrap
                               // (A = several bf16 terms of a split fp32 operand side by side, see LinearTcArgs::w_k)
  int n_fastest;               // persistent kernel: consecutive tiles walk N first (CTAs that run at the same time share the
                               // activation rows in L2: the operand that does not fit when M is 10^4..10^6 rows)
  int tap_off[32];             // [nphase][ntaps] row shift of the A box
};

// ------------------------------------------------------------------------------------------------
// epilogues: called once per (row, 16 consecutive columns)
// ------------------------------------------------------------------------------------------------
// Snake1d: x + (alpha + 1e-9)^-1 * sin(alpha x)^2 with the reciprocal precomputed per channel (the reference's own
// form) and the SFU sine: |alpha x| stays below ~1e2 in this decoder, where __sinf's absolute error (~1e-5) is far
// under the fp16 rounding of the stored activation.
__device__ __forceinline__ float snake_eval(float v, float alpha, float inv_alpha) {
  const float s = __sinf(alpha * v);
  return fmaf(s * s, inv_alpha, v);
}

struct EpiConv {
  struct Params {
    const float* bias;
    const float* alpha;
    const __half* residual;
    __half* out_raw;
    __half* out_act;
    int Tq, Tout, Cout, ostride;
  };
  // the 16 residual values of a chunk (zeros when the layer has none): the persistent kernel requests them before it
  // waits for the accumulator, so that their latency overlaps the tile's MMAs
  __device__ static void load_residual(const Params& p, int b, int phase, int m, int n0, uint4& r0, uint4& r1) {
    r0 = r1 = make_uint4(0u, 0u, 0u, 0u);
    if (!p.residual || m >= p.Tq || n0 >= p.Cout) return;
    const size_t base = ((size_t)b * p.Tout + (size_t)m * p.ostride + phase) * p.Cout + n0;
    r0 = *reinterpret_cast<const uint4*>(p.residual + base);
    r1 = *reinterpret_cast<const uint4*>(p.residual + base + 8);
  }
  __device__ static void apply(const Params& p, int b, int phase, int m, int n0, float (&v)[16]) {
    uint4 r0, r1;
    load_residual(p, b, phase, m, n0, r0, r1);
    apply(p, b, phase, m, n0, v, r0, r1);
  }
  __device__ static void apply(const Params& p, int b, int phase, int m, int n0, float (&v)[16], const uint4& r0, const uint4& r1) {
    if (m >= p.Tq || n0 >= p.Cout) return;
    const size_t base = ((size_t)b * p.Tout + (size_t)m * p.ostride + phase) * p.Cout + n0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + i);
      v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    }
    if (p.residual) {
      const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
      const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] += __low2float(h0[i]); v[2 * i + 1] += __high2float(h0[i]);
        v[8 + 2 * i] += __low2float(h1[i]); v[8 + 2 * i + 1] += __high2float(h1[i]);
      }
    }
    if (p.out_raw) {
      uint4 o[2];
      __half2* h = reinterpret_cast<__half2*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(p.out_raw + base) = o[0];
      *reinterpret_cast<uint4*>(p.out_raw + base + 8) = o[1];
    }
    if (p.out_act) {
      uint4 o[2];
      __half2* h = reinterpret_cast<__half2*>(o);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 al = *reinterpret_cast<const float4*>(p.alpha + n0 + i);
        const float4 ia = *reinterpret_cast<const float4*>(p.alpha + p.Cout + n0 + i);  // [C] alpha | [C] 1/(alpha+1e-9)
        h[i / 2] = __floats2half2_rn(snake_eval(v[i], al.x, ia.x), snake_eval(v[i + 1], al.y, ia.y));
        h[i / 2 + 1] = __floats2half2_rn(snake_eval(v[i + 2], al.z, ia.z), snake_eval(v[i + 3], al.w, ia.w));
      }
      *reinterpret_cast<uint4*>(p.out_act + base) = o[0];
      *reinterpret_cast<uint4*>(p.out_act + base + 8) = o[1];
    }
  }
};

// ---- linear-layer epilogues of the bf16 sampler path (row = sequence row * npos + position) ---------------
struct EpiLinear {
  struct Params {
    int mode;             // EPI_STORE / EPI_RESID / EPI_SWIGLU / EPI_QKV
    int R, N;             // valid rows / output features
    float* out_f32;       // STORE: [R][ldo] (or permuted), RESID: residual stream h [R][ldo]
    __nv_bfloat16* out_bf16;  // SWIGLU: act [R][ldo]; QKV: q [R][d_model]
    int ldo;
    int perm_S, perm_V;   // STORE: see GemvArgs
    const float* rope;
    KvView kv;
    const StepState* state;
    int pos0, npos, layer, d_model;
    int atomic;
    int aux;              // EPI_SWIGLU_SPLIT3: F
  };
  // what the RoPE / KV-append epilogue of one 16-column chunk reads from global memory besides the accumulator: eight
  // (cos, sin) pairs and the destination row in the paged cache.  The fused step kernel requests them before it waits for
  // the accumulator (their L2 latency overlaps the tile's MMAs).
  struct QkvPre {
    float4 cs[4];
    size_t dst;  // element offset into q (section 0) or the K/V pages (sections 1, 2)
  };
  __device__ static void prefetch_qkv(const Params& p, int m, int n0, QkvPre& q) {
    const int D = p.d_model, sec = n0 / D, within = n0 % D;
    const int hd = within / kHeadDim, e = within % kHeadDim;
    const int b = m / p.npos, j = m % p.npos;
    const int pos = (p.state ? p.state->offset - p.npos : p.pos0) + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) q.cs[i] = make_float4(1.f, 0.f, 1.f, 0.f);
    q.dst = 0;
    if (m >= p.R || n0 >= p.N) return;
    if (sec != 2) {
      const float4* cs = reinterpret_cast<const float4*>(p.rope + ((size_t)pos * (kHeadDim / 2) + (e >> 1)) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) q.cs[i] = cs[i];
    }
    q.dst = sec == 0 ? (size_t)m * D + within : p.kv.row(p.layer, sec - 1, b, pos, hd) + e;
  }
  __device__ static void apply_qkv(const Params& p, int m, int n0, float (&v)[16], const QkvPre& q) {
    if (m >= p.R || n0 >= p.N) return;
    const int sec = n0 / p.d_model;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // two (cos,sin) pairs per float4; identity for the V section
      const float4 c = q.cs[i];
      const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
      v[4 * i] = x0 * c.x - x1 * c.y; v[4 * i + 1] = x1 * c.x + x0 * c.y;
      v[4 * i + 2] = x2 * c.z - x3 * c.w; v[4 * i + 3] = x3 * c.z + x2 * c.w;
    }
    uint4 o[2];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    __nv_bfloat16* dst = (sec == 0 ? p.out_bf16 : reinterpret_cast<__nv_bfloat16*>(p.kv.pages)) + q.dst;
    *reinterpret_cast<uint4*>(dst) = o[0];
    *reinterpret_cast<uint4*>(dst + 8) = o[1];
  }
  __device__ static void apply(const Params& p, int /*b*/, int /*phase*/, int m, int n0, float (&v)[16]) {
    if (m >= p.R || n0 >= p.N) return;
    if (p.mode == EPI_STORE) {
      size_t o = (size_t)m * p.ldo + n0;
      if (p.perm_S > 0) {
        const int bb = m / p.perm_S, j = m % p.perm_S, kk = n0 / p.perm_V;
        o = (((size_t)bb * (p.N / p.perm_V) + kk) * p.perm_S + j) * p.perm_V + (n0 % p.perm_V);
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p.out_f32 + o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if (p.mode == EPI_RESID) {
      float* o = p.out_f32 + (size_t)m * p.ldo + n0;
      if (p.atomic) {  // split-K partial: accumulate into the fp32 residual stream with vector reductions
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                       "f"(v[i + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 h = *reinterpret_cast<float4*>(o + i);
          h.x += v[i]; h.y += v[i + 1]; h.z += v[i + 2]; h.w += v[i + 3];
          *reinterpret_cast<float4*>(o + i) = h;
        }
      }
    } else if (p.mode == EPI_SWIGLU) {
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a0 = v[4 * i], b0 = v[4 * i + 1], a1 = v[4 * i + 2], b1 = v[4 * i + 3];
        // SiLU(a) * b; fast division / exponential: the result is rounded to bf16 (the IEEE division's out-of-line slow
        // path is a branch per quotient, which serialises the eight of a chunk)
        h[i] = __floats2bfloat162_rn(__fdividef(a0, 1.f + __expf(-a0)) * b0, __fdividef(a1, 1.f + __expf(-a1)) * b1);
      }
      *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)m * p.ldo + (n0 >> 1)) = o;
    } else if (p.mode == EPI_SWIGLU_SPLIT3) {
      // fp32-activation prefill: exact SiLU (same expression as gemv_kernel<EPI_SWIGLU>), the 8 results of the chunk stored as
      // three bf16 terms at [m][j], [m][F + j], [m][2F + j]
      __nv_bfloat16 t[3][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = v[2 * i], b0 = v[2 * i + 1];
        split3(a0 / (1.f + expf(-a0)) * b0, t[0][i], t[1][i], t[2][i]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k)
        *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)m * p.ldo + (size_t)k * p.aux + (n0 >> 1)) = *reinterpret_cast<const uint4*>(t[k]);
    } else if (p.mode == EPI_QKV_F32) {
      // RoPE (llama.py:633-650), q -> fp32 [R][d], K/V -> fp32 pages (what gemv_kernel<EPI_QKV> writes)
      const int D = p.d_model, sec = n0 / D, within = n0 % D;
      const int hd = within / kHeadDim, e = within % kHeadDim;
      const int b = m / p.npos, j = m % p.npos;
      const int pos = (p.state ? p.state->offset - p.npos : p.pos0) + j;
      if (sec != 2) {
        const float4* cs = reinterpret_cast<const float4*>(p.rope + ((size_t)pos * (kHeadDim / 2) + (e >> 1)) * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 c = cs[i];
          const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
          v[4 * i] = x0 * c.x - x1 * c.y; v[4 * i + 1] = x1 * c.x + x0 * c.y;
          v[4 * i + 2] = x2 * c.z - x3 * c.w; v[4 * i + 3] = x3 * c.z + x2 * c.w;
        }
      }
      float* dst = sec == 0 ? p.out_f32 + (size_t)m * D + within
                            : reinterpret_cast<float*>(p.kv.pages) + p.kv.row(p.layer, sec - 1, b, pos, hd) + e;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {  // EPI_QKV
      const int D = p.d_model, sec = n0 / D, within = n0 % D;
      const int hd = within / kHeadDim, e = within % kHeadDim;
      const int b = m / p.npos, j = m % p.npos;
      const int pos = (p.state ? p.state->offset - p.npos : p.pos0) + j;
      if (sec != 2) {
        const float4* cs = reinterpret_cast<const float4*>(p.rope + ((size_t)pos * (kHeadDim / 2) + (e >> 1)) * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // two (cos,sin) pairs per float4
          const float4 c = cs[i];
          const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
          v[4 * i] = x0 * c.x - x1 * c.y; v[4 * i + 1] = x1 * c.x + x0 * c.y;
          v[4 * i + 2] = x2 * c.z - x3 * c.w; v[4 * i + 3] = x3 * c.z + x2 * c.w;
        }
      }
      uint4 o[2];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      __nv_bfloat16* dst = sec == 0 ? p.out_bf16 + (size_t)m * D + within
                                    : reinterpret_cast<__nv_bfloat16*>(p.kv.pages) + p.kv.row(p.layer, sec - 1, b, pos, hd) + e;
      *reinterpret_cast<uint4*>(dst) = o[0];
      *reinterpret_cast<uint4*>(dst + 8) = o[1];
    }
  }
};

// ---- linear-layer epilogues of the Segment-AVCLIP visual tower (avclip.cu; motionformer_src/vit_helper.py) -------------
struct EpiVit {
  struct Params {
    int mode;                  // VIT_STORE_BF16 / VIT_RESID_F32 / VIT_PATCH (kernels.h)
    int gelu;                  // VIT_STORE_BF16: exact (erf) GELU after the bias (nn.GELU(), vit_helper.py:486)
    int M, N, ldo;
    const float* bias;         // [N]
    __nv_bfloat16* out_bf16;   // VIT_STORE_BF16: [M][ldo]
    float* out_f32;            // VIT_RESID_F32: residual stream, updated in place [M][ldo]; VIT_PATCH: token rows
    const float* pos;          // VIT_PATCH: [rows_in][N] position table added to the embedded tubelets
    int rows_in, rows_out, row_off;  // VIT_PATCH: GEMM row m -> token row (m / rows_in) * rows_out + row_off + m % rows_in
  };
  __device__ static void load_residual(const Params&, int, int, int, int, uint4&, uint4&) {}
  __device__ static void apply(const Params& p, int b, int phase, int m, int n0, float (&v)[16], const uint4&, const uint4&) {
    apply(p, b, phase, m, n0, v);
  }
  __device__ static void apply(const Params& p, int /*b*/, int /*phase*/, int m, int n0, float (&v)[16]) {
    if (m >= p.M || n0 >= p.N) return;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + i);
      v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    }
    if (p.mode == VIT_STORE_BF16) {
      if (p.gelu) {
        // exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the
        // bf16 rounding of the stored value): one reciprocal, one exponential and five FMAs instead of erff's ~25 instructions -
        // the fc1 epilogue (128 values per thread and tile) was longer than the tile's MMAs (fc1 681 TFLOP/s vs 935 for q|k|v)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x = v[i], z = fabsf(x) * 0.70710678118654752f;
          const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
          const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
          const float e = 1.f - poly * __expf(-z * z);       // erf(|x| / sqrt 2)
          v[i] = 0.5f * x * (1.f + copysignf(e, x));
        }
      }
      uint4 o[2];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      __nv_bfloat16* dst = p.out_bf16 + (size_t)m * p.ldo + n0;
      *reinterpret_cast<uint4*>(dst) = o[0];
      *reinterpret_cast<uint4*>(dst + 8) = o[1];
    } else if (p.mode == VIT_RESID_F32) {
      float* o = p.out_f32 + (size_t)m * p.ldo + n0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 h = *reinterpret_cast<float4*>(o + i);
        h.x += v[i]; h.y += v[i + 1]; h.z += v[i + 2]; h.w += v[i + 3];
        *reinterpret_cast<float4*>(o + i) = h;
      }
    } else {  // VIT_PATCH
      const int seg = m / p.rows_in, tok = m % p.rows_in;
      const float* ps = p.pos + (size_t)tok * p.N + n0;
      float* o = p.out_f32 + ((size_t)seg * p.rows_out + p.row_off + tok) * p.ldo + n0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(ps + i);
        *reinterpret_cast<float4*>(o + i) = make_float4(v[i] + q.x, v[i + 1] + q.y, v[i + 2] + q.z, v[i + 3] + q.w);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// KSUB: 64-wide K blocks per pipeline stage.  The single-thread producer / MMA loops pay ~0.25 us of serialized
// mbarrier-wait + fence + commit latency per stage; with thin tiles (decode: 64 rows x 32-64 weight rows) that
// latency, not bandwidth, bounded the mainloop, so linear layers put 4 K blocks (16 UMMAs) behind each barrier.
// TERMS > 1 (fp32-equivalent GEMMs of the fp32-activation prefill): operand A holds TERMS bf16 terms of a split fp32
// matrix side by side ([R][TERMS * Kw], term t of K block kb at column (t * g.bwrap + kb) * 64); a stage carries the TERMS A
// tiles of one K block and ONE B tile, multiplied TERMS times - the weight tile is fetched once per K block, not per term.
// CK > 1: the CK CTAs of a cluster (1, 1, CK) share one output tile and split its K range (g.ksplit == CK); after the mainloop
// CTA r finalises columns [r, r + 1) * BLOCK_N / CK: the others send it their partial accumulators with st.async (data +
// transaction-count completion on the owner's mbarrier) into the idle stage ring, the owner adds the CK partials in rank order
// (deterministic) and runs the epilogue on complete sums - every element keeps one owner and there are no float atomics.
template <int BLOCK_N, int BLOCK_K, int STAGES, int FMT, class Epi, int TILE_M = 128, int KSUB = 1, int TERMS = 1, int CK = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcShape g,
               const typename Epi::Params ep) {
  constexpr int SW = BLOCK_K * 2;
  static_assert(TILE_M == 64 || TILE_M == 128, "UMMA M for cta_group::1");
  constexpr int A_BYTES = TILE_M * BLOCK_K * 2, B_BYTES = BLOCK_N * BLOCK_K * 2, SUB_BYTES = TERMS * A_BYTES + B_BYTES;
  constexpr int AT_BYTES = TERMS * A_BYTES;  // offset of the B tile inside a sub-stage
  constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
  constexpr int TMEM_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage buffers must stay 1024B aligned for the swizzle");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N");
  static_assert(TERMS == 1 || KSUB == 1, "split operands use one K block per stage");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* xbar = tmem_full + 1;  // CK > 1: completes when the peers' partial sums of this CTA's columns have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 1);
  static_assert(CK == 1 || (TILE_M == 128 && (BLOCK_N / CK) % 16 == 0 && BLOCK_N % CK == 0), "cluster split-K tile shape");
  static_assert(CK == 1 || (size_t)(CK - 1) * 128 * (BLOCK_N / CK + 4) * 4 <= (size_t)STAGES * STAGE_BYTES,
                "the receive buffer aliases the stage ring");

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int m0 = blockIdx.x * TILE_M, n0 = blockIdx.y * BLOCK_N;
  const int split = blockIdx.z % g.ksplit, zz = blockIdx.z / g.ksplit;
  const int b = zz / g.nphase, phase = zz % g.nphase;
  // units = (tap, K block) pairs; a stage covers up to KSUB consecutive units (KSUB > 1 only with a single tap)
  const int total_units = g.ntaps * g.kblocks;
  const int u_begin = (total_units * split) / g.ksplit, u_end = (total_units * (split + 1)) / g.ksplit;
  const int iters = (u_end - u_begin + KSUB - 1) / KSUB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    if (CK > 1) {
      mbar_init(xbar, 1);
      mbar_expect_tx(xbar, (uint32_t)((CK - 1) * 128 * (BLOCK_N / CK) * 4));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (CK > 1) cluster_sync_all();  // every peer's exchange barrier is initialised and armed before anything can reach it

  // NOTE: triggering the dependent grid BEFORE this grid's own griddepcontrol.wait was measured to break the chain
  // (the dependent's wait then no longer covers our prerequisite); the trigger is issued after the wait, below.

  // warps 0 and 1 walk their loops with all 32 lanes (uniform control flow) and one elected lane issues: inside
  // `if (lane == 0)` every tensor copy / MMA costs an ELECT + R2UR loop (~0.13 us per copy, ~41 ns per MMA; measured
  // 0.54 us per 64-wide K block on the prefill GEMMs before this change)
  if (warp == 0) {
    // weights (operand B) do not depend on the previous kernel: fill the first ring pass before waiting for it
    const int pre = (g.pdl & 4) ? min(iters, STAGES) : 0;
    for (int it = 0; it < pre; ++it) {
      const int u0 = u_begin + it * KSUB, nsub = min(KSUB, u_end - u0);
      mbar_expect_tx_elect(&full[it], nsub * SUB_BYTES);
      for (int sub = 0; sub < nsub; ++sub) {
        const int gi = u0 + sub, tap = gi / g.kblocks, kb = gi % g.kblocks;
        const int kbB = (TERMS == 1 && g.bwrap) ? kb % g.bwrap : kb;
        tma_load_3d_elect(smem + it * STAGE_BYTES + sub * SUB_BYTES + AT_BYTES, &tmB, &full[it], kbB * BLOCK_K, n0, phase * g.ntaps + tap);
      }
    }
    if (g.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      const int u0 = u_begin + it * KSUB, nsub = min(KSUB, u_end - u0);
      uint8_t* ss = smem + s * STAGE_BYTES;
      if (it >= pre) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx_elect(&full[s], nsub * SUB_BYTES);
      }
      for (int sub = 0; sub < nsub; ++sub) {
        const int gi = u0 + sub, tap = gi / g.kblocks, kb = gi % g.kblocks;
        const int kbB = (TERMS == 1 && g.bwrap) ? kb % g.bwrap : kb;
        uint8_t* sa = ss + sub * SUB_BYTES;
        if (it >= pre) tma_load_3d_elect(sa + AT_BYTES, &tmB, &full[s], kbB * BLOCK_K, n0, phase * g.ntaps + tap);
#pragma unroll
        for (int t = 0; t < TERMS; ++t)
          tma_load_3d_elect(sa + t * A_BYTES, &tmA, &full[s], (t * g.bwrap + kb) * BLOCK_K, m0 + g.tap_off[phase * g.ntaps + tap], b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(TILE_M, BLOCK_N, FMT);
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      const int nsub = min(KSUB, u_end - (u_begin + it * KSUB));
      for (int sub = 0; sub < nsub; ++sub) {
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES + sub * SUB_BYTES);
        const uint64_t bdesc = make_smem_desc<SW>(a_addr + AT_BYTES);
#pragma unroll
        for (int t = 0; t < TERMS; ++t) {
          const uint64_t adesc = make_smem_desc<SW>(a_addr + t * A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)  // +32 bytes (>>4 = 2) per UMMA_K inside the swizzle atom
            umma_bf16_f16_elect(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | sub | t | k) != 0);
        }
      }
      umma_commit_elect(&empty[s]);  // frees the smem slot once these MMAs have read it
    }
    umma_commit_elect(tmem_full);    // accumulator complete
    __syncwarp();
  } else {
    if (g.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // our prerequisite has completed: the dependent grid may now start its prologue and weight prefetch
    if (g.pdl & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
  }
  if constexpr (CK > 1) {
    // all MMAs of all CTAs of the cluster have completed (the epilogue warps arrive after their accumulator wait): the stage
    // rings are idle and may receive partial sums
    cluster_sync_all();
    if (warp >= 2) {
      constexpr int OWN = BLOCK_N / CK, RST = OWN + 4;  // columns finalised per CTA; receive row stride (floats, conflict-free)
      float* recv = reinterpret_cast<float*>(smem);     // [CK - 1 senders in rank order][128 rows][RST]
      const uint32_t my = cluster_ctarank();
      const int q = warp & 3, row = q * 32 + lane, m = m0 + row;
      constexpr int kChunks = BLOCK_N / 16, kHalf = (kChunks + 1) / 2;
      const int c_begin = (warp - 2) < 4 ? 0 : kHalf * 16, c_end = (warp - 2) < 4 ? kHalf * 16 : BLOCK_N;
      const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 16) {
        const uint32_t owner = (uint32_t)(c / OWN);
        if (owner == my) continue;
        float v[16];
        tmem_ld16(tq + c, v);
        const uint32_t slot = my < owner ? my : my - 1;
        const uint32_t ra = mapa_u32(smem_u32(recv + ((size_t)slot * 128 + row) * RST + (c % OWN)), owner);
        const uint32_t rb = mapa_u32(smem_u32(xbar), owner);
#pragma unroll
        for (int i = 0; i < 16; i += 4) st_async_f4(ra + 4 * i, v[i], v[i + 1], v[i + 2], v[i + 3], rb);
      }
      mbar_wait(xbar, 0);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 16) {
        if ((uint32_t)(c / OWN) != my) continue;
        float own[16], v[16];
        tmem_ld16(tq + c, own);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)CK; ++r) {  // rank order: the sum does not depend on arrival order
          if (r == my) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += own[i];
          } else {
            const float* src = recv + ((size_t)(r < my ? r : r - 1) * 128 + row) * RST + (c % OWN);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const uint4 o = lds_u4(src + i);
              v[i] += __uint_as_float(o.x); v[i + 1] += __uint_as_float(o.y);
              v[i + 2] += __uint_as_float(o.z); v[i + 3] += __uint_as_float(o.w);
            }
          }
        }
        Epi::apply(ep, b, phase, m, n0 + c, v);
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    // M=128: accumulator row i sits in lane i.  M=64: rows 16q..16q+15 sit in lanes 32q..32q+15 (half-filled quadrants)
    const int m = TILE_M == 128 ? m0 + q * 32 + lane : m0 + q * 16 + lane;
    const bool row_ok = TILE_M == 128 || lane < 16;
    // the two warps of a quadrant split the accumulator columns
    constexpr int kChunks = BLOCK_N / 16, kHalf = (kChunks + 1) / 2;
    const int c_begin = (warp - 2) < 4 ? 0 : kHalf * 16, c_end = (warp - 2) < 4 ? kHalf * 16 : BLOCK_N;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      if (row_ok) Epi::apply(ep, b, phase, m, n0 + c, v);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// persistent variant (codec convs): one CTA per SM walks the (m, n, batch*phase) tiles; the accumulator is double
// buffered in TMEM so the epilogue of tile i (bias + residual + Snake + two fp16 stores per element, the long pole of
// the narrow layers) overlaps the TMA / MMA mainloop of tile i+1, and barrier init / TMEM allocation / launch happen
// once per layer instead of once per 128-row tile.  Same tile arithmetic as gemm_tc_kernel (KSUB = 1, no split-K).
// ------------------------------------------------------------------------------------------------
// KSUB > 1: a stage holds KSUB K blocks of each operand (KSUB activation sub-tiles, then KSUB weight sub-tiles), loaded by
// ONE 4-D tensor box per operand (maps from make_map_kblocks) - for channel counts that force a narrow K block (Cin = 96:
// 32-wide blocks) the copies and barrier round trips per tap drop from 3 to 1.
// OCC = 2: two CTAs per SM (half the ring each, 2 x 2 accumulator buffers of <= 128 TMEM columns): for the narrow layers whose
// tiles are paced by the serial work of one producer / MMA / epilogue chain rather than by bytes in flight.
// EW: epilogue warps (8 or 16 = two or four per TMEM lane quadrant, splitting the accumulator columns).  With a short K (the
// tower's K = 768: 3.2 us of MMAs per 128 x 256 tile) the tile rate is set by the epilogue, whose cost per warp is instruction
// and store-transaction bound; sixteen warps halve it (tower: 186.8 -> 179.7 ms per 256 segments; the codec's convolutions
// measured the same with 8 and 16).  What then bounds these GEMMs is the operand ingest: 590 KB per tile at the ~75 GB/s per SM
// that four 48 KB stages in flight sustain = 7.8 us per tile = ~950 TFLOP/s; only tiles that share an operand between two SMs
// (cta_group::2) would lower the bytes per FLOP.
template <int BLOCK_N, int BLOCK_K, int STAGES, int FMT, class Epi, int KSUB = 1, int OCC = 1, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW, OCC)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcShape g,
                          const typename Epi::Params ep, const int m_tiles, const int n_tiles) {
  constexpr int SW = BLOCK_K * 2;
  constexpr int TILE_M = kTileM;
  constexpr int A_BYTES = TILE_M * BLOCK_K * 2, B_BYTES = BLOCK_N * BLOCK_K * 2, STAGE_BYTES = KSUB * (A_BYTES + B_BYTES);
  constexpr int ACC_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;  // per accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(OCC * TMEM_COLS <= 512, "the co-resident CTAs share the SM's 512 TMEM columns");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage buffers must stay 1024B aligned for the swizzle");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2] accumulator buffer written
  uint64_t* tempty = tfull + 2;       // [2] accumulator buffer drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int kgroups = g.kblocks / KSUB;  // stages per tap
  const int iters = g.ntaps * kgroups;
  const int zdim = g.batch * g.nphase;
  const int ntiles = m_tiles * n_tiles * zdim;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // tile -> (m0, n0, batch, phase); m fastest so that neighbouring CTAs share the weight tile in L2
  auto tile_coords = [&](int t, int& m0, int& n0, int& b, int& phase) {
    int zz;
    if (g.n_fastest) {
      n0 = (t % n_tiles) * BLOCK_N;
      const int r = t / n_tiles;
      m0 = (r % m_tiles) * TILE_M;
      zz = r / m_tiles;
    } else {
      m0 = (t % m_tiles) * TILE_M;
      const int r = t / m_tiles;
      n0 = (r % n_tiles) * BLOCK_N;
      zz = r / n_tiles;
    }
    b = zz / g.nphase;
    phase = zz % g.nphase;
  };

  // warps 0 and 1 run their loops with all 32 lanes (uniform control flow) and one elected lane issues: inside
  // `if (lane == 0)` every tensor copy / MMA costs an ELECT + R2UR loop (~0.13 us per copy, ~41 ns per MMA)
  if (warp == 0) {
    int git = 0;  // ring iterations since the kernel started
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, b, phase;
      tile_coords(t, m0, n0, b, phase);
      for (int it = 0; it < iters; ++it, ++git) {
        const int s = git % STAGES;
        const uint32_t ph = (git / STAGES) & 1;
        const int tap = it / kgroups, kb = it % kgroups;
        uint8_t* ss = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx_elect(&full[s], STAGE_BYTES);
        if (KSUB == 1) {
          tma_load_3d_elect(ss + A_BYTES, &tmB, &full[s], kb * BLOCK_K, n0, phase * g.ntaps + tap);
          tma_load_3d_elect(ss, &tmA, &full[s], kb * BLOCK_K, m0 + g.tap_off[phase * g.ntaps + tap], b);
        } else {
          tma_load_4d_elect(ss + KSUB * A_BYTES, &tmB, &full[s], 0, n0, kb * KSUB, phase * g.ntaps + tap);
          tma_load_4d_elect(ss, &tmA, &full[s], 0, m0 + g.tap_off[phase * g.ntaps + tap], kb * KSUB, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(TILE_M, BLOCK_N, FMT);
    int git = 0, tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      const int ab = tc & 1;
      mbar_wait(&tempty[ab], ((tc >> 1) & 1) ^ 1);  // the epilogue has drained this buffer (first use passes)
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ab * ACC_COLS;
      for (int it = 0; it < iters; ++it, ++git) {
        const int s = git % STAGES;
        const uint32_t ph = (git / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int sub = 0; sub < KSUB; ++sub) {
          const uint64_t adesc = make_smem_desc<SW>(a_addr + sub * A_BYTES),
                         bdesc = make_smem_desc<SW>(a_addr + KSUB * A_BYTES + sub * B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_bf16_f16_elect(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | sub | k) != 0);
        }
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&tfull[ab]);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    // the EW / 4 warps of a quadrant split the accumulator columns into equal runs of 16-column chunks
    constexpr int kParts = EW / 4, kChunks = BLOCK_N / 16, kHalf = (kChunks + kParts - 1) / kParts;
    static_assert(EW == 8 || EW == 16, "epilogue warps");
    const int part = (warp - 2) >> 2;
    const int c_begin = part * kHalf * 16, c_end = (part + 1) * kHalf * 16 < BLOCK_N ? (part + 1) * kHalf * 16 : BLOCK_N;
    int tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      int m0, n0, b, phase;
      tile_coords(t, m0, n0, b, phase);
      const int ab = tc & 1;
      const int m = m0 + q * 32 + lane;
      // residual rows of this warp's chunks: in flight while the tile's MMAs run
      uint4 res[kHalf][2];
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) Epi::load_residual(ep, b, phase, m, n0 + c, res[i][0], res[i][1]);
      }
      mbar_wait(&tfull[ab], (tc >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) {
          float v[16];
          tmem_ld16(tacc + c, v);
          Epi::apply(ep, b, phase, m, n0 + c, v, res[i][0], res[i][1]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[ab])) : "memory");
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent GEMM (cta_group::2): the two CTAs of a cluster (2, 1, 1) compute one 256 x BLOCK_N tile with
// UMMA M = 256 - CTA r holds rows [128 r, 128 r + 128) of A, rows [BLOCK_N / 2 * r, ...) of W and its own 128 x BLOCK_N
// accumulator in TMEM; the leader CTA's MMA warp issues tcgen05.mma.cta_group::2, which reads both CTAs' shared memory.
// Per 128 x 256 outputs a CTA ingests (128 + 128) x 64 x 2 B = 32 KB per K block instead of 48 KB: the single-CTA kernel is
// bound by operand ingest (~75 GB/s per SM = bytes in flight / latency), so the tile rate rises by up to 1.5 x.
// Protocol (CUTLASS sm100 2-SM pipeline): both producers load their halves with cp.async.bulk.tensor...cta_group::2 signalling
// the LEADER's full barrier (peer bit of the barrier address cleared); the leader's producer arms it with the bytes of both;
// tcgen05.commit.cta_group::2 ... multicast::cluster frees the stage in both CTAs and publishes the accumulator to both
// epilogues; the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier (remote mbarrier.arrive).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d_2sm_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc_elect(uint64_t* bar) {  // arrives on `bar` of both CTAs of the pair
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

template <int BLOCK_N, int STAGES, int FMT, class Epi, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tc2_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcShape g,
                           const typename Epi::Params ep, const int m_tiles, const int n_tiles) {
  constexpr int BLOCK_K = 64, SW = 128, HALF_N = BLOCK_N / 2;
  constexpr int A_BYTES = kTileM * BLOCK_K * 2, B_BYTES = HALF_N * BLOCK_K * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_COLS = BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256, TMEM_COLS = 2 * ACC_COLS;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256 && B_BYTES % 1024 == 0, "tile shape");
  static_assert(EW == 8 || EW == 16, "epilogue warps");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // used in the leader CTA only
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2] accumulator buffer complete (both CTAs)
  uint64_t* tempty = tfull + 2;       // [2] leader only: drained by the 2 x EW epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int iters = g.kblocks;
  const int ntiles = m_tiles * n_tiles;  // tiles of 256 rows x BLOCK_N columns

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // both CTAs' barriers exist before anything can signal them; both are resident before the paired alloc
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto tile_coords = [&](int t, int& m0, int& n0) {  // N first: the pairs that run together share activation rows in L2
    n0 = (t % n_tiles) * BLOCK_N;
    m0 = (t / n_tiles) * (2 * kTileM);
  };

  if (warp == 0) {
    int git = 0;
    for (int t = pair; t < ntiles; t += npairs) {
      int m0, n0;
      tile_coords(t, m0, n0);
      for (int it = 0; it < iters; ++it, ++git) {
        const int s = git % STAGES;
        uint8_t* ss = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ((git / STAGES) & 1) ^ 1);
        if (rank == 0) mbar_expect_tx_elect(&full[s], 2 * STAGE_BYTES);
        tma_load_3d_2sm_elect(ss + A_BYTES, &tmB, &full[s], it * BLOCK_K, n0 + (int)rank * HALF_N, 0);
        tma_load_3d_2sm_elect(ss, &tmA, &full[s], it * BLOCK_K, m0 + (int)rank * kTileM, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * kTileM, BLOCK_N, FMT);
      int git = 0, tc = 0;
      for (int t = pair; t < ntiles; t += npairs, ++tc) {
        const int ab = tc & 1;
        mbar_wait(&tempty[ab], ((tc >> 1) & 1) ^ 1);  // both epilogues have drained this buffer (first use passes)
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ab * ACC_COLS;
        for (int it = 0; it < iters; ++it, ++git) {
          const int s = git % STAGES;
          mbar_wait(&full[s], (git / STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc<SW>(a_addr), bdesc = make_smem_desc<SW>(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) umma2_f16_elect(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          umma2_commit_mc_elect(&empty[s]);
        }
        umma2_commit_mc_elect(&tfull[ab]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    constexpr int kParts = EW / 4, kChunks = BLOCK_N / 16, kHalf = (kChunks + kParts - 1) / kParts;
    const int part = (warp - 2) >> 2;
    const int c_begin = part * kHalf * 16, c_end = (part + 1) * kHalf * 16 < BLOCK_N ? (part + 1) * kHalf * 16 : BLOCK_N;
    const uint32_t rempty = mapa_u32(smem_u32(&tempty[0]), 0);  // the leader's accumulator-empty barriers
    int tc = 0;
    for (int t = pair; t < ntiles; t += npairs, ++tc) {
      int m0, n0;
      tile_coords(t, m0, n0);
      const int ab = tc & 1;
      const int m = m0 + (int)rank * kTileM + q * 32 + lane;
      mbar_wait(&tfull[ab], (tc >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) {
          float v[16];
          tmem_ld16(tacc + c, v);
          Epi::apply(ep, 0, 0, m, n0 + c, v);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rempty + 8u * (uint32_t)ab) : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees its half of the paired allocation while the other may still use its own
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// Fused ResidualUnit of the DAC decoder / encoder (dac 1.0.0 ResidualUnit: Snake -> conv k7 dilated -> Snake -> conv k1, + x)
// for C <= 256 channels, where the stand-alone k1 convolution is pure memory traffic (reads h and x, writes x and
// Snake(x): 49 % of the HBM peak at 192 channels, r02_codec_launches_b16.csv) and the k7 convolution writes h only to have
// it read back.  Per 128-row tile:  acc1 = conv7(act) (7 taps x C / BK K blocks from the ring)  ->  the epilogue warps turn
// acc1 into h = Snake(acc1 + b7) in fp16 and store it into shared memory in the UMMA K-major swizzled layout  ->
// acc2 = h x W1^T (W1 K blocks streamed through the same ring)  ->  epilogue: + b1 + x, raw and Snake'd outputs.
// h never leaves the SM: the unit moves 4 activation passes instead of 6 and one launch instead of two.
// TMEM: acc1 at column 0, acc2 at column ACC (two accumulators of C <= 256 columns).  The k7 mainloop of tile i + 1 overlaps
// the second epilogue of tile i; the first epilogue (acc1 -> h) is on the MMA warp's critical path, so it hands h over K block
// by K block (the k1 GEMM starts on block 0 while the others are produced) and reads its bias / Snake parameters from shared
// memory.  A variant skewed by one tile (k1 of tile i - 1 issued behind the k7 mainloop of tile i, acc2 written over acc1)
// measured slower (9.23 vs 8.45 ms per 16 clips): with two 256-column buffers the second epilogue then sits between two k7
// mainloops instead of under one.
// ------------------------------------------------------------------------------------------------
#ifdef VAURA_RU_TIMING
// debug build only (profiles/ru_timing.py): %globaltimer stamps of CTA 0, tiles 8..15: [tile][role 0 = MMA warp, 1 = epilogue warp 2,
// 2 = producer][8 events]
__device__ unsigned long long g_ru_timing[8 * 3 * 8];
#define RU_STAMP(tile_i, role, ev)                                                                     \
  do {                                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && (tile_i) >= 8 && (tile_i) < 16) {                              \
      unsigned long long t_;                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
      g_ru_timing[(((tile_i) - 8) * 3 + (role)) * 8 + (ev)] = t_;                                      \
    }                                                                                                  \
  } while (0)
#else
#define RU_STAMP(tile_i, role, ev) do {} while (0)
#endif

struct RuParams {
  const float *bias7, *alpha2;   // conv k7 bias; Snake after it ([C] alpha | [C] 1 / (alpha + 1e-9))
  EpiConv::Params k1;            // conv k1 epilogue: bias, Snake of the consumer, residual x, raw / activated outputs
};

template <int C, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ru_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW7,
                     const __grid_constant__ CUtensorMap tmW1, const TcShape g, const RuParams rp, const int m_tiles) {
  constexpr int SW = BLOCK_K * 2, KB = C / BLOCK_K;
  constexpr int A_BYTES = kTileM * BLOCK_K * 2, B_BYTES = C * BLOCK_K * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int H_BYTES = KB * A_BYTES;
  constexpr int ACC = C <= 32 ? 32 : C <= 64 ? 64 : C <= 128 ? 128 : 256, TMEM_COLS = 2 * ACC;
  static_assert(C % BLOCK_K == 0 && C % 16 == 0 && C <= 256, "channel count");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage buffers must stay 1024B aligned for the swizzle");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* hs = smem + STAGES * STAGE_BYTES;  // h tile: KB sub-tiles [128 rows][BLOCK_K] fp16, swizzled like a TMA-written A tile
  float* par = reinterpret_cast<float*>(hs + H_BYTES);  // [3][C]: conv k7 bias | Snake alpha | 1 / (alpha + 1e-9)
  uint64_t* full = reinterpret_cast<uint64_t*>(par + 3 * C);
  uint64_t* empty = full + STAGES;
  uint64_t* acc1_full = empty + STAGES;
  uint64_t* h_ready = acc1_full + 1;   // [KB] 8 arrivals each: every epilogue warp has stored its part of K block kb of h
  uint64_t* acc2_full = h_ready + KB;
  uint64_t* acc2_free = acc2_full + 1; // 8 arrivals: acc2 drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_free + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int units7 = g.ntaps * KB;  // ring stages of the k7 mainloop per tile; KB more carry the W1 K blocks
  const int ntiles = m_tiles * g.batch;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW7) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    for (int kb = 0; kb < KB; ++kb) mbar_init(&h_ready[kb], 8);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < C; i += kGemmThreads) {
    par[i] = rp.bias7[i];
    par[C + i] = rp.alpha2[i];
    par[2 * C + i] = rp.alpha2[C + i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    int git = 0;
    int tcp = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcp) {
      const int m0 = (t % m_tiles) * kTileM, b = t / m_tiles;
      RU_STAMP(tcp, 2, 0);
      for (int it = 0; it < units7 + KB; ++it, ++git) {
        if (it == units7) RU_STAMP(tcp, 2, 1);
        const int s = git % STAGES;
        const uint32_t ph = (git / STAGES) & 1;
        uint8_t* ss = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ph ^ 1);
        if (it < units7) {
          const int tap = it / KB, kb = it % KB;
          mbar_expect_tx_elect(&full[s], STAGE_BYTES);
          tma_load_3d_elect(ss + A_BYTES, &tmW7, &full[s], kb * BLOCK_K, 0, tap);
          tma_load_3d_elect(ss, &tmA, &full[s], kb * BLOCK_K, m0 + g.tap_off[tap], b);
        } else {
          mbar_expect_tx_elect(&full[s], B_BYTES);
          tma_load_3d_elect(ss + A_BYTES, &tmW1, &full[s], (it - units7) * BLOCK_K, 0, 0);
        }
      }
      RU_STAMP(tcp, 2, 2);
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kTileM, C, 0);
    int git = 0, tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      const uint32_t tph = tc & 1;
      RU_STAMP(tc, 0, 0);
      for (int it = 0; it < units7; ++it, ++git) {
        const int s = git % STAGES;
        mbar_wait(&full[s], (git / STAGES) & 1);
        tcgen05_fence_after();
        if (it == 0) RU_STAMP(tc, 0, 1);
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc<SW>(a_addr), bdesc = make_smem_desc<SW>(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) umma_bf16_f16_elect(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(acc1_full);
      RU_STAMP(tc, 0, 2);
      mbar_wait(acc2_free, tph ^ 1);       // the previous tile's second epilogue has drained acc2 (first use passes)
      RU_STAMP(tc, 0, 3);
      for (int kb = 0; kb < KB; ++kb, ++git) {
        const int s = git % STAGES;
        mbar_wait(&h_ready[kb], tph);      // K block kb of this tile's h is in shared memory
        if (kb == 0) RU_STAMP(tc, 0, 4);
        mbar_wait(&full[s], (git / STAGES) & 1);
        tcgen05_fence_after();
        const uint64_t adesc = make_smem_desc<SW>(smem_u32(hs + kb * A_BYTES));
        const uint64_t bdesc = make_smem_desc<SW>(smem_u32(smem + s * STAGE_BYTES + A_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) umma_bf16_f16_elect(tmem_base + ACC, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(acc2_full);
      RU_STAMP(tc, 0, 5);
    }
    __syncwarp();
  } else {
    // first epilogue: the eight warps split the columns K block by K block (two warps per TMEM lane quadrant take the two
    // halves of every K block), so that the MMA warp can start the k1 GEMM on K block 0 while the others are produced
    const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
    constexpr int kChunks = C / 16, kHalf = (kChunks + 1) / 2;
    constexpr int CPB = BLOCK_K / 16, CPH = (CPB + 1) / 2;  // 16-column chunks per K block / per warp and K block
    const int c_begin = half == 0 ? 0 : kHalf * 16, c_end = half == 0 ? kHalf * 16 : C;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    int tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      const int m0 = (t % m_tiles) * kTileM, b = t / m_tiles, m = m0 + row;
      const uint32_t tph = tc & 1;
      if (warp == 2) RU_STAMP(tc, 1, 0);
      // residual rows of the second epilogue: in flight while the tile's MMAs run
      uint4 res[kHalf][2];
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) EpiConv::load_residual(rp.k1, b, 0, m, c, res[i][0], res[i][1]);
      }
      // ---- first epilogue: h = Snake(acc1 + b7) -> shared memory, K-major, 128B / 64B swizzle (what a tensor copy would write)
      mbar_wait(acc1_full, tph);
      tcgen05_fence_after();
      if (warp == 2) RU_STAMP(tc, 1, 1);
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int i = 0; i < CPH; ++i) {
          const int cc = half * CPH + i;  // chunk inside the K block
          if (cc < CPB) {
            const int c = kb * BLOCK_K + 16 * cc;
            float v[16];
            tmem_ld16(tq + c, v);
            uint4 o[2];
            __half2* h2 = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(par + c + e);
              const float4 al = *reinterpret_cast<const float4*>(par + C + c + e);
              const float4 ia = *reinterpret_cast<const float4*>(par + 2 * C + c + e);
              h2[e / 2] = __floats2half2_rn(snake_eval(v[e] + bb.x, al.x, ia.x), snake_eval(v[e + 1] + bb.y, al.y, ia.y));
              h2[e / 2 + 1] = __floats2half2_rn(snake_eval(v[e + 2] + bb.z, al.z, ia.z), snake_eval(v[e + 3] + bb.w, al.w, ia.w));
            }
            const int j = 2 * cc;  // first 16-byte chunk inside the row of this K block
            uint8_t* rowp = hs + kb * A_BYTES + row * SW;
            const int sw = SW == 128 ? (row & 7) : ((row >> 1) & 3);
            *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = o[0];
            *reinterpret_cast<uint4*>(rowp + (((j + 1) ^ sw) << 4)) = o[1];
          }
        }
        // generic-proxy stores -> visible to the tensor core's (async-proxy) operand reads, then tell the MMA warp
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_ready[kb])) : "memory");
      }
      // ---- second epilogue: acc2 + b1 + x -> raw x' and Snake(x') for the consumer
      if (warp == 2) RU_STAMP(tc, 1, 2);
      mbar_wait(acc2_full, tph);
      tcgen05_fence_after();
      if (warp == 2) RU_STAMP(tc, 1, 3);
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) {
          float v[16];
          tmem_ld16(tq + ACC + c, v);
          EpiConv::apply(rp.k1, b, 0, m, c, v, res[i][0], res[i][1]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (warp == 2) RU_STAMP(tc, 1, 4);
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc2_free)) : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// Same unit for C <= 128, skewed by one tile: with 3 x ACC <= 384 TMEM columns there is room for two acc1 buffers, and with a
// second h tile in shared memory the MMA warp never waits for the first epilogue (a third of a tile in the kernel above,
// profiles/r02_codec_fused_ru.summary.txt).  Used for the encoder's 64- and 128-channel units (launch_ru_fused):
//   MMA warp        k7(0) | k7(1) k1(0) | k7(2) k1(1) | ...      k7(t) -> acc1[t & 1], k1(t): hs[t & 1] x W1 -> acc2
//   epilogue warps  epi1(0) | epi1(1) epi2(0) | epi1(2) epi2(1) | ...   epi1(t): acc1[t & 1] -> hs[t & 1], epi2(t): acc2 -> outputs
//   producer        the ring stages in the MMA warp's order
// acc1[t & 1] is free for k7(t + 2) because k1(t), issued before it, waited for every K block of h(t); hs[t & 1] is free for
// epi1(t + 2) because epi2(t), which precedes it, waited for k1(t)'s commit; acc2 is single-buffered (acc2_free as above).
// ------------------------------------------------------------------------------------------------
template <int C, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ru_fused_skew_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW7,
                          const __grid_constant__ CUtensorMap tmW1, const TcShape g, const RuParams rp, const int m_tiles) {
  constexpr int SW = BLOCK_K * 2, KB = C / BLOCK_K;
  constexpr int A_BYTES = kTileM * BLOCK_K * 2, B_BYTES = C * BLOCK_K * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int H_BYTES = KB * A_BYTES;
  constexpr int ACC = C <= 32 ? 32 : C <= 64 ? 64 : 128, TMEM_COLS = ACC == 128 ? 512 : 4 * ACC;  // 3 x ACC, rounded up to 2^n
  static_assert(C % BLOCK_K == 0 && C % 16 == 0 && C <= 128, "channel count");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage buffers must stay 1024B aligned for the swizzle");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* hs = smem + STAGES * STAGE_BYTES;  // two h tiles: KB sub-tiles [128 rows][BLOCK_K] fp16 each, swizzled like a TMA-written A tile
  float* par = reinterpret_cast<float*>(hs + 2 * H_BYTES);  // [3][C]: conv k7 bias | Snake alpha | 1 / (alpha + 1e-9)
  uint64_t* full = reinterpret_cast<uint64_t*>(par + 3 * C);
  uint64_t* empty = full + STAGES;
  uint64_t* acc1_full = empty + STAGES;  // [2]
  uint64_t* h_ready = acc1_full + 2;     // [2][KB] 8 arrivals each
  uint64_t* acc2_full = h_ready + 2 * KB;
  uint64_t* acc2_free = acc2_full + 1;   // 8 arrivals: acc2 drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_free + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int units7 = g.ntaps * KB;
  const int ntiles = m_tiles * g.batch;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW7) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&acc1_full[i], 1);
    for (int i = 0; i < 2 * KB; ++i) mbar_init(&h_ready[i], 8);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < C; i += kGemmThreads) {
    par[i] = rp.bias7[i];
    par[C + i] = rp.alpha2[i];
    par[2 * C + i] = rp.alpha2[C + i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    int git = 0;
    auto load_k7 = [&](int t) {
      const int m0 = (t % m_tiles) * kTileM, b = t / m_tiles;
      for (int it = 0; it < units7; ++it, ++git) {
        const int s = git % STAGES;
        uint8_t* ss = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ((git / STAGES) & 1) ^ 1);
        const int tap = it / KB, kb = it % KB;
        mbar_expect_tx_elect(&full[s], STAGE_BYTES);
        tma_load_3d_elect(ss + A_BYTES, &tmW7, &full[s], kb * BLOCK_K, 0, tap);
        tma_load_3d_elect(ss, &tmA, &full[s], kb * BLOCK_K, m0 + g.tap_off[tap], b);
      }
    };
    auto load_w1 = [&]() {
      for (int kb = 0; kb < KB; ++kb, ++git) {
        const int s = git % STAGES;
        uint8_t* ss = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ((git / STAGES) & 1) ^ 1);
        mbar_expect_tx_elect(&full[s], B_BYTES);
        tma_load_3d_elect(ss + A_BYTES, &tmW1, &full[s], kb * BLOCK_K, 0, 0);
      }
    };
    if (my_tiles > 0) load_k7(blockIdx.x);
    for (int tc = 0; tc < my_tiles; ++tc) {
      if (tc + 1 < my_tiles) load_k7(blockIdx.x + (tc + 1) * gridDim.x);
      load_w1();
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kTileM, C, 0);
    int git = 0;
    auto mma_k7 = [&](int tc) {  // tile tc of this CTA -> acc1[tc & 1]
      const uint32_t acc = tmem_base + (uint32_t)((tc & 1) * ACC);
      for (int it = 0; it < units7; ++it, ++git) {
        const int s = git % STAGES;
        mbar_wait(&full[s], (git / STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc<SW>(a_addr), bdesc = make_smem_desc<SW>(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) umma_bf16_f16_elect(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(&acc1_full[tc & 1]);
    };
    if (my_tiles > 0) mma_k7(0);
    for (int tc = 0; tc < my_tiles; ++tc) {
      if (tc + 1 < my_tiles) mma_k7(tc + 1);
      mbar_wait(acc2_free, (tc & 1) ^ 1);  // the previous tile's second epilogue has drained acc2 (first use passes)
      const uint32_t hph = (tc >> 1) & 1;
      for (int kb = 0; kb < KB; ++kb, ++git) {
        const int s = git % STAGES;
        mbar_wait(&h_ready[(tc & 1) * KB + kb], hph);  // K block kb of this tile's h is in shared memory
        mbar_wait(&full[s], (git / STAGES) & 1);
        tcgen05_fence_after();
        const uint64_t adesc = make_smem_desc<SW>(smem_u32(hs + (tc & 1) * H_BYTES + kb * A_BYTES));
        const uint64_t bdesc = make_smem_desc<SW>(smem_u32(smem + s * STAGE_BYTES + A_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
          umma_bf16_f16_elect(tmem_base + 2 * ACC, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(acc2_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
    constexpr int kChunks = C / 16, kHalf = (kChunks + 1) / 2;
    constexpr int CPB = BLOCK_K / 16, CPH = (CPB + 1) / 2;  // 16-column chunks per K block / per warp and K block
    const int c_begin = half == 0 ? 0 : kHalf * 16, c_end = half == 0 ? kHalf * 16 : C;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    // second epilogue of tile tc: acc2 + b1 + x -> raw x' and Snake(x') for the consumer
    auto epi2 = [&](int tc) {
      const int t = blockIdx.x + tc * gridDim.x;
      const int m = (t % m_tiles) * kTileM + row, b = t / m_tiles;
      uint4 res[kHalf][2];
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) EpiConv::load_residual(rp.k1, b, 0, m, c, res[i][0], res[i][1]);
      }
      mbar_wait(acc2_full, tc & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int c = c_begin + 16 * i;
        if (c < c_end) {
          float v[16];
          tmem_ld16(tq + 2 * ACC + c, v);
          EpiConv::apply(rp.k1, b, 0, m, c, v, res[i][0], res[i][1]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc2_free)) : "memory");
    };
    for (int tc = 0; tc < my_tiles; ++tc) {
      // ---- first epilogue of tile tc: h = Snake(acc1 + b7) -> hs[tc & 1], K block by K block
      uint8_t* hb = hs + (tc & 1) * H_BYTES;
      const uint32_t ta = tq + (uint32_t)((tc & 1) * ACC);
      mbar_wait(&acc1_full[tc & 1], (tc >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int i = 0; i < CPH; ++i) {
          const int cc = half * CPH + i;  // chunk inside the K block
          if (cc < CPB) {
            const int c = kb * BLOCK_K + 16 * cc;
            float v[16];
            tmem_ld16(ta + c, v);
            uint4 o[2];
            __half2* h2 = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(par + c + e);
              const float4 al = *reinterpret_cast<const float4*>(par + C + c + e);
              const float4 ia = *reinterpret_cast<const float4*>(par + 2 * C + c + e);
              h2[e / 2] = __floats2half2_rn(snake_eval(v[e] + bb.x, al.x, ia.x), snake_eval(v[e + 1] + bb.y, al.y, ia.y));
              h2[e / 2 + 1] = __floats2half2_rn(snake_eval(v[e + 2] + bb.z, al.z, ia.z), snake_eval(v[e + 3] + bb.w, al.w, ia.w));
            }
            const int j = 2 * cc;  // first 16-byte chunk inside the row of this K block
            uint8_t* rowp = hb + kb * A_BYTES + row * SW;
            const int sw = SW == 128 ? (row & 7) : ((row >> 1) & 3);
            *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = o[0];
            *reinterpret_cast<uint4*>(rowp + (((j + 1) ^ sw) << 4)) = o[1];
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_ready[(tc & 1) * KB + kb])) : "memory");
      }
      if (tc > 0) epi2(tc - 1);
    }
    if (my_tiles > 0) epi2(my_tiles - 1);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// Fused decode step of the bf16 path for up to 64 sequence rows (one new position per row): ONE cooperative kernel runs
// the 24 layers + final norm + heads that transformer_pass_bf16 issues as 170 launches.  Each phase is the same tile
// arithmetic as the stand-alone kernels (rmsnorm_bf16_kernel, gemm_tc_kernel with UMMA M = 64 and 4 K blocks per stage,
// attn_bf16_kernel restated per warp); phases are separated by a device-wide barrier instead of a kernel boundary, so the
// ~4 us of launch + prologue (barrier init, TMEM allocation, tensor-map fetch, pipeline fill from a cold start) per kernel
// is paid once per step.  Every GEMM phase has at most one tile per CTA (144 / 128 tiles on 148 SMs).
// Cross-CTA data (h, xn, q, K/V, attention output, SwiGLU output) is read with L2 loads or TMA after the barrier's
// acquire; a proxy fence on the reader's side (after the acquire, before any TMA is issued) orders the generic-proxy stores
// of the previous phase before the async-proxy reads (a second fence on the writer's side measured 1.6 % slower).
// ------------------------------------------------------------------------------------------------
namespace fused {
#ifndef VAURA_FUSED_KSUB
#define VAURA_FUSED_KSUB 4  // K blocks per ring stage (A/B builds: -DVAURA_FUSED_KSUB=2 -> twice as many half-size stages)
#endif
constexpr int kKsub = VAURA_FUSED_KSUB, kBlockK = 64;
constexpr int kAccCols = 64;  // TMEM columns of the accumulator (BN <= 64)
constexpr int kMaxParts = kFusedKsplit;  // K slices of a residual GEMM (wo, w2) = partial-sum slices a row CTA adds up
constexpr int kBMax = 64 * kBlockK * 2;                   // 8 KB weight sub-tile (BN = 64; BN = 32 uses half of it)
// TM = UMMA M = rows of the activation tile: 64 (up to 64 sequence rows) or 128 (up to 128, e.g. 64 clips with CFG)
template <int TM>
struct Geo {
  static constexpr int kABytes = TM * kBlockK * 2;                 // 8 / 16 KB activation sub-tile
  static constexpr int kSubBytes = kABytes + kBMax;
  static constexpr int kStageBytes = kKsub * kSubBytes;            // 64 / 96 KB per stage
  static constexpr int kStages = (TM == 64 ? 12 : 8) / kKsub;      // 192 KB of ring either way (3 / 2 stages of 4 K blocks)
  static constexpr int kRing = kStages * kStageBytes;
};
constexpr int kAttnScratch = 10 * (kHeadDim + kMaxCtx) * 4;                  // per-warp q + scores
// attention staging: every warp owns kAttnSlots slots of one 16-position run of K or V rows (16 x 192 B) inside the idle
// GEMM ring, each with its own mbarrier
constexpr int kAttnSlots = 6, kRunBytes = 16 * kHeadDim * 2;
constexpr int kAttnBars = 10 * kAttnSlots * 8;
constexpr int kSmem = 3 * 65536 + 256 + kAttnScratch + 512 + 1024;
static_assert(10 * kAttnSlots * kRunBytes <= 3 * 65536 && kAttnBars <= 512, "attention staging fits the GEMM ring");
static_assert(Geo<64>::kRing == 3 * 65536 && Geo<128>::kRing == 3 * 65536, "ring size");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// device-wide barrier: no bulk copy is in flight when a CTA arrives (its mainloop has drained), so the release /
// acquire pair costs its idle latency (~0.75 us, profiles/r01_probe_sync_latency_under_tma.txt)
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned target, unsigned long long* arrival = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (arrival) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(*arrival));  // every warp of the CTA is done with the phase
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    if ((int)(ld_acquire_u32(counter) - target) < 0) {
      const long long t0 = clock64();
      while ((int)(ld_acquire_u32(counter) - target) < 0)
        if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
}
}  // namespace fused

// per-CTA pipeline state that survives across the GEMM tiles of a step
struct FusedPipe {
  uint8_t* smem;
  uint64_t *full, *empty, *tmem_full;
  uint32_t tmem_base;
  int git;     // ring iterations so far (identical in the producer and the MMA thread)
  int tiles;   // accumulator uses so far (epilogue warps)
};

// one 64 x BN output tile: A = activations [64 x K] (map tmA, rows 0..63), B = weight rows [n0, n0 + BN) of layer `layer`
// (map tmB), K blocks [kb0, kb1); warp 0 = TMA, warp 1 = MMA issue, warps 2-9 = epilogue
template <int BN, int TM>
__device__ __forceinline__ void fused_gemm_tile(FusedPipe& pp, const CUtensorMap* tmA, const CUtensorMap* tmB, int layer, int n0,
                                                int kb0, int kb1, const EpiLinear::Params& ep,
                                                unsigned long long* dbg = nullptr) {
  // dbg (profiles/fused_timing.py): [0..7] stage loads issued, [8..15] stage landed, [16..23] stage MMAs issued,
  // [24] accumulator complete, [25] epilogue done, [26] entry
  auto dstamp = [&](int i) {
    if (dbg) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[i] = t;
    }
  };
  using namespace fused;
  using GE = Geo<TM>;
  constexpr int kStages = GE::kStages, kStageBytes = GE::kStageBytes, kABytes = GE::kABytes;
  constexpr int B_BYTES = BN * kBlockK * 2;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int iters = (kb1 - kb0 + kKsub - 1) / kKsub;
  const int rot = blockIdx.x % iters;
  if (warp == 0) {
    // the whole warp walks the stages (uniform control flow); one elected lane issues the copies
    if (dbg && lane == 0) dstamp(26);
    for (int it = 0; it < iters; ++it) {
      const int gi = pp.git + it, s = gi % kStages;
      const uint32_t ph = (gi / kStages) & 1;
      // K groups are visited in a CTA-dependent rotation (neutral in measurements, kept: the CTAs of a phase read the
      // same activation matrix, this keeps them from asking the same L2 lines at the same time)
      const int grp = (it + rot) % iters;
      const int u0 = kb0 + grp * kKsub;
      uint8_t* ss = pp.smem + s * kStageBytes;
      mbar_wait(&pp.empty[s], ph ^ 1);
      // one box per operand covers the stage's kKsub K blocks (a copy instruction costs the issuing thread ~0.13 us
      // whatever its size, profiles/probes/tma_rate.cu); a box that runs past kb1 or past K still delivers its full
      // byte count
      mbar_expect_tx_elect(&pp.full[s], kKsub * (kABytes + B_BYTES));
      tma_load_4d_elect(ss + kKsub * kABytes, tmB, &pp.full[s], 0, n0, u0, layer);
      tma_load_4d_elect(ss, tmA, &pp.full[s], 0, 0, u0, 0);
      if (dbg && lane == 0 && it < 8) dstamp(it);
    }
    __syncwarp();
  } else if (warp == 1) {
    // the whole warp walks the stages (uniform control flow); one elected lane issues the MMAs and the commits
    constexpr uint32_t idesc = make_idesc(TM, BN, 1);
    for (int it = 0; it < iters; ++it) {
      const int gi = pp.git + it, s = gi % kStages;
      const uint32_t ph = (gi / kStages) & 1;
      mbar_wait(&pp.full[s], ph);
      tcgen05_fence_after();
      if (dbg && lane == 0 && it < 8) dstamp(8 + it);
      const int nsub = min(kKsub, kb1 - (kb0 + (it + rot) % iters * kKsub));
      // stage layout: kKsub activation sub-tiles, then kKsub weight sub-tiles (each [rows x 128 B], 128B-swizzled);
      // descriptors differ only in their 14-bit address field: +2 per 32-byte K step inside the swizzle atom
      const uint32_t st_addr = smem_u32(pp.smem + s * kStageBytes);
      static_assert((make_smem_desc<128>(0) >> 32) == 0x40004040ull, "descriptor high word");
      // broadcast from lane 0: the compiler then keeps the descriptor words in uniform registers (UIADD3 per MMA instead of four
      // R2UR.BROADCAST from vector registers)
      const uint32_t a_lo = __shfl_sync(0xffffffffu, (st_addr & 0x3FFFF) >> 4, 0);
      const uint32_t b_lo = __shfl_sync(0xffffffffu, ((st_addr + kKsub * kABytes) & 0x3FFFF) >> 4, 0);
      // one rolled loop over the stage's K blocks (four MMAs each): the tile is paced by the SM's L2 port, not by MMA issue, and
      // the kernel is several times the L1.5 instruction cache - smaller code, not faster issue, is what helps here
#pragma unroll 1
      for (int sub = 0; sub < nsub; ++sub)
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16_f16_elect_lo(pp.tmem_base, a_lo + sub * (kABytes >> 4) + 2 * k, b_lo + sub * (B_BYTES >> 4) + 2 * k, idesc,
                                 (it | sub | k) != 0);
      umma_commit_elect(&pp.empty[s]);
      if (dbg && lane == 0 && it < 8) dstamp(16 + it);
    }
    umma_commit_elect(pp.tmem_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    // UMMA M = 128: accumulator row i sits in lane i; M = 64: rows 16q..16q+15 sit in lanes 32q..32q+15
    const int m = TM == 128 ? q * 32 + lane : q * 16 + lane;
    const bool row_ok = TM == 128 || lane < 16;
    constexpr int kChunks = BN / 16, kHalf = (kChunks + 1) / 2;
    const int c_begin = (warp - 2) < 4 ? 0 : kHalf * 16, c_end = (warp - 2) < 4 ? kHalf * 16 : BN;
    // wqkv tiles (BN = 32: one chunk per warp): RoPE pairs and the cache row are requested before the accumulator wait
    const bool qkv_pre = BN == 32 && ep.mode == EPI_QKV;
    EpiLinear::QkvPre pre;
    if (qkv_pre && row_ok) EpiLinear::prefetch_qkv(ep, m, n0 + c_begin, pre);
    mbar_wait(pp.tmem_full, pp.tiles & 1);
    tcgen05_fence_after();
    if (warp == 2 && lane == 0) dstamp(24);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 16) {
      float v[16];
      tmem_ld16(pp.tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      if (row_ok) {
        if (qkv_pre) EpiLinear::apply_qkv(ep, m, n0 + c, v, pre);
        else EpiLinear::apply(ep, 0, 0, m, n0 + c, v);
      }
    }
    tcgen05_fence_before();
    if (warp == 2 && lane == 0) dstamp(25);
  }
  pp.git += iters;
  pp.tiles += 1;
}

// DET: reproducible mode (FusedStepArgs::part), a separate instantiation so that the default kernel's code is untouched
template <int TM, bool DET>
__global__ void __launch_bounds__(kGemmThreads, 1)
decode_step_fused_bf16(const __grid_constant__ CUtensorMap tm_xn, const __grid_constant__ CUtensorMap tm_attn,
                       const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wqkv,
                       const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_w13,
                       const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_heads,
                       const __grid_constant__ FusedStepArgs a) {
  using namespace fused;
  using GE = Geo<TM>;
  constexpr int kStages = GE::kStages, kStageBytes = GE::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  FusedPipe pp;
  pp.smem = smem;
  pp.full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  pp.empty = pp.full + kStages;
  pp.tmem_full = pp.empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pp.tmem_full + 1);
  float* scratch = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);
  uint64_t* attn_bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + 256 + kAttnScratch);
  pp.git = 0;
  pp.tiles = 0;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // provably warp-uniform
  const int cta = blockIdx.x, G = gridDim.x;
  const int R = a.R, D = a.D, F = a.F;
  const unsigned epoch = a.state->epoch;
  const int offset = a.state->offset;
  const int p = offset - 1;  // position fed by this step
  if (cta == 0 && tid == 0 && a.step_times) {  // step-to-step latency (bench.py: p50)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.step_times[offset] = t;
  }
  const unsigned nbar = (unsigned)(5 * a.L + 1 + (a.fuse_io ? 1 : 0));
  unsigned bi = 0;
  int stamp_i = 0;
  auto stamp = [&]() {  // optional phase timestamps of CTA 0 (profiles/fused_timing.py)
    if (a.timing && cta == a.timing_cta && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.timing[stamp_i] = t;
    }
    ++stamp_i;
  };
  auto sync_all = [&]() {
    stamp();
    ++bi;
    // optional: when every CTA arrives at the five barriers of the last layer (timing[1300 + phase * G + cta];
    // profiles/fused_arrivals.py): which CTAs a phase waits for
    // (compiled in with -DVAURA_FUSED_ARRIVALS only: the stamps cost 0.3 % of a step even when they are off)
    unsigned long long* arrival = nullptr;
#ifdef VAURA_FUSED_ARRIVALS
    if (a.timing) {
      const int pi = (int)bi - (2 + 5 * (a.L - 1));
      if (pi >= 0 && pi < 5 && 1300 + pi * G + cta < 2048) arrival = a.timing + 1300 + pi * G + cta;
    }
#endif
    grid_sync(&a.state->barrier, (epoch * nbar + bi) * (unsigned)G, arrival);
    stamp();
  };

  if (warp == 0 && lane == 0) {
    const CUtensorMap* maps[8] = {&tm_xn, &tm_attn, &tm_act, &tm_wqkv, &tm_wo, &tm_w13, &tm_w2, &tm_heads};
    for (int i = 0; i < 8; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(maps[i]) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&pp.full[s], 1);
      mbar_init(&pp.empty[s], 1);
    }
    mbar_init(pp.tmem_full, 1);
    for (int i = 0; i < 10 * kAttnSlots; ++i) mbar_init(&attn_bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kAccCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pp.tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // ---- RMSNorm of the residual row `cta` (llama.py:147-158): fp32 math, bf16 output = the next GEMM's A operand.
  //      embed = true (first phase of a step with fuse_io): the row is built here from the conditioning row and the 9
  //      folded token tables (llama.py:455-472, what embed_kernel does) and stored to h on the way ----
  //      nparts > 0 (after a split-K residual GEMM with a.part): the row is first completed, h[row] += part[0][row] + ... +
  //      part[nparts-1][row] in that fixed order (every output element has one owner: no float atomics, results are
  //      reproducible from run to run), and stored back ----
  auto rmsnorm_phase = [&](const float* w, bool embed, int nparts = 0) {
    if (cta < R) {
      float* red = scratch;
      float* hrow = a.h + (size_t)cta * D;
      const int n4 = D >> 2;
      float4 v[2], gw[2];
      float ss = 0.f;
      // the norm weights do not depend on the row: requested first, in flight together with the row
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kGemmThreads;
        gw[i] = c < n4 ? __ldg(reinterpret_cast<const float4*>(w) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kGemmThreads;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < n4) {
          if (!embed) {
            v[i] = __ldcg(reinterpret_cast<const float4*>(hrow) + c);
            if (DET && nparts > 0) {
              float4 pv[fused::kMaxParts];
#pragma unroll
              for (int sp = 0; sp < fused::kMaxParts; ++sp)
                if (sp < nparts) pv[sp] = __ldcg(reinterpret_cast<const float4*>(a.part + ((size_t)sp * R + cta) * D) + c);
#pragma unroll
              for (int sp = 0; sp < fused::kMaxParts; ++sp)
                if (sp < nparts) { v[i].x += pv[sp].x; v[i].y += pv[sp].y; v[i].z += pv[sp].z; v[i].w += pv[sp].w; }
              reinterpret_cast<float4*>(hrow)[c] = v[i];
            }
          } else {
            const int C = a.cond_dim, TD = D - C, f = 4 * c;
            if (f < C) {
              int vrow = p / a.atpvf;
              if (vrow > a.cond_tokens) vrow = a.cond_tokens;  // >= Tv -> empty_video_emb row (llama.py:569-572)
              v[i] = __ldg(reinterpret_cast<const float4*>(a.cond_rows + ((size_t)cta * (a.cond_tokens + 1) + vrow) * C + f));
            } else {
              const int bt = cta % a.batch;  // CFG halves share the token sequence (vaura_model.py:795)
              for (int k = 0; k < a.Kc; ++k) {  // python sum() adds the codebooks in order (llama.py:455-460)
                const int tok = a.seq[((size_t)bt * a.Kc + k) * a.S + p];
                const float4 tv = __ldg(reinterpret_cast<const float4*>(a.tables + ((size_t)k * (a.vocab + 1) + tok) * TD + (f - C)));
                v[i].x += tv.x; v[i].y += tv.y; v[i].z += tv.z; v[i].w += tv.w;
              }
            }
            reinterpret_cast<float4*>(hrow)[c] = v[i];
          }
        }
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
      ss = warp_sum(ss);
      if (lane == 0) red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
      for (int i = 0; i < kGemmThreads / 32; ++i) tot += red[i];
      const float rs = rsqrtf(tot / (float)D + a.eps);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kGemmThreads;
        if (c < n4) {
          const float4 g = gw[i];
          uint2 o;
          *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[i].x * rs * g.x, v[i].y * rs * g.y);
          *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[i].z * rs * g.z, v[i].w * rs * g.w);
          *reinterpret_cast<uint2*>(a.xn + (size_t)cta * D + 4 * c) = o;
        }
      }
    }
  };

  // ---- attention (llama.py:246-255): one warp per (row, head) over the paged bf16 cache, fp32 softmax.  A 16-position
  //      run of K or V rows of one head is contiguous inside its page (3 KB): the warp stages the runs of its items -
  //      all K runs, then all V runs, then the next item - with one bulk copy each into a private ring of kAttnSlots
  //      slots in the idle GEMM ring, so up to 18 KB per warp are in flight independent of registers and the V rows
  //      (and the next item's K rows) stream in while the scores are computed.  Positions are consumed 32 at a time.
  // the warp's items and their page tables do not change during the launch: lane i holds page i (<= 16 pages for 256
  // positions of 16)
  const int att_item_stride = G * (kGemmThreads / 32), att_item0 = warp * G + cta;  // CTA-fastest: every SM gets ~7 items
  const int att_nitems = att_item0 < R * a.H ? (R * a.H - att_item0 + att_item_stride - 1) / att_item_stride : 0;  // <= 2
  int att_pg[2] = {0, 0}, att_row[2] = {0, 0}, att_hd[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (i < att_nitems) {
      const int item = att_item0 + i * att_item_stride;
      att_row[i] = item / a.H;
      att_hd[i] = item % a.H;
      if (lane < a.kv.max_pages_per_seq) att_pg[i] = a.kv.page_table[att_row[i] * a.kv.max_pages_per_seq + lane];
    }
  int att_slot = 0;         // ring position of the next copy to issue (kernel lifetime)
  int att_rslot = 0;        // ring position of the next copy to consume
  unsigned att_rphase = 0;  // bit s = parity of the phase the consumer waits for on slot s
  auto attention_phase = [&](int layer) {
    float* qs = scratch + warp * (kHeadDim + kMaxCtx);
    float* sc = qs + kHeadDim;
    const __nv_bfloat16* kvp = reinterpret_cast<const __nv_bfloat16*>(a.kv.pages);
    const int nctx = p + 1, psz = a.kv.page_size, psh = 31 - __clz(psz);
    const size_t page_stride = (size_t)a.kv.nhead * psz * kHeadDim;
    const int nitems = att_nitems;
    const int nr = (nctx + 15) >> 4, total = nitems * 2 * nr;
    uint8_t* stage = smem + warp * (kAttnSlots * kRunBytes);
    uint64_t* abar = attn_bars + warp * kAttnSlots;
    // q of the first item: requested before anything else so that its latency overlaps the first copies
    unsigned short qraw[3] = {0, 0, 0};
    if (nitems)
#pragma unroll
      for (int i = 0; i < 3; ++i)
        qraw[i] = __ldcg(reinterpret_cast<const unsigned short*>(a.q) + (size_t)att_row[0] * D + att_hd[0] * kHeadDim + lane + 32 * i);
    int issued = 0, consumed = 0;
    int astamp_i = 0;
    auto astamp = [&]() {  // sub-phase timestamps of (timing CTA, thread 0) in the last layer: timing[900 ...]
      if (a.timing && cta == a.timing_cta && tid == 0 && layer == a.L - 1) {
        unsigned long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        a.timing[900 + astamp_i] = tt;
      }
      ++astamp_i;
    };
    astamp();
    // issue cursor: item, K (0) / V (1), run; copies go out in consumption order (all K runs, all V runs, next item)
    int c_it = 0, c_kv = 0, c_run = 0;
    auto issue_one = [&]() {
      const int j = c_run * 16;
      const int page = __shfl_sync(0xffffffffu, c_it ? att_pg[1] : att_pg[0], j >> psh);
      const int hd = c_it ? att_hd[1] : att_hd[0];
      const __nv_bfloat16* src = kvp + ((size_t)(layer * 2 + c_kv) * a.kv.num_pages + page) * page_stride +
                                 ((size_t)hd * psz + (j & (psz - 1))) * kHeadDim;
      mbar_expect_tx_elect(&abar[att_slot], kRunBytes);
      bulk_load_1d_elect(stage + att_slot * kRunBytes, src, kRunBytes, &abar[att_slot]);
      att_slot = att_slot + 1 == kAttnSlots ? 0 : att_slot + 1;
      ++issued;
      if (++c_run == nr) { c_run = 0; if (++c_kv == 2) { c_kv = 0; ++c_it; } }
    };
    auto refill = [&]() {
      __syncwarp();  // every lane is done reading the slots counted in `consumed`
      while (issued < total && issued - consumed < kAttnSlots) issue_one();
    };
    // wait for the next copy in consumption order and return its slot
    auto staged = [&]() -> const uint8_t* {
      const int sl = att_rslot;
      mbar_wait(&abar[sl], (att_rphase >> sl) & 1u);
      att_rphase ^= 1u << sl;
      att_rslot = sl + 1 == kAttnSlots ? 0 : sl + 1;
      return stage + sl * kRunBytes;
    };
    refill();
    astamp();
    for (int it = 0; it < nitems; ++it) {
      const int row = it ? att_row[1] : att_row[0], hd = it ? att_hd[1] : att_hd[0];
      __syncwarp();
      if (it)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          qraw[i] = __ldcg(reinterpret_cast<const unsigned short*>(a.q) + (size_t)row * D + hd * kHeadDim + lane + 32 * i);
#pragma unroll
      for (int i = 0; i < 3; ++i) qs[lane + 32 * i] = __bfloat162float(__ushort_as_bfloat16(qraw[i]));
      __syncwarp();
      // scores: 4 lanes per key position; lane t takes the 16-byte chunks t, t+4, t+8 of the 192-byte row (conflict-free:
      // a quarter warp reads two rows x four chunks); 4 x 8 positions per iteration
      const int g = lane >> 2, t = lane & 3;
      float q24[24];
#pragma unroll
      for (int i = 0; i < 24; ++i) q24[i] = qs[(4 * (i >> 3) + t) * 8 + (i & 7)];
      float mx = -INFINITY;
      astamp();
      for (int j0 = 0; j0 < nctx; j0 += 32) {
        const bool two = j0 + 16 < nctx;
        const uint8_t* runA = staged();
        if (j0 == 0) astamp();
        const uint8_t* runB = two ? staged() : runA;
        // two rows per half (rows g and 8 + g of run A, then of run B): six independent 8-term chains in flight
        float sd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (h2 == 1 && !two) break;
          const uint8_t* kr = (h2 ? runB : runA) + (g * 12 + t) * 16;
          uint4 kk[2][3];
#pragma unroll
          for (int u2 = 0; u2 < 2; ++u2)
#pragma unroll
            for (int c = 0; c < 3; ++c) kk[u2][c] = lds_u4(kr + u2 * (8 * 12 * 16) + 64 * c);  // rows past nctx: read, not used
          float ps[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const uint32_t w = e == 0 ? kk[u2][c].x : e == 1 ? kk[u2][c].y : e == 2 ? kk[u2][c].z : kk[u2][c].w;
                ps[u2][c] = fmaf(q24[c * 8 + 2 * e], bf16_lo(w), ps[u2][c]);
                ps[u2][c] = fmaf(q24[c * 8 + 2 * e + 1], bf16_hi(w), ps[u2][c]);
              }
          sd[2 * h2] = (ps[0][0] + ps[0][1]) + ps[0][2];
          sd[2 * h2 + 1] = (ps[1][0] + ps[1][1]) + ps[1][2];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], 2);
#pragma unroll
        for (int u = 0; u < 4; ++u) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], 1);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = j0 + 8 * u + g;
          if (jj < nctx) {
            const float sv = sd[u] * a.scale;
            if (t == 0) sc[jj] = sv;
            mx = fmaxf(mx, sv);
          }
        }
        consumed += two ? 2 : 1;
        refill();
      }
      astamp();
      mx = warp_max(mx);
      __syncwarp();
      float sum = 0.f;
      for (int jj = lane; jj < ((nctx + 31) & ~31); jj += 32) {  // positions past the context get probability 0
        const float e = jj < nctx ? expf(sc[jj] - mx) : 0.f;
        __syncwarp();
        // within a block of 32 positions: even positions first, then odd ones (what a P.V lane reads is contiguous)
        sc[(jj & ~31) + (lane & 1) * 16 + (lane >> 1)] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      __syncwarp();
      // P.V: lanes 0-11 take the even positions, lanes 12-23 the odd ones, 8 dims (one 16-byte chunk) each; eight
      // independent accumulators, eight value rows in flight
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      const int half = lane >= 12 ? 1 : 0, dl = lane < 24 ? lane - 12 * half : 0;
      astamp();
      for (int j0 = 0; j0 < nctx; j0 += 32) {
        const bool two = j0 + 16 < nctx;
        const uint8_t* runA = staged() + (half * 12 + dl) * 16;
        if (j0 == 0) astamp();
        const uint8_t* runB = two ? staged() + (half * 12 + dl) * 16 : runA;
        float pj[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 pv = lds_u4(sc + j0 + half * 16 + 4 * i);
          pj[4 * i] = __uint_as_float(pv.x); pj[4 * i + 1] = __uint_as_float(pv.y);
          pj[4 * i + 2] = __uint_as_float(pv.z); pj[4 * i + 3] = __uint_as_float(pv.w);
        }
#pragma unroll
        for (int b8 = 0; b8 < 2; ++b8) {
          if (b8 == 1 && !two) break;
          const uint8_t* rb = b8 ? runB : runA;
          uint4 vv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)  // row 2 i + half of the run; rows past the context may hold anything (NaN bit patterns)
            vv[i] = j0 + 16 * b8 + 2 * i + half < nctx ? lds_u4(rb + i * 24 * 16) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float pw = pj[8 * b8 + i];
            o[0] = fmaf(pw, bf16_lo(vv[i].x), o[0]); o[1] = fmaf(pw, bf16_hi(vv[i].x), o[1]);
            o[2] = fmaf(pw, bf16_lo(vv[i].y), o[2]); o[3] = fmaf(pw, bf16_hi(vv[i].y), o[3]);
            o[4] = fmaf(pw, bf16_lo(vv[i].z), o[4]); o[5] = fmaf(pw, bf16_hi(vv[i].z), o[5]);
            o[6] = fmaf(pw, bf16_lo(vv[i].w), o[6]); o[7] = fmaf(pw, bf16_hi(vv[i].w), o[7]);
          }
        }
        consumed += two ? 2 : 1;
        refill();
      }
      astamp();
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += __shfl_down_sync(0xffffffffu, o[i], 12);
      if (lane < 12) {
        const float inv = 1.f / sum;
        uint4 ov;
        *reinterpret_cast<__nv_bfloat162*>(&ov.x) = __floats2bfloat162_rn(o[0] * inv, o[1] * inv);
        *reinterpret_cast<__nv_bfloat162*>(&ov.y) = __floats2bfloat162_rn(o[2] * inv, o[3] * inv);
        *reinterpret_cast<__nv_bfloat162*>(&ov.z) = __floats2bfloat162_rn(o[4] * inv, o[5] * inv);
        *reinterpret_cast<__nv_bfloat162*>(&ov.w) = __floats2bfloat162_rn(o[6] * inv, o[7] * inv);
        *reinterpret_cast<uint4*>(a.attn + (size_t)row * D + hd * kHeadDim + 8 * lane) = ov;
      }
    }
  };

  EpiLinear::Params ep{};
  // the position is known here: no dependent load of state->offset in front of the RoPE / KV-append epilogue
  ep.R = R; ep.rope = a.rope; ep.kv = a.kv; ep.state = nullptr; ep.pos0 = p; ep.npos = 1; ep.d_model = D;
  ep.perm_S = 0; ep.perm_V = 0;

  // The RMSNorm after a split-K GEMM (wo, w2) needs every tile's reductions, the GEMM after it needs every row's norm:
  // instead of two device-wide barriers the tile CTAs count themselves on state->tiles_done (release) and only the R
  // CTAs that normalise a row wait for the count (acquire); the one barrier that follows covers both.
  const unsigned tiles_wo = (unsigned)(D / 64 * a.wo_ksplit), tiles_w2 = (unsigned)(D / 64 * a.w2_ksplit);
  const unsigned tiles_base = epoch * (unsigned)a.L * (tiles_wo + tiles_w2);
  unsigned tiles_seen = 0;
  auto tiles_arrive = [&]() {
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&a.state->tiles_done) : "memory");
  };
  auto tiles_wait = [&]() {
    if (tid == 0) {
      const unsigned target = tiles_base + tiles_seen;
      const long long t0 = clock64();
      while ((int)(ld_acquire_u32(&a.state->tiles_done) - target) < 0)
        if (clock64() - t0 > 4000000000ll) __trap();
    }
    __syncthreads();
  };

  // ---- L2 prefetch of the weight tile this CTA multiplies in the NEXT GEMM phase.  A GEMM phase is short (one tile per
  //      CTA) and its ring holds about half of the tile's operands, so weight bytes that come from HBM cost two memory
  //      round trips per phase; requested one phase ahead they are L2 hits when the tile runs, and HBM works in the
  //      background of a phase instead of on its critical path.  Issued by warp 2 (an epilogue warp, idle until the
  //      accumulator is complete): one request per weight row segment. ----
  auto prefetch_rows = [&](const __nv_bfloat16* w, size_t row0, int nrows, size_t ld, size_t col0, int ncols) {
    if (warp != 2 || !a.l2_prefetch) return;
    for (int r = lane; r < nrows; r += 32)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w + (row0 + r) * ld + col0), "r"(ncols * 2) : "memory");
  };
  auto prefetch_wqkv = [&](int l) {
    if (cta < 3 * D / 32) prefetch_rows(a.w_qkv, (size_t)l * 3 * D + cta * 32, 32, D, 0, D);
  };
  auto prefetch_splitk = [&](const __nv_bfloat16* w, int l, int K, int ksplit) {  // wo / w2 tile of this CTA
    const int nt = D / 64;
    if (cta >= nt * ksplit) return;
    const int split = cta / nt, kb = K / kBlockK, k0 = kb * split / ksplit, k1 = kb * (split + 1) / ksplit;
    prefetch_rows(w, (size_t)l * D + (cta % nt) * 64, 64, K, (size_t)k0 * kBlockK, (k1 - k0) * kBlockK);
  };
  auto prefetch_w13 = [&](int l) {
    if (cta < 2 * F / 64) prefetch_rows(a.w_13, (size_t)l * 2 * F + cta * 64, 64, D, 0, D);
  };
  auto prefetch_heads = [&]() {
    for (int t = cta; t < a.NH / 64; t += G) prefetch_rows(a.w_heads, (size_t)t * 64, 64, D, 0, D);
  };

  // where a wo / w2 tile of K slice `split` puts its 64 x 64 sums: its own slice of a.part with plain stores (the row CTAs
  // add the slices in order, rmsnorm_phase), or - a.part == nullptr, the round-1 scheme kept for A/B runs - straight into h
  // with vector float reductions, whose order is not fixed
  auto resid_target = [&](int split) {
    ep.N = D; ep.ldo = D; ep.perm_S = 0;
    if constexpr (DET) { ep.mode = EPI_STORE; ep.out_f32 = a.part + (size_t)split * R * D; ep.atomic = 0; }
    else { ep.mode = EPI_RESID; ep.out_f32 = a.h; ep.atomic = 1; }
  };

  prefetch_wqkv(0);
  rmsnorm_phase(a.attn_norm, a.fuse_io);
  sync_all();
  for (int l = 0; l < a.L; ++l) {
    // wqkv: 144 tiles of 32 output features, RoPE + KV append + bf16 q in the epilogue
    prefetch_splitk(a.w_o, l, D, a.wo_ksplit);
    if (cta < 3 * D / 32) {
      ep.mode = EPI_QKV; ep.N = 3 * D; ep.out_bf16 = a.q; ep.out_f32 = nullptr; ep.ldo = D; ep.layer = l; ep.atomic = 0;
      fused_gemm_tile<32, TM>(pp, &tm_xn, &tm_wqkv, l, cta * 32, 0, D / kBlockK, ep,
                              a.timing && cta == a.timing_cta && l == a.L - 1 ? a.timing + 910 : nullptr);
    }
    sync_all();
    attention_phase(l);
    sync_all();
    // wo + residual: 24 N tiles x wo_ksplit K slices, fp32 vector reductions into h; then the FFN norm of the rows
    {
      const int nt = D / 64;
      prefetch_w13(l);
      if (cta < (int)tiles_wo) {
        const int split = cta / nt, kb = D / kBlockK;
        resid_target(split);
        fused_gemm_tile<64, TM>(pp, &tm_attn, &tm_wo, l, (cta % nt) * 64, kb * split / a.wo_ksplit, kb * (split + 1) / a.wo_ksplit, ep);
        tiles_arrive();
      }
      tiles_seen += tiles_wo;
      if (cta < R) {
        tiles_wait();
        rmsnorm_phase(a.ffn_norm + (size_t)l * D, false, DET ? a.wo_ksplit : 0);
      }
    }
    sync_all();
    // w1|w3 (rows interleaved) + SiLU * mul: 128 tiles of 64 rows = 32 hidden units
    prefetch_splitk(a.w_2, l, F, a.w2_ksplit);
    if (cta < 2 * F / 64) {
      ep.mode = EPI_SWIGLU; ep.N = 2 * F; ep.out_bf16 = a.act; ep.ldo = F; ep.atomic = 0;
      fused_gemm_tile<64, TM>(pp, &tm_xn, &tm_w13, l, cta * 64, 0, D / kBlockK, ep,
                              a.timing && cta == a.timing_cta && l == a.L - 1 ? a.timing + 940 : nullptr);
    }
    sync_all();
    // w2 + residual; then the next layer's attention norm (the final norm after the last layer)
    {
      const int nt = D / 64;
      if (l + 1 < a.L) prefetch_wqkv(l + 1); else prefetch_heads();
      if (cta < (int)tiles_w2) {
        const int split = cta / nt, kb = F / kBlockK;
        resid_target(split);
        fused_gemm_tile<64, TM>(pp, &tm_act, &tm_w2, l, (cta % nt) * 64, kb * split / a.w2_ksplit, kb * (split + 1) / a.w2_ksplit, ep);
        tiles_arrive();
      }
      tiles_seen += tiles_w2;
      if (cta < R) {
        tiles_wait();
        rmsnorm_phase(l + 1 < a.L ? a.attn_norm + (size_t)(l + 1) * D : a.final_norm, false, DET ? a.w2_ksplit : 0);
      }
    }
    sync_all();
  }
  // heads: NH / 64 tiles (144 for 9 x 1024), plain fp32 store
  for (int t = cta; t < a.NH / 64; t += G) {
    ep.mode = EPI_STORE; ep.N = a.NH; ep.out_f32 = a.logits; ep.ldo = a.NH; ep.atomic = 0;
    fused_gemm_tile<64, TM>(pp, &tm_xn, &tm_heads, 0, t * 64, 0, D / kBlockK, ep);
    __syncthreads();  // the accumulator is single-buffered: drain it before the next tile's MMAs
  }
  if (a.fuse_io) {
    // CFG / sampling / mask-fix / write-back (vaura_model.py:775-827, what sample_kernel does): one warp per (clip, codebook)
    sync_all();
    const int nrows = a.sample.B * a.sample.K;
    for (int u = warp * G + cta; u < nrows; u += G * (kGemmThreads / 32)) sample_row(a.sample, u / a.sample.K, u % a.sample.K, lane, offset);
  }
  if (cta == 0 && tid == 0) {  // every CTA read offset / epoch before its first barrier arrival
    a.state->epoch = epoch + 1;
    if (a.fuse_io) a.state->offset = offset + 1;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pp.tmem_base), "r"(kAccCols));
}

// One-time per-device set-up of a kernel (dynamic shared-memory opt-in is a per-device function attribute): a process that
// drives several GPUs must repeat it on each of them.  Returns the device ordinal clamped into the table.
static int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D K-major map: dims {K, rows, outer}; box {block_k, box_rows, 1}; 16-bit elements
static bool make_map(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t outer, uint64_t row_stride_el,
                     uint64_t outer_stride_el, int block_k, int box_rows, bool f16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {K, rows, outer};
  cuuint64_t strides[2] = {row_stride_el * 2, outer_stride_el * 2};
  cuuint32_t box[3] = {(cuuint32_t)block_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [outer][rows][K] bf16 seen as (64, rows, K / 64, outer): one box = nblk K blocks of box_rows rows, landing in shared
// memory as nblk consecutive 128B-swizzled [box_rows x 64] sub-tiles
static bool make_map_kblocks(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t outer, uint64_t row_stride_el,
                             uint64_t outer_stride_el, int box_rows, int nblk, int block_k = 64, bool f16 = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || K % block_k || (block_k != 64 && block_k != 32)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)block_k, rows, K / block_k, outer};
  cuuint64_t strides[3] = {row_stride_el * 2, (cuuint64_t)block_k * 2, outer_stride_el * 2};
  cuuint32_t box[4] = {(cuuint32_t)block_k, (cuuint32_t)box_rows, (cuuint32_t)nblk, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc_make_map_kblocks(void* map, const void* base, uint64_t K, uint64_t rows, uint64_t outer, uint64_t row_stride_el,
                         uint64_t outer_stride_el, int box_rows, int nblk) {
  return make_map_kblocks(static_cast<CUtensorMap*>(map), base, K, rows, outer, row_stride_el, outer_stride_el, box_rows, nblk);
}

template <int BLOCK_N, int BLOCK_K, int STAGES, int FMT, class Epi, int TILE_M = 128, int KSUB = 1, int TERMS = 1, int CK = 1>
static cudaError_t launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcShape& g, const typename Epi::Params& ep,
                             int m_tiles, int n_tiles, cudaStream_t st) {
  constexpr int smem = STAGES * KSUB * (TERMS * TILE_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + 1024 + 256;
  static_assert(smem <= 227 * 1024, "stage ring exceeds shared memory");
  if (KSUB > 1 && g.ntaps != 1) return cudaErrorInvalidValue;
  if (CK > 1 && g.ksplit != CK) return cudaErrorInvalidValue;
  auto kern = gemm_tc_kernel<BLOCK_N, BLOCK_K, STAGES, FMT, Epi, TILE_M, KSUB, TERMS, CK>;
  static bool attr[64] = {false};
  const int slot = current_device_slot();
  if (!attr[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr[slot] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m_tiles, n_tiles, g.batch * g.nphase * g.ksplit);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr1[2];
  int na = 0;
  if (g.pdl) {
    attr1[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr1[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CK > 1) {
    attr1[na].id = cudaLaunchAttributeClusterDimension;
    attr1[na].val.clusterDim.x = 1; attr1[na].val.clusterDim.y = 1; attr1[na].val.clusterDim.z = CK;
    ++na;
  }
  cfg.attrs = attr1;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, g, ep);
}
template <int BLOCK_N, int BLOCK_K, int STAGES, int FMT, class Epi, int KSUB = 1, int OCC = 1, int EW = 8>
static cudaError_t launch_tc_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const TcShape& g,
                                        const typename Epi::Params& ep, int m_tiles, int n_tiles, cudaStream_t st) {
  constexpr int smem = STAGES * KSUB * (kTileM * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + 1024 + 256;
  static_assert(OCC * (smem + 1024) <= 227 * 1024, "stage ring exceeds shared memory");
  if (g.kblocks % KSUB) return cudaErrorInvalidValue;
  auto kern = gemm_tc_persistent_kernel<BLOCK_N, BLOCK_K, STAGES, FMT, Epi, KSUB, OCC, EW>;
  static int sms_tab[64] = {0};
  const int slot = current_device_slot();
  if (!sms_tab[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_tab[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  const int sms = sms_tab[slot];
  const int ntiles = m_tiles * n_tiles * g.batch * g.nphase;
  kern<<<dim3(ntiles < OCC * sms ? ntiles : OCC * sms), dim3(64 + 32 * EW), smem, st>>>(ta, tb, g, ep, m_tiles, n_tiles);
  return cudaGetLastError();
}

template <int C, int BLOCK_K, int STAGES>
static cudaError_t launch_ru_fused_t(const RuArgs& a, const int* taps7_host, int B, cudaStream_t st) {
  constexpr int smem = STAGES * (kTileM * BLOCK_K * 2 + C * BLOCK_K * 2) + (C / BLOCK_K) * kTileM * BLOCK_K * 2 + 3 * C * 4 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "ring + h tile exceed shared memory");
  CUtensorMap ta, t7, t1;
  if (!make_map(&ta, a.act, C, a.T, B, C, (uint64_t)a.T * C, BLOCK_K, kTileM, true) ||
      !make_map(&t7, a.W7, C, C, 7, C, (uint64_t)C * C, BLOCK_K, C, true) ||
      !make_map(&t1, a.W1, C, C, 1, C, (uint64_t)C * C, BLOCK_K, C, true))
    return cudaErrorUnknown;
  auto kern = gemm_ru_fused_kernel<C, BLOCK_K, STAGES>;
  static int sms_tab[64] = {0};
  const int slot = current_device_slot();
  if (!sms_tab[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_tab[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  TcShape g{};
  g.ntaps = 7; g.nphase = 1; g.kblocks = C / BLOCK_K; g.batch = B; g.ksplit = 1;
  for (int i = 0; i < 7; ++i) g.tap_off[i] = taps7_host[i];
  RuParams rp{};
  rp.bias7 = a.bias7; rp.alpha2 = a.alpha2;
  rp.k1 = EpiConv::Params{a.bias1, a.alpha_next, a.x, a.out_raw, a.out_act, a.T, a.T, C, 1};
  const int mt = (a.T + kTileM - 1) / kTileM, ntiles = mt * B, sms = sms_tab[slot];
  kern<<<dim3(ntiles < sms ? ntiles : sms), dim3(kGemmThreads), smem, st>>>(ta, t7, t1, g, rp, mt);
  return cudaGetLastError();
}

#ifdef VAURA_RU_TIMING
extern "C" int vaura_debug_ru_timing(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ru_timing, sizeof(g_ru_timing));
}
#endif

template <int C, int BLOCK_K, int STAGES>
static cudaError_t launch_ru_fused_skew_t(const RuArgs& a, const int* taps7_host, int B, cudaStream_t st) {
  constexpr int KB = C / BLOCK_K;
  constexpr int smem = STAGES * (kTileM * BLOCK_K * 2 + C * BLOCK_K * 2) + 2 * KB * kTileM * BLOCK_K * 2 + 3 * C * 4 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "ring + two h tiles exceed shared memory");
  CUtensorMap ta, t7, t1;
  if (!make_map(&ta, a.act, C, a.T, B, C, (uint64_t)a.T * C, BLOCK_K, kTileM, true) ||
      !make_map(&t7, a.W7, C, C, 7, C, (uint64_t)C * C, BLOCK_K, C, true) ||
      !make_map(&t1, a.W1, C, C, 1, C, (uint64_t)C * C, BLOCK_K, C, true))
    return cudaErrorUnknown;
  auto kern = gemm_ru_fused_skew_kernel<C, BLOCK_K, STAGES>;
  static int sms_tab[64] = {0};
  const int slot = current_device_slot();
  if (!sms_tab[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_tab[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  TcShape g{};
  g.ntaps = 7; g.nphase = 1; g.kblocks = KB; g.batch = B; g.ksplit = 1;
  for (int i = 0; i < 7; ++i) g.tap_off[i] = taps7_host[i];
  RuParams rp{};
  rp.bias7 = a.bias7; rp.alpha2 = a.alpha2;
  rp.k1 = EpiConv::Params{a.bias1, a.alpha_next, a.x, a.out_raw, a.out_act, a.T, a.T, C, 1};
  const int mt = (a.T + kTileM - 1) / kTileM, ntiles = mt * B, sms = sms_tab[slot];
  kern<<<dim3(ntiles < sms ? ntiles : sms), dim3(kGemmThreads), smem, st>>>(ta, t7, t1, g, rp, mt);
  return cudaGetLastError();
}

bool ru_fused_supported(int C) { return C == 192 || C == 96 || C == 128 || C == 64 || C == 256; }

// One launch per ResidualUnit (see gemm_ru_fused_kernel)
cudaError_t launch_ru_fused(const RuArgs& a, const int* taps7_host, int B, cudaStream_t st) {
  // The variant skewed by one tile (two acc1 buffers, two h tiles) for the encoder's 64- and 128-channel units: DAC encode of
  // 64 clips 21.7 -> 21.0 ms.  The decoder's 96-channel units gain 3-4 % at dilation 1 and 3 and lose 5 % at dilation 9
  // (629 / 618 / 570 us against 646 / 645 / 542 us per launch, profiles/scripts/r02_run76.sh): they stay on the kernel above.
  if (knobs().codec_ru_skew) {
    switch (a.C) {
      case 128: return launch_ru_fused_skew_t<128, 64, 4>(a, taps7_host, B, st);
      case 64: return launch_ru_fused_skew_t<64, 64, 8>(a, taps7_host, B, st);
    }
  }
  switch (a.C) {
    case 256: return launch_ru_fused_t<256, 64, 3>(a, taps7_host, B, st);
    case 192: return launch_ru_fused_t<192, 64, 4>(a, taps7_host, B, st);
    case 128: return launch_ru_fused_t<128, 64, 5>(a, taps7_host, B, st);
    case 96: return launch_ru_fused_t<96, 32, 10>(a, taps7_host, B, st);
    case 64: return launch_ru_fused_t<64, 64, 8>(a, taps7_host, B, st);
  }
  return cudaErrorInvalidValue;
}

bool conv_tc_supported(int Cin, int Cout, int ntaps, int nphase) {
  if (Cin % 32 != 0 || Cout % 16 != 0 || ntaps * nphase > 32) return false;
  const int bk = (Cin % 64 == 0) ? 64 : 32;
  int bn = Cout;
  if (bn > 256) bn = (Cout % 256 == 0) ? 256 : ((Cout % 192 == 0) ? 192 : 128);
  if (Cout % bn != 0) return false;
  const int ok[][2] = {{256, 64}, {192, 64}, {128, 64}, {96, 64}, {64, 64}, {32, 64}, {96, 32}, {32, 32}, {16, 32}};
  for (auto& c : ok)
    if (c[0] == bn && c[1] == bk) return true;
  return false;
}

// codec conv as implicit GEMM (see ConvArgs in kernels.h)
cudaError_t launch_conv_tc(const ConvArgs& a, const int* tap_off_host, int B, cudaStream_t st) {
  const int bk = (a.Cin % 64 == 0) ? 64 : 32;
  if (a.Cin % bk != 0 || a.Cout % 16 != 0 || a.ntaps * a.nphase > 32) return cudaErrorInvalidValue;
  int bn = a.Cout;
  if (bn > 256) bn = (a.Cout % 256 == 0) ? 256 : ((a.Cout % 192 == 0) ? 192 : 128);
  if (a.Cout % bn != 0) return cudaErrorInvalidValue;
  CUtensorMap ta, tb;
  if (!make_map(&ta, a.in, a.Cin, a.Tin, B, a.Cin, (uint64_t)a.Tin * a.Cin, bk, kTileM, true)) return cudaErrorUnknown;
  if (!make_map(&tb, a.W, a.Cin, a.Cout, (uint64_t)a.ntaps * a.nphase, a.Cin, (uint64_t)a.Cout * a.Cin, bk, bn, true))
    return cudaErrorUnknown;
  TcShape g{};
  g.ntaps = a.ntaps; g.nphase = a.nphase; g.kblocks = a.Cin / bk; g.batch = B; g.ksplit = 1; g.pdl = 0;
  for (int i = 0; i < a.ntaps * a.nphase; ++i) g.tap_off[i] = tap_off_host[i];
  EpiConv::Params ep{a.bias, a.alpha, a.residual, a.out_raw, a.out_act, a.Tq, a.Tout, a.Cout, a.ostride};
  const int mt = (a.Tq + kTileM - 1) / kTileM, nt = a.Cout / bn;
  // persistent tile loop (double-buffered TMEM accumulator) unless VAURA_CONV_PERSISTENT=0
  const bool persistent = knobs().conv_persistent, ksub3 = knobs().conv_ksub;
  // 96 input channels = three 32-wide K blocks per tap: one stage (one tensor box per operand) per tap
  if (persistent && ksub3 && bk == 32 && a.Cin == 96 && (bn == 96 || bn == 32 || bn == 16)) {
    CUtensorMap ta4, tb4;
    if (!make_map_kblocks(&ta4, a.in, a.Cin, a.Tin, B, a.Cin, (uint64_t)a.Tin * a.Cin, kTileM, 3, 32, true) ||
        !make_map_kblocks(&tb4, a.W, a.Cin, a.Cout, (uint64_t)a.ntaps * a.nphase, a.Cin, (uint64_t)a.Cout * a.Cin, bn, 3, 32, true))
      return cudaErrorUnknown;
    // VAURA_CONV_OCC2=1: two CTAs per SM for the 96-channel layers (measured neutral: 9.58 vs 9.54 ms per 16 clips - these
    // layers are bound by bytes in flight per SM, not by the per-CTA serial chain)
    const bool occ2 = knobs().conv_occ2;
    if (bn == 96 && occ2) return launch_tc_persistent<96, 32, 2, 0, EpiConv, 3, 2>(ta4, tb4, g, ep, mt, nt, st);
    if (bn == 96) return launch_tc_persistent<96, 32, 4, 0, EpiConv, 3>(ta4, tb4, g, ep, mt, nt, st);
    if (bn == 32) return launch_tc_persistent<32, 32, 4, 0, EpiConv, 3>(ta4, tb4, g, ep, mt, nt, st);
    return launch_tc_persistent<16, 32, 4, 0, EpiConv, 3>(ta4, tb4, g, ep, mt, nt, st);
  }
#define TC_CASE(BN, BK, ST)                                                                                    \
  if (bn == BN && bk == BK)                                                                                    \
    return persistent ? launch_tc_persistent<BN, BK, ST, 0, EpiConv>(ta, tb, g, ep, mt, nt, st)                \
                      : launch_tc<BN, BK, ST, 0, EpiConv>(ta, tb, g, ep, mt, nt, st);
  {
    if (persistent && knobs().conv_occ2 && bn == 96 && bk == 64) return launch_tc_persistent<96, 64, 3, 0, EpiConv, 1, 2>(ta, tb, g, ep, mt, nt, st);
  }
  TC_CASE(256, 64, 4)
  TC_CASE(192, 64, 4)
  TC_CASE(128, 64, 6)
  TC_CASE(96, 64, 6)
  TC_CASE(64, 64, 6)
  TC_CASE(32, 64, 6)
  TC_CASE(96, 32, 8)
  TC_CASE(32, 32, 8)
  TC_CASE(16, 32, 8)
#undef TC_CASE
  return cudaErrorInvalidValue;
}

// Fused decode step (decode_step_fused_bf16): tensor maps are rebuilt per call (they are baked into the captured graph node)
bool fused_step_supported(int R, int D, int F, int NH) {
  return R >= 1 && R <= 128 && D % 64 == 0 && F % 64 == 0 && NH % 64 == 0 && (3 * D) % 32 == 0 && D / 4 <= 2 * kGemmThreads;
}

size_t fused_part_bytes(int R, int D) { return (size_t)kFusedKsplit * R * D * sizeof(float); }

template <int TM, bool DET>
static cudaError_t launch_decode_fused_t(const FusedStepArgs& a, const void* wqkv, const void* wo, const void* w13, const void* w2,
                                         const void* w_heads, cudaStream_t st) {
  static int sms_tab[64] = {0};
  const int slot = current_device_slot();
  if (!sms_tab[slot]) {
    cudaError_t e = cudaFuncSetAttribute(decode_step_fused_bf16<TM, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::kSmem);
    if (e != cudaSuccess) return e;
    int dev = 0, occ = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_step_fused_bf16<TM, DET>, kGemmThreads, fused::kSmem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    sms_tab[slot] = n;
  }
  const int sms = sms_tab[slot];
  const int need = a.D / 64 * (a.wo_ksplit > a.w2_ksplit ? a.wo_ksplit : a.w2_ksplit);
  if (a.wo_ksplit < 1 || a.w2_ksplit < 1 || (a.part && (a.wo_ksplit > kFusedKsplit || a.w2_ksplit > kFusedKsplit)))
    return cudaErrorInvalidValue;
  // one tile per CTA per phase, one residual row per CTA in the norm phases
  if (3 * a.D / 32 > sms || 2 * a.F / 64 > sms || need > sms || a.R > sms) return cudaErrorInvalidValue;
  // attention phase: at most two (row, head) items per warp, lane i holds page i, 16-position runs inside a page
  if (a.R * a.H > 2 * sms * (kGemmThreads / 32) || a.kv.max_pages_per_seq > 32 || a.kv.page_size < 16 ||
      (a.kv.page_size & (a.kv.page_size - 1)))
    return cudaErrorInvalidValue;
  const uint64_t D = a.D, F = a.F, L = a.L;
  CUtensorMap m_xn, m_attn, m_act, m_wqkv, m_wo, m_w13, m_w2, m_heads;
  constexpr int KS = fused::kKsub;
  bool ok = make_map_kblocks(&m_xn, a.xn, D, a.R, 1, D, (uint64_t)a.R * D, TM, KS) &&
            make_map_kblocks(&m_attn, a.attn, D, a.R, 1, D, (uint64_t)a.R * D, TM, KS) &&
            make_map_kblocks(&m_act, a.act, F, a.R, 1, F, (uint64_t)a.R * F, TM, KS) &&
            make_map_kblocks(&m_wqkv, wqkv, D, 3 * D, L, D, 3 * D * D, 32, KS) &&
            make_map_kblocks(&m_wo, wo, D, D, L, D, D * D, 64, KS) &&
            make_map_kblocks(&m_w13, w13, D, 2 * F, L, D, 2 * F * D, 64, KS) &&
            make_map_kblocks(&m_w2, w2, F, D, L, F, D * F, 64, KS) &&
            make_map_kblocks(&m_heads, w_heads, D, a.NH, 1, D, (uint64_t)a.NH * D, 64, KS);
  if (!ok) return cudaErrorUnknown;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = fused::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  FusedStepArgs b = a;
  b.w_qkv = static_cast<const __nv_bfloat16*>(wqkv);
  b.w_o = static_cast<const __nv_bfloat16*>(wo);
  b.w_13 = static_cast<const __nv_bfloat16*>(w13);
  b.w_2 = static_cast<const __nv_bfloat16*>(w2);
  b.w_heads = static_cast<const __nv_bfloat16*>(w_heads);
  b.l2_prefetch = knobs().fused_l2_prefetch;
  return cudaLaunchKernelEx(&cfg, decode_step_fused_bf16<TM, DET>, m_xn, m_attn, m_act, m_wqkv, m_wo, m_w13, m_w2, m_heads, b);
}

cudaError_t launch_decode_fused_bf16(const FusedStepArgs& a, const void* wqkv, const void* wo, const void* w13, const void* w2,
                                     const void* w_heads, cudaStream_t st) {
  if (!fused_step_supported(a.R, a.D, a.F, a.NH)) return cudaErrorInvalidValue;
  const bool tm64 = a.R <= 64 && !knobs().fused_tm128;
  if (a.part) return tm64 ? launch_decode_fused_t<64, true>(a, wqkv, wo, w13, w2, w_heads, st)
                          : launch_decode_fused_t<128, true>(a, wqkv, wo, w13, w2, w_heads, st);
  return tm64 ? launch_decode_fused_t<64, false>(a, wqkv, wo, w13, w2, w_heads, st)
              : launch_decode_fused_t<128, false>(a, wqkv, wo, w13, w2, w_heads, st);
}

static int mt_for(int R) { return (R + kTileM - 1) / kTileM; }

// Linear layer of the bf16 sampler path: out = A[R][K] (bf16) x W[N][K]^T (bf16) with a fused epilogue.
cudaError_t launch_linear_tc(const LinearTcArgs& a, cudaStream_t st) {
  if (a.K % 64 != 0 || a.N % a.block_n != 0) return cudaErrorInvalidValue;
  const int wk = a.w_k > 0 ? a.w_k : a.K;  // K extent of W (A may hold several bf16 terms side by side)
  if (wk % 64 != 0 || a.K % wk != 0) return cudaErrorInvalidValue;
  const bool split3 = a.K == 3 * wk;       // three terms: one weight tile per K block multiplied with the three A tiles
  // up to 64 rows: UMMA M=64 halves the activation tile, so twice as many weight bytes fit in flight per SM
  const int tile_m = (a.R <= 64 && !split3) ? 64 : kTileM;
  CUtensorMap ta, tb;
  if (!make_map(&ta, a.A, a.K, a.R, 1, a.lda, (uint64_t)a.R * a.lda, 64, tile_m, false)) return cudaErrorUnknown;
  if (!make_map(&tb, a.W, wk, a.N, 1, wk, (uint64_t)a.N * wk, 64, a.block_n, false)) return cudaErrorUnknown;
  TcShape g{};
  g.ntaps = 1; g.nphase = 1; g.kblocks = split3 ? wk / 64 : a.K / 64; g.batch = 1;
  g.bwrap = wk != a.K ? wk / 64 : 0;
  g.ksplit = (a.epi == EPI_RESID && a.ksplit > 1) ? a.ksplit : 1;
  if (g.ksplit > g.kblocks) g.ksplit = g.kblocks;
  g.pdl = a.pdl;
  EpiLinear::Params ep{};
  ep.mode = a.epi; ep.R = a.R; ep.N = a.N; ep.out_f32 = a.out_f32; ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  ep.ldo = a.ldo; ep.perm_S = a.perm_S; ep.perm_V = a.perm_V; ep.rope = a.rope; ep.kv = a.kv; ep.state = a.state;
  ep.pos0 = a.pos0; ep.npos = a.npos; ep.layer = a.layer; ep.d_model = a.d_model; ep.atomic = g.ksplit > 1;
  ep.aux = a.aux;
  const int mt = (a.R + tile_m - 1) / tile_m, nt = a.N / a.block_n;
  if (split3 && a.ksplit == 0 && (a.epi != EPI_STORE || a.perm_S == 0)) {
    // auto: when the tiles of this GEMM cover less than half of the SMs (prompt prefill: one or two row tiles), split K over
    // clusters of 4 or 2 CTAs (partial sums through DSMEM, rank-ordered: deterministic) and widen the 1536-wide tiles
    const bool ck_on = knobs().prefill_ck;
    // 256-wide tiles from N = 8192 (w1|w3); q|k|v (N = 4608) runs 128-wide tiles in pairs = 144 CTAs instead of 72: prefill of a
    // 167-position window 3.59 -> 3.33 ms (128-wide for w1|w3 as well: 3.51).  VAURA_PREFILL_BN256_FROM overrides.
    const int bn_thr = knobs().prefill_bn256_from;
    const int bn = a.N % 256 == 0 && a.N >= bn_thr ? 256 : (a.N % 128 == 0 ? 128 : 0);
    const int kb = wk / 64;
    if (ck_on && bn) {
      const int tiles = mt_for(a.R) * (a.N / bn);
      // a B200 co-schedules 33 clusters of 4 CTAs that need a whole SM each (profiles/probes/cluster_occ.cu): more would run
      // as a second wave
      const int ck = (tiles <= 32 && kb >= 8) ? 4 : ((tiles * 2 <= 148 && kb >= 4) ? 2 : 1);
      if (ck > 1) {
        CUtensorMap tb2;
        if (!make_map(&tb2, a.W, wk, a.N, 1, wk, (uint64_t)a.N * wk, 64, bn, false)) return cudaErrorUnknown;
        g.ksplit = ck;
        ep.atomic = 0;
        const int nt2 = a.N / bn;
        if (bn == 256 && ck == 4) return launch_tc<256, 64, 2, 1, EpiLinear, 128, 1, 3, 4>(ta, tb2, g, ep, mt, nt2, st);
        if (bn == 256 && ck == 2) return launch_tc<256, 64, 2, 1, EpiLinear, 128, 1, 3, 2>(ta, tb2, g, ep, mt, nt2, st);
        if (bn == 128 && ck == 4) return launch_tc<128, 64, 3, 1, EpiLinear, 128, 1, 3, 4>(ta, tb2, g, ep, mt, nt2, st);
        return launch_tc<128, 64, 3, 1, EpiLinear, 128, 1, 3, 2>(ta, tb2, g, ep, mt, nt2, st);
      }
    }
  }
  if (split3) {
    if (g.ksplit != 1) return cudaErrorInvalidValue;
    switch (a.block_n) {  // stage = 3 x 16 KB of A + the weight tile
      case 32: return launch_tc<32, 64, 4, 1, EpiLinear, 128, 1, 3>(ta, tb, g, ep, mt, nt, st);
      case 64: return launch_tc<64, 64, 3, 1, EpiLinear, 128, 1, 3>(ta, tb, g, ep, mt, nt, st);
      case 128: return launch_tc<128, 64, 3, 1, EpiLinear, 128, 1, 3>(ta, tb, g, ep, mt, nt, st);
      case 256: return launch_tc<256, 64, 2, 1, EpiLinear, 128, 1, 3>(ta, tb, g, ep, mt, nt, st);
    }
    return cudaErrorInvalidValue;
  }
  if (tile_m == 64) {
    switch (a.block_n) {
      case 16: return launch_tc<16, 64, 4, 1, EpiLinear, 64, 4>(ta, tb, g, ep, mt, nt, st);
      case 32: return launch_tc<32, 64, 4, 1, EpiLinear, 64, 4>(ta, tb, g, ep, mt, nt, st);
      case 64: return launch_tc<64, 64, 3, 1, EpiLinear, 64, 4>(ta, tb, g, ep, mt, nt, st);
      case 128: return launch_tc<128, 64, 2, 1, EpiLinear, 64, 4>(ta, tb, g, ep, mt, nt, st);
    }
    return cudaErrorInvalidValue;
  }
  switch (a.block_n) {
    case 16: return launch_tc<16, 64, 8, 1, EpiLinear>(ta, tb, g, ep, mt, nt, st);
    case 32: return launch_tc<32, 64, 8, 1, EpiLinear>(ta, tb, g, ep, mt, nt, st);
    case 64: return launch_tc<64, 64, 6, 1, EpiLinear>(ta, tb, g, ep, mt, nt, st);
    case 128: return launch_tc<128, 64, 6, 1, EpiLinear>(ta, tb, g, ep, mt, nt, st);
  }
  return cudaErrorInvalidValue;
}

// Linear layer of the Segment-AVCLIP tower: out = A[M][K] (bf16) x W[N][K]^T (bf16) + bias, epilogue per VitLinearArgs::mode.
// Persistent tile loop with a double-buffered TMEM accumulator (gemm_tc_persistent_kernel): M is 10^4..10^6 token rows, so
// the epilogue of one 128 x BLOCK_N tile overlaps the mainloop of the next.
cudaError_t launch_vit_linear(const VitLinearArgs& a, cudaStream_t st) {
  if (a.K % 64 != 0 || a.M <= 0) return cudaErrorInvalidValue;
  const int bn = a.N % 256 == 0 ? 256 : (a.N % 128 == 0 ? 128 : 0);
  if (!bn) return cudaErrorInvalidValue;
  CUtensorMap ta, tb;
  if (!make_map(&ta, a.A, a.K, a.M, 1, a.lda, (uint64_t)a.M * a.lda, 64, kTileM, false)) return cudaErrorUnknown;
  if (!make_map(&tb, a.W, a.K, a.N, 1, a.K, (uint64_t)a.N * a.K, 64, bn, false)) return cudaErrorUnknown;
  TcShape g{};
  g.ntaps = 1; g.nphase = 1; g.kblocks = a.K / 64; g.batch = 1; g.ksplit = 1; g.pdl = 0;
  g.n_fastest = !knobs().avclip_m_fastest;  // M fastest = the codec's tile order (A/B measurement)
  EpiVit::Params ep{};
  ep.mode = a.mode; ep.gelu = a.gelu; ep.M = a.M; ep.N = a.N; ep.ldo = a.ldo; ep.bias = a.bias;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16); ep.out_f32 = a.out_f32; ep.pos = a.pos;
  ep.rows_in = a.rows_in; ep.rows_out = a.rows_out; ep.row_off = a.row_off;
  const int mt = (a.M + kTileM - 1) / kTileM, nt = a.N / bn;
  // CTA-pair tiles (gemm_tc2_persistent_kernel) by default: q|k|v 187 -> 171 us, fc1 358 -> 245 us, fc2 212 -> 192 us per 32
  // segments (12.96 vs 15.19 ms of GEMMs per forward).  VAURA_AVCLIP_2CTA=0: one CTA per tile (A/B measurement)
  const bool two_cta = knobs().avclip_2cta;
  if (two_cta && bn == 256) {
    constexpr int ST = 6, EWP = 16;
    constexpr int smem2 = ST * (kTileM * 64 * 2 + 128 * 64 * 2) + 1024 + 256;
    CUtensorMap tb2;
    if (!make_map(&tb2, a.W, a.K, a.N, 1, a.K, (uint64_t)a.N * a.K, 64, 128, false)) return cudaErrorUnknown;
    auto kern = gemm_tc2_persistent_kernel<256, ST, 1, EpiVit, EWP>;
    static int sms_tab[64] = {0};
    const int slot = current_device_slot();
    if (!sms_tab[slot]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
      if (e != cudaSuccess) return e;
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms_tab[slot], cudaDevAttrMultiProcessorCount, dev);
    }
    const int mt2 = (a.M + 2 * kTileM - 1) / (2 * kTileM);
    int grid = sms_tab[slot] & ~1;
    if (2 * mt2 * nt < grid) grid = 2 * mt2 * nt;
    kern<<<dim3(grid), dim3(64 + 32 * EWP), smem2, st>>>(ta, tb2, g, ep, mt2, nt);
    return cudaGetLastError();
  }
  const bool ew16 = !knobs().avclip_ew8;  // eight epilogue warps: A/B measurement
  if (bn == 256) return ew16 ? launch_tc_persistent<256, 64, 4, 1, EpiVit, 1, 1, 16>(ta, tb, g, ep, mt, nt, st)
                             : launch_tc_persistent<256, 64, 4, 1, EpiVit>(ta, tb, g, ep, mt, nt, st);
  return launch_tc_persistent<128, 64, 6, 1, EpiVit>(ta, tb, g, ep, mt, nt, st);
}

}  // namespace vaura
