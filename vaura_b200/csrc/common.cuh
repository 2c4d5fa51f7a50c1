// Shared device helpers for the vaura_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaura {

constexpr int kHeadDim = 96;   // d_model / nhead of the shipped model (llama_9cbs.yaml); kernels are built for it
constexpr int kMaxCtx = 256;   // RoPE table rows == max context (llama.py:317, :364-368)
constexpr int kUnknown = -1;   // "not generated yet" marker (vaura_model.py:482)

// Device-resident loop state so that one captured CUDA graph can be replayed for every decode step.
struct StepState {
  int offset;    // column about to be sampled; the step consumes columns [offset - npos, offset)
  int done;      // arrival counter of the sampling kernel's CTAs
  unsigned epoch;    // persistent kernel: launches since the state was (re)initialised
  unsigned barrier;  // persistent kernel: monotonically increasing device-wide barrier counter
  int pad0[28];
  unsigned tiles_done;  // fused step kernel: split-K tiles finished (own 128-byte line; see decode_step_fused_bf16)
  int pad1[31];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// two bf16 packed in a 32-bit word -> fp32 (exact: bf16 is the top half of an fp32)
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// x = t1 + t2 + t3 with bf16 terms: 3 x 8 significant bits cover the 24 of an fp32, so bf16 x bf16 products of the terms
// accumulated in fp32 reproduce the fp32 product (what decode_cluster.cu does with mma.sync, here for tcgen05 prefill)
__device__ __forceinline__ void split3(float x, __nv_bfloat16& t1, __nv_bfloat16& t2, __nv_bfloat16& t3) {
  t1 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(t1);
  t2 = __float2bfloat16_rn(r1);
  t3 = __float2bfloat16_rn(r1 - __bfloat162float(t2));
}

// 16-byte streaming load: weights are read exactly once per step, keep them out of L1.
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Paged KV addressing (include/vaura_b200.h: vaura_kv_cache).
struct KvView {
  void* pages;
  const int32_t* page_table;
  int num_pages, page_size, max_pages_per_seq, nhead;
  // element offset of (layer, kv, seq b, position p, head h, dim 0)
  __device__ __forceinline__ size_t row(int layer, int kv, int b, int p, int h) const {
    int page = page_table[b * max_pages_per_seq + p / page_size];
    return ((((size_t)(layer * 2 + kv) * num_pages + page) * nhead + h) * page_size + (p % page_size)) * kHeadDim;
  }
};

// Philox4x32-10 (Salmon et al., SC'11).  counter = (clip_id, offset, codebook, stream_id), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

}  // namespace vaura
