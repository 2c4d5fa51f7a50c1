// CFG combine -> temperature -> softmax -> top-k / top-p filter -> Philox inverse-CDF draw (or argmax)
// -> delay-pattern mask-fix -> prompt-preserving write-back.  One warp per (clip, codebook) row of
// V = 1024 logits; every selection is done with warp shuffles / votes on registers.
//
// Reference code replaced (paths relative to /root/reference):
//   CFG combine            models/vaura_model.py:810-813   (uncond + (cond - uncond) * scale)
//   sampling switch        models/vaura_model.py:816-825
//   sample_top_k           utils/utils.py:163-177  (keep probs >= k-th largest, ties kept, renormalise)
//   sample_top_p           utils/utils.py:180-196  (keep sorted prefix while cumsum - p_i <= top_p)
//   multinomial            utils/utils.py:139-160  (torch.multinomial; ours: Philox4x32-10 + inverse CDF in
//                                                   vocabulary order, statistically equivalent)
//   mask-fix / write-back  models/vaura_model.py:536-544
#include "common.cuh"
#include "kernels.h"

namespace vaura {

constexpr int kEPT = 32;  // elements per lane: V = 32 * kEPT = 1024

// Largest uint32 t such that count(bits >= t) >= k, i.e. the k-th largest value (non-negative floats
// order like their bit patterns).
__device__ __forceinline__ uint32_t kth_largest_bits(const uint32_t (&bits)[kEPT], int k) {
  uint32_t t = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t cand = t | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < kEPT; ++i) c += (bits[i] >= cand);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) t = cand;
  }
  return t;
}

// top-p: kept set = { i : sum of probs strictly larger than p_i  <= top_p }.  The predicate
// g(t) = [ sum_{x > t} x > top_p ] is monotone (true for small t); find the largest t with g true,
// kept = { x > t }.  If g(0) is false (top_p >= total) everything is kept.
__device__ __forceinline__ uint32_t top_p_threshold_bits(const float (&pr)[kEPT], const uint32_t (&bits)[kEPT],
                                                          float top_p, bool& keep_all) {
  auto mass_above = [&](uint32_t t) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kEPT; ++i) s += (bits[i] > t) ? pr[i] : 0.f;
    return warp_sum(s);
  };
  keep_all = !(mass_above(0u) > top_p);
  uint32_t t = 0;
  if (!keep_all) {
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t cand = t | (1u << bit);
      if (mass_above(cand) > top_p) t = cand;
    }
  }
  return t;
}

__global__ void __launch_bounds__(128) sample_kernel(SampleArgs a) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);  // (b, k)
  const int offset = a.state ? a.state->offset : a.offset;
  if (row < a.B * a.K) {
    const int b = row / a.K, k = row % a.K;
    const int V = a.V;
    // lane owns the contiguous slice [lane*32, lane*32+32) so the inverse CDF runs in vocabulary order
    float x[kEPT];
    {
      const float4* c4 = reinterpret_cast<const float4*>(a.logits + ((size_t)b * a.K + k) * V + lane * kEPT);
#pragma unroll
      for (int i = 0; i < kEPT / 4; ++i) {
        float4 v = c4[i];
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
      if (a.use_cfg) {
        const float4* u4 = reinterpret_cast<const float4*>(a.logits + ((size_t)(a.B + b) * a.K + k) * V + lane * kEPT);
#pragma unroll
        for (int i = 0; i < kEPT / 4; ++i) {
          float4 u = u4[i];
          x[4 * i] = u.x + (x[4 * i] - u.x) * a.cfg_scale;
          x[4 * i + 1] = u.y + (x[4 * i + 1] - u.y) * a.cfg_scale;
          x[4 * i + 2] = u.z + (x[4 * i + 2] - u.z) * a.cfg_scale;
          x[4 * i + 3] = u.w + (x[4 * i + 3] - u.w) * a.cfg_scale;
        }
      }
    }
    if (a.logits_out) {
      float4* o4 = reinterpret_cast<float4*>(a.logits_out + (((size_t)offset * a.B + b) * a.K + k) * V + lane * kEPT);
#pragma unroll
      for (int i = 0; i < kEPT / 4; ++i) o4[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }

    int token;
    if (!(a.use_sampling && a.temp > 0.f)) {
      // argmax, first index on ties (vaura_model.py:825)
      float best = x[0];
      int bi = 0;
#pragma unroll
      for (int i = 1; i < kEPT; ++i)
        if (x[i] > best) { best = x[i]; bi = i; }
      bi += lane * kEPT;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      token = bi;
      if (a.probs_out) {
#pragma unroll
        for (int i = 0; i < kEPT; ++i)
          a.probs_out[((size_t)b * a.K + k) * V + lane * kEPT + i] = (lane * kEPT + i == bi) ? 1.f : 0.f;
      }
    } else {
      // softmax(logits / temp)  (vaura_model.py:817)
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < kEPT; ++i) { x[i] = x[i] / a.temp; m = fmaxf(m, x[i]); }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kEPT; ++i) { x[i] = expf(x[i] - m); s += x[i]; }
      s = warp_sum(s);
      uint32_t bits[kEPT];
#pragma unroll
      for (int i = 0; i < kEPT; ++i) { x[i] = x[i] / s; bits[i] = __float_as_uint(x[i]); }

      if (a.top_p > 0.f) {
        bool keep_all;
        const uint32_t t = top_p_threshold_bits(x, bits, a.top_p, keep_all);
        if (!keep_all) {
#pragma unroll
          for (int i = 0; i < kEPT; ++i) if (!(bits[i] > t)) x[i] = 0.f;
        }
      } else if (a.top_k > 0 && a.top_k < V) {
        const uint32_t t = kth_largest_bits(bits, a.top_k);
#pragma unroll
        for (int i = 0; i < kEPT; ++i) if (bits[i] < t) x[i] = 0.f;
      }

      // inclusive scan of per-lane masses, then locate u * total
      float local = 0.f;
#pragma unroll
      for (int i = 0; i < kEPT; ++i) local += x[i];
      float incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const float total = __shfl_sync(0xffffffffu, incl, 31);
      if (a.probs_out) {
#pragma unroll
        for (int i = 0; i < kEPT; ++i) a.probs_out[((size_t)b * a.K + k) * V + lane * kEPT + i] = x[i] / total;
      }
      const int clip = a.clip_ids ? a.clip_ids[b] : b;
      const uint4 rnd = philox4x32_10(make_uint4((uint32_t)clip, (uint32_t)offset, (uint32_t)k, 0u),
                                      make_uint2(a.seed_lo, a.seed_hi));
      const float u01 = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
      const float target = u01 * total;
      const unsigned hit = __ballot_sync(0xffffffffu, incl > target);
      int cand = -1;
      if (hit) {
        const int src = __ffs(hit) - 1;
        if (lane == src) {
          float run = incl - local;
#pragma unroll
          for (int i = 0; i < kEPT; ++i) {
            run += x[i];
            if (cand < 0 && run > target && x[i] > 0.f) cand = lane * kEPT + i;
          }
        }
        cand = __shfl_sync(0xffffffffu, cand, src);
      }
      if (cand < 0) {  // rounding guard: last kept index
        int last = -1;
#pragma unroll
        for (int i = 0; i < kEPT; ++i) if (x[i] > 0.f) last = lane * kEPT + i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        cand = last;
      }
      token = cand;
    }

    if (lane == 0) {
      if (a.tokens_out) a.tokens_out[row] = token;
      if (a.sequence) {
        const int t = offset - 1 - k;  // timestep held by column `offset` of codebook k
        if (!(t >= 0 && t < a.T)) token = V;  // mask-fix to the special id (vaura_model.py:536-537)
        int32_t* cell = a.sequence + ((size_t)b * a.K + k) * a.S + offset;
        if (*cell == kUnknown) *cell = token;  // keep prompt tokens (vaura_model.py:540-544)
      }
    }
  }
  // last CTA to finish advances the device-resident loop state
  if (a.state) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int prev = atomicAdd(&a.state->done, 1);
      if (prev == (int)gridDim.x - 1) {
        a.state->done = 0;
        a.state->offset = offset + 1;
        __threadfence();
      }
    }
  }
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st) {
  if (a.V != 32 * kEPT) return cudaErrorInvalidValue;
  const int rows = a.B * a.K;
  sample_kernel<<<(rows + 3) / 4, 128, 0, st>>>(a);
  return cudaGetLastError();
}

__global__ void set_state_kernel(StepState* s, int offset) {
  s->offset = offset;
  s->done = 0;
}

cudaError_t launch_set_state(StepState* s, int offset, cudaStream_t st) {
  set_state_kernel<<<1, 1, 0, st>>>(s, offset);
  return cudaGetLastError();
}

}  // namespace vaura
