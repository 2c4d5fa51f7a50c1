// Per-row sampling stage shared by sample_kernel (sampling.cu) and the persistent decode-step kernel
// (decode_persistent.cu).  See sampling.cu for the reference lines each part replaces.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace vaura {

constexpr int kEPT = 32;  // elements per lane: V = 32 * kEPT = 1024

// Threshold t with { i : bits_i >= t } == the k largest values plus ties of the k-th (non-negative floats order like
// their bit patterns): bit-serial search for the largest t with count(bits >= t) >= k, two bits per round (three
// candidate counts as independent chains, one reduction round trip), leaving early once a candidate keeps exactly k
// values - every later refinement would keep the same set.
__device__ __forceinline__ uint32_t kth_largest_bits(const uint32_t (&bits)[kEPT], int k) {
  uint32_t t = 0;
  {  // bit 30 alone (31 value bits)
    const uint32_t cand = 1u << 30;
    int c = 0;
#pragma unroll
    for (int i = 0; i < kEPT; ++i) c += (bits[i] >= cand);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c == k) return cand;
    if (c > k) t = cand;
  }
  for (int bit = 28; bit >= 0; bit -= 2) {
    const uint32_t c1 = t | (1u << bit), c2 = t | (2u << bit), c3 = t | (3u << bit);
    int n1a = 0, n1b = 0, n2a = 0, n2b = 0, n3a = 0, n3b = 0;
#pragma unroll
    for (int i = 0; i < kEPT; i += 2) {
      n1a += (bits[i] >= c1); n1b += (bits[i + 1] >= c1);
      n2a += (bits[i] >= c2); n2b += (bits[i + 1] >= c2);
      n3a += (bits[i] >= c3); n3b += (bits[i + 1] >= c3);
    }
    // counts are <= 1024: two share one warp reduction, the third takes its own (three 11-bit fields do not fit)
    const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)(n1a + n1b) | ((unsigned)(n2a + n2b) << 16));
    const int n3 = __reduce_add_sync(0xffffffffu, n3a + n3b);
    const int n1 = tot & 0xffff, n2 = tot >> 16;
    if (n3 >= k) { t = c3; if (n3 == k) break; }
    else if (n2 >= k) { t = c2; if (n2 == k) break; }
    else if (n1 >= k) { t = c1; if (n1 == k) break; }
  }
  return t;
}

// a / b for normal-range operands without the out-of-line slow path of the compiler's IEEE division (its range check
// is a branch per quotient, which serialises the 32 independent divisions of a lane): reciprocal r = 1/b rounded to
// nearest, then the two residual corrections of the standard sequence.  Equal to a / b when neither the operands nor
// the quotient are subnormal or overflow; within one ulp otherwise.
__device__ __forceinline__ float div_by(float a, float b, float r) {
  float q = a * r;
  q = fmaf(fmaf(-b, q, a), r, q);
  return fmaf(fmaf(-b, q, a), r, q);
}

// top-p: kept set = { i : sum of probs strictly larger than p_i  <= top_p }.  The predicate
// g(t) = [ sum_{x > t} x > top_p ] is monotone (true for small t); find the largest t with g true,
// kept = { x > t }.  If g(0) is false (top_p >= total) everything is kept.
__device__ __forceinline__ uint32_t top_p_threshold_bits(const float (&pr)[kEPT], const uint32_t (&bits)[kEPT],
                                                          float top_p, bool& keep_all) {
  auto mass_above = [&](uint32_t t) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < kEPT; i += 4) {
      s0 += (bits[i] > t) ? pr[i] : 0.f;
      s1 += (bits[i + 1] > t) ? pr[i + 1] : 0.f;
      s2 += (bits[i + 2] > t) ? pr[i + 2] : 0.f;
      s3 += (bits[i + 3] > t) ? pr[i + 3] : 0.f;
    }
    return warp_sum((s0 + s1) + (s2 + s3));
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

// One warp samples the token of (clip b, codebook k) for column `offset` and writes it back.
static __device__ __noinline__ void sample_row(const SampleArgs& a, int b, int k, int lane, int offset) {
  const int V = a.V;
  // lane owns the contiguous slice [lane*32, lane*32+32) so the inverse CDF runs in vocabulary order
  float x[kEPT];
  {
    const float4* c4 = reinterpret_cast<const float4*>(a.logits + ((size_t)b * a.K + k) * V + lane * kEPT);
#pragma unroll
    for (int i = 0; i < kEPT / 4; ++i) {
      float4 v = __ldcg(c4 + i);
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
    if (a.use_cfg) {
      const float4* u4 = reinterpret_cast<const float4*>(a.logits + ((size_t)(a.B + b) * a.K + k) * V + lane * kEPT);
#pragma unroll
      for (int i = 0; i < kEPT / 4; ++i) {
        float4 u = __ldcg(u4 + i);
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
    // the draw's uniform first: its ten dependent Philox rounds overlap the softmax arithmetic
    const int clip = a.clip_ids ? a.clip_ids[b] : b;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)clip, (uint32_t)offset, (uint32_t)k, a.stream_id),
                                    make_uint2(a.seed_lo, a.seed_hi));
    const float u01 = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
    // softmax(logits / temp)  (vaura_model.py:817); sums run as four interleaved chains per lane
    if (a.temp != 1.0f) {
      const float rt = 1.0f / a.temp;
#pragma unroll
      for (int i = 0; i < kEPT; ++i) x[i] = div_by(x[i], a.temp, rt);
    }
    float m0 = fmaxf(x[0], x[1]), m1 = fmaxf(x[2], x[3]);
#pragma unroll
    for (int i = 4; i < kEPT; i += 4) { m0 = fmaxf(m0, fmaxf(x[i], x[i + 1])); m1 = fmaxf(m1, fmaxf(x[i + 2], x[i + 3])); }
    const float m = warp_max(fmaxf(m0, m1));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < kEPT; i += 4) {
      x[i] = expf(x[i] - m); x[i + 1] = expf(x[i + 1] - m); x[i + 2] = expf(x[i + 2] - m); x[i + 3] = expf(x[i + 3] - m);
      s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3];
    }
    const float s = warp_sum((s0 + s1) + (s2 + s3));
    const float rs = 1.0f / s;
    uint32_t bits[kEPT];
#pragma unroll
    for (int i = 0; i < kEPT; ++i) { x[i] = div_by(x[i], s, rs); bits[i] = __float_as_uint(x[i]); }

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
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
    for (int i = 0; i < kEPT; i += 4) { l0 += x[i]; l1 += x[i + 1]; l2 += x[i + 2]; l3 += x[i + 3]; }
    const float local = (l0 + l1) + (l2 + l3);
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    if (a.probs_out) {
      const float rtot = 1.0f / total;
      float4* po = reinterpret_cast<float4*>(a.probs_out + ((size_t)b * a.K + k) * V + lane * kEPT);
#pragma unroll
      for (int i = 0; i < kEPT; i += 4)
        po[i / 4] = make_float4(div_by(x[i], total, rtot), div_by(x[i + 1], total, rtot), div_by(x[i + 2], total, rtot),
                                div_by(x[i + 3], total, rtot));
    }
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
    if (a.tokens_out) a.tokens_out[b * a.K + k] = token;
    if (a.sequence) {
      const int t = offset - 1 - k;  // timestep held by column `offset` of codebook k
      if (!(t >= 0 && t < a.T)) token = V;  // mask-fix to the special id (vaura_model.py:536-537)
      int32_t* cell = a.sequence + ((size_t)b * a.K + k) * a.S + offset;
      if (*cell == kUnknown) *cell = token;  // keep prompt tokens (vaura_model.py:540-544)
    }
  }
}

}  // namespace vaura
