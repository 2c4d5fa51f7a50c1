// DAC decoder stages on the CUDA cores (fp16 storage, fp32 accumulate) in channels-last layout
// [B][T][C].  Every convolution is an implicit GEMM over (time, Cout) tiles with a loop over taps;
// ConvTranspose1d is run as `stride` polyphase GEMMs of two taps each.  Bias, residual add, Snake
// and tanh are fused into the epilogues, so each activation tensor is written once per consumer.
//
// Replaces (arithmetic of pip descript-audio-codec 1.0.0, called at
// /root/reference/models/modules/dac/model.py:41-48):
//   from_codes_kernel     ResidualVectorQuantize.from_codes  (sum_k out_proj_k(codebook_k[codes_k]))
//   conv_gemm_kernel      WNConv1d k7 (dil 1/3/9) / k1, WNConvTranspose1d, Snake1d, residual add
//   conv_out_tanh_kernel  final WNConv1d(C->1, k7) + tanh
#include "common.cuh"
#include "kernels.h"

namespace vaura {

__global__ void __launch_bounds__(256) from_codes_kernel(const int32_t* __restrict__ codes, const __half* __restrict__ tables,
                                                         __half* __restrict__ z, int Kc, int T, int Vc, int latent) {
  const int t = blockIdx.x, b = blockIdx.y;
  __shared__ int code[16];
  if (threadIdx.x < Kc) {  // clamped: an out-of-range code never reads outside the table (the host mirror raises on it)
    const int c = codes[((size_t)b * Kc + threadIdx.x) * T + t];
    code[threadIdx.x] = c < 0 ? 0 : (c >= Vc ? Vc - 1 : c);
  }
  __syncthreads();
  for (int c = threadIdx.x * 2; c < latent; c += blockDim.x * 2) {
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < Kc; ++k) {
      const __half2 v = *reinterpret_cast<const __half2*>(tables + ((size_t)k * Vc + code[k]) * latent + c);
      s0 += __low2float(v);
      s1 += __high2float(v);
    }
    *reinterpret_cast<__half2*>(z + ((size_t)b * T + t) * latent + c) = __floats2half2_rn(s0, s1);
  }
}

cudaError_t launch_from_codes(const int32_t* codes, const __half* tables, __half* z, int B, int Kc, int T, int Vc,
                              int latent, cudaStream_t st) {
  from_codes_kernel<<<dim3(T, B), 256, 0, st>>>(codes, tables, z, Kc, T, Vc, latent);
  return cudaGetLastError();
}

__device__ __forceinline__ float snake_f(float v, float alpha, float inv_alpha) {
  const float s = __sinf(alpha * v);
  return fmaf(s * s, inv_alpha, v);
}

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) conv_gemm_kernel(ConvArgs a) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int q0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int b = blockIdx.z / a.nphase, phase = blockIdx.z % a.nphase;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const __half* in = a.in + (size_t)b * a.Tin * a.Cin;
  const __half* W = a.W + (size_t)phase * a.ntaps * a.Cout * a.Cin;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader: 64 rows x 4 chunks of 4 halfs
  for (int tap = 0; tap < a.ntaps; ++tap) {
    const int off = a.tap_off[phase * a.ntaps + tap];
    const __half* Wt = W + (size_t)tap * a.Cout * a.Cin;
    for (int k0 = 0; k0 < a.Cin; k0 += TK) {
      {
        const int t = q0 + lrow + off;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < a.Tin && q0 + lrow < a.Tq) {
          const uint2 raw = *reinterpret_cast<const uint2*>(in + (size_t)t * a.Cin + k0 + lk);
          const __half2 h0 = *reinterpret_cast<const __half2*>(&raw.x), h1 = *reinterpret_cast<const __half2*>(&raw.y);
          v = make_float4(__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1));
        }
        As[lk][lrow] = v.x; As[lk + 1][lrow] = v.y; As[lk + 2][lrow] = v.z; As[lk + 3][lrow] = v.w;
        const int co = n0 + lrow;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co < a.Cout) {
          const uint2 raw = *reinterpret_cast<const uint2*>(Wt + (size_t)co * a.Cin + k0 + lk);
          const __half2 h0 = *reinterpret_cast<const __half2*>(&raw.x), h1 = *reinterpret_cast<const __half2*>(&raw.y);
          w = make_float4(__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1));
        }
        Bs[lk][lrow] = w.x; Bs[lk + 1][lrow] = w.y; Bs[lk + 2][lrow] = w.z; Bs[lk + 3][lrow] = w.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float am[4] = {av.x, av.y, av.z, av.w}, bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= a.Tq) continue;
    const size_t t_out = (size_t)q * a.ostride + phase;
    const size_t base = ((size_t)b * a.Tout + t_out) * a.Cout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      float v = acc[i][j] + a.bias[co];
      if (a.residual) v += __half2float(a.residual[base + co]);
      if (a.out_raw) a.out_raw[base + co] = __float2half_rn(v);
      if (a.out_act) a.out_act[base + co] = __float2half_rn(snake_f(v, a.alpha[co], a.alpha[a.Cout + co]));
    }
  }
}

cudaError_t launch_conv_gemm(const ConvArgs& a, int B, cudaStream_t st) {
  if (a.Cin % TK != 0) return cudaErrorInvalidValue;
  dim3 grid((a.Tq + TM - 1) / TM, (a.Cout + TN - 1) / TN, B * a.nphase);
  conv_gemm_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// final conv (C -> 1, k=7, pad 3) + tanh; weights fp32 [7][C].
// One warp walks kOutPerWarp consecutive outputs of a clip: lane l holds channels [4l, 4l+4) (C <= 128: one coalesced
// 8-byte load per lane per input row), every input row feeds the seven outputs it overlaps through seven rotating
// partial sums, and the output that has seen its last row is reduced across the warp and written.  The input is read
// once, row-contiguously (the per-output version read every row seven times through strided 16-byte pieces).
constexpr int kOutPerWarp = 128;
__global__ void __launch_bounds__(256) conv_out_tanh_kernel(const __half* __restrict__ in, const float* __restrict__ W,
                                                            const float* __restrict__ bias, __half* __restrict__ wav, int T,
                                                            int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int t0 = (blockIdx.x * (blockDim.x >> 5) + warp) * kOutPerWarp;
  if (t0 >= T) return;
  const int t1 = min(t0 + kOutPerWarp, T);
  const bool live = 4 * lane < C;
  float w[7][4];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[j][e] = live ? W[j * C + 4 * lane + e] : 0.f;
  const float b0 = bias[0];
  const __half* base = in + (size_t)b * T * C + 4 * lane;
  // acc[j]: partial sum of output (r - j + 3) after row r has been added, i.e. acc[0] is the youngest output
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int U = 8;  // rows in flight
  for (int r0 = t0 - 3; r0 < t1 + 3; r0 += U) {
    uint2 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u;
      raw[u] = (live && r >= 0 && r < T && r < t1 + 3) ? __ldg(reinterpret_cast<const uint2*>(base + (size_t)r * C)) : make_uint2(0u, 0u);
    }
    float done[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const __half2 h0 = *reinterpret_cast<const __half2*>(&raw[u].x), h1 = *reinterpret_cast<const __half2*>(&raw[u].y);
      const float x0 = __low2float(h0), x1 = __high2float(h0), x2 = __low2float(h1), x3 = __high2float(h1);
      // row r is tap j of output r - j + 3: the oldest live output (tap 6) completes with this row
#pragma unroll
      for (int j = 0; j < 7; ++j) acc[j] = fmaf(x0, w[j][0], fmaf(x1, w[j][1], fmaf(x2, w[j][2], fmaf(x3, w[j][3], acc[j]))));
      done[u] = acc[6];
#pragma unroll
      for (int j = 6; j > 0; --j) acc[j] = acc[j - 1];
      acc[0] = 0.f;
    }
    // the U outputs completed by these rows: U interleaved butterfly reductions, lane u writes output r0 + u - 3
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) done[u] += __shfl_xor_sync(0xffffffffu, done[u], o);
    float mine = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) mine = lane == u ? done[u] : mine;
    const int t = r0 + lane - 3;
    if (lane < U && t >= t0 && t < t1) wav[(size_t)b * T + t] = __float2half_rn(tanhf(mine + b0));
  }
}

cudaError_t launch_conv_out_tanh(const __half* in, const float* W, const float* bias, __half* wav, int B, int T, int C,
                                 cudaStream_t st) {
  if (C % 8 != 0 || C > 128) return cudaErrorInvalidValue;
  const int warps = (T + kOutPerWarp - 1) / kOutPerWarp;
  conv_out_tanh_kernel<<<dim3((warps + 7) / 8, B), 256, 0, st>>>(in, W, bias, wav, T, C);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------------
// encode direction (dac 1.0.0 Encoder / ResidualVectorQuantize.forward, reached through
// DacModelWrapper.encode, models/modules/dac/model.py:30-39).  The encoder's ResidualUnits, strided convolutions (run as
// three-tap convolutions over frames of `stride` samples) and the final k3 convolution go through the same implicit-GEMM
// kernels as the decoder; only the two ends are new.
// ------------------------------------------------------------------------------------------------------------------------
// first layer: WNConv1d(1 -> C, k 7, pad 3) on the waveform, raw output + Snake with the first residual unit's alpha
__global__ void __launch_bounds__(256) enc_conv_in_kernel(const float* __restrict__ wav, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ alpha,
                                                          __half* __restrict__ out_raw, __half* __restrict__ out_act, int L, int C) {
  // thread = (time step, 8 channels)
  const int c8 = C / 8;
  const size_t total = (size_t)L * c8;
  const int b = blockIdx.y;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8) * 8, t = (int)(i / c8);
    float x[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int r = t + j - 3;
      x[j] = (r >= 0 && r < L) ? __ldg(wav + (size_t)b * L + r) : 0.f;
    }
    uint4 raw, act;
    __half2* hr = reinterpret_cast<__half2*>(&raw);
    __half2* ha = reinterpret_cast<__half2*>(&act);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = c0 + e + u;
        float acc = __ldg(bias + c);
#pragma unroll
        for (int j = 0; j < 7; ++j) acc = fmaf(__ldg(W + c * 7 + j), x[j], acc);
        v[u] = acc;
      }
      hr[e / 2] = __floats2half2_rn(v[0], v[1]);
      ha[e / 2] = __floats2half2_rn(snake_f(v[0], __ldg(alpha + c0 + e), __ldg(alpha + C + c0 + e)),
                                    snake_f(v[1], __ldg(alpha + c0 + e + 1), __ldg(alpha + C + c0 + e + 1)));
    }
    const size_t o = ((size_t)b * L + t) * C + c0;
    *reinterpret_cast<uint4*>(out_raw + o) = raw;
    *reinterpret_cast<uint4*>(out_act + o) = act;
  }
}

cudaError_t launch_enc_conv_in(const float* wav, const float* W, const float* bias, const float* alpha, __half* out_raw,
                               __half* out_act, int B, int L, int C, cudaStream_t st) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = (size_t)L * (C / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  enc_conv_in_kernel<<<dim3(blocks, B), 256, 0, st>>>(wav, W, bias, alpha, out_raw, out_act, L, C);
  return cudaGetLastError();
}

// residual vector quantisation of one latent frame per CTA (dac/nn/quantize.py VectorQuantize.decode_latents +
// ResidualVectorQuantize.forward): for every codebook k: e = in_proj_k(residual), cosine nearest neighbour among the
// l2-normalised code vectors (first index on ties, like torch.max), residual -= out_proj_k(codebook_k[idx]) (fp32 table).
constexpr int kRvqThreads = 256;
__global__ void __launch_bounds__(kRvqThreads)
rvq_encode_kernel(const __half* __restrict__ z, const float* __restrict__ w_in, const float* __restrict__ b_in,
                  const float* __restrict__ cb_norm, const float* __restrict__ tables, int32_t* __restrict__ codes, int Kc, int T,
                  int Vc, int latent, int Dc) {
  extern __shared__ float rsm[];
  float* res = rsm;                 // [latent]
  float* e = rsm + latent;          // [Dc] (<= 32)
  float* bestv = e + 32;            // [warps]
  int* besti = reinterpret_cast<int*>(bestv + kRvqThreads / 32);
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < latent; c += kRvqThreads) res[c] = __half2float(z[((size_t)b * T + t) * latent + c]);
  __syncthreads();
  for (int k = 0; k < Kc; ++k) {
    // in_proj: Dc dot products of length `latent`, one warp each (round robin)
    for (int d = warp; d < Dc; d += kRvqThreads / 32) {
      const float* w = w_in + ((size_t)k * Dc + d) * latent;
      float acc = 0.f;
      for (int c = lane; c < latent; c += 32) acc = fmaf(__ldg(w + c), res[c], acc);
      acc = warp_sum(acc);
      if (lane == 0) e[d] = acc + __ldg(b_in + k * Dc + d);
    }
    __syncthreads();
    float ev[32], nrm = 0.f;
    for (int d = 0; d < Dc; ++d) { ev[d] = e[d]; nrm += ev[d] * ev[d]; }
    const float inv = 1.f / fmaxf(sqrtf(nrm), 1e-12f);  // F.normalize: x / max(||x||, eps)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = tid; j < Vc; j += kRvqThreads) {
      const float* cv = cb_norm + ((size_t)k * Vc + j) * Dc;
      float dot = 0.f, cn = 0.f;
      for (int d = 0; d < Dc; ++d) { const float c = __ldg(cv + d); dot = fmaf(ev[d] * inv, c, dot); cn = fmaf(c, c, cn); }
      // -dist = -(|e|^2 - 2 e.c + |c|^2) with |e| = 1 after normalisation; the common -1 is dropped
      const float sc = 2.f * dot - cn;
      if (sc > bv) { bv = sc; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { bestv[warp] = bv; besti[warp] = bi; }
    __syncthreads();
    bv = bestv[0]; bi = besti[0];
    for (int w2 = 1; w2 < kRvqThreads / 32; ++w2)
      if (bestv[w2] > bv || (bestv[w2] == bv && besti[w2] < bi)) { bv = bestv[w2]; bi = besti[w2]; }
    if (tid == 0) codes[((size_t)b * Kc + k) * T + t] = bi;
    const float* tb = tables + ((size_t)k * Vc + bi) * latent;
    for (int c = tid; c < latent; c += kRvqThreads) res[c] -= __ldg(tb + c);
    __syncthreads();
  }
}

cudaError_t launch_rvq_encode(const __half* z, const float* w_in, const float* b_in, const float* cb_norm, const float* tables,
                              int32_t* codes, int B, int Kc, int T, int Vc, int latent, int Dc, cudaStream_t st) {
  if (Dc > 32 || Dc < 1) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)latent + 32 + 2 * (kRvqThreads / 32)) * 4;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  rvq_encode_kernel<<<dim3(T, B), kRvqThreads, smem, st>>>(z, w_in, b_in, cb_norm, tables, codes, Kc, T, Vc, latent, Dc);
  return cudaGetLastError();
}

}  // namespace vaura
