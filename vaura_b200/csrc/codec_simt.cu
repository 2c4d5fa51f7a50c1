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
  if (threadIdx.x < Kc) code[threadIdx.x] = codes[((size_t)b * Kc + threadIdx.x) * T + t];
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

// final conv (C -> 1, k=7, pad 3) + tanh; weights fp32 [7][C]
__global__ void __launch_bounds__(256) conv_out_tanh_kernel(const __half* __restrict__ in, const float* __restrict__ W,
                                                            const float* __restrict__ bias, __half* __restrict__ wav, int T,
                                                            int C) {
  extern __shared__ float ws[];  // [7][C]
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (t >= T) return;
  float s = bias[0];
  for (int j = 0; j < 7; ++j) {
    const int ti = t + j - 3;
    if (ti < 0 || ti >= T) continue;
    const uint4* p = reinterpret_cast<const uint4*>(in + ((size_t)b * T + ti) * C);
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 raw = p[c8];
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      const float* w = ws + j * C + c8 * 8;
#pragma unroll
      for (int e = 0; e < 4; ++e) s = fmaf(__low2float(h[e]), w[2 * e], fmaf(__high2float(h[e]), w[2 * e + 1], s));
    }
  }
  wav[(size_t)b * T + t] = __float2half_rn(tanhf(s));
}

cudaError_t launch_conv_out_tanh(const __half* in, const float* W, const float* bias, __half* wav, int B, int T, int C,
                                 cudaStream_t st) {
  if (C % 8 != 0) return cudaErrorInvalidValue;
  conv_out_tanh_kernel<<<dim3((T + 255) / 256, B), 256, 7 * C * sizeof(float), st>>>(in, W, bias, wav, T, C);
  return cudaGetLastError();
}

}  // namespace vaura
