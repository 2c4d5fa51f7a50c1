// bf16-activation sampler path (VAURA_PRECISION_BF16, batch >= 16 and prefill): the projections run on
// tcgen05 (gemm_tcgen05.cu); this file holds the two memory-bound stages between them.
//   rmsnorm_bf16_kernel  llama.py:147-158, fp32 math, bf16 output (the GEMM's A operand)
//   attn_bf16_kernel     llama.py:246-255 over the paged bf16 KV cache, fp32 softmax / accumulation
#include "common.cuh"
#include "kernels.h"

namespace vaura {

// one CTA per row, one float4 per thread: a single L2 round trip instead of a serial loop per warp
__global__ void __launch_bounds__(512) rmsnorm_bf16_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ out, int R, int D, size_t ldh, float eps,
                                                           int pdl) {
  if (pdl) {  // programmatic dependent launch: wait for the producer of h, then let our consumer start its prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pdl & 8) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  __shared__ float red[16];
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = h + (size_t)r * ldh;
  const int n4 = D >> 2;
  float4 v[2];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * blockDim.x;
    v[i] = c < n4 ? *reinterpret_cast<const float4*>(x + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float rs = rsqrtf(tot / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * blockDim.x;
    if (c < n4) {
      const float4 g = *reinterpret_cast<const float4*>(w + 4 * c);
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[i].x * rs * g.x, v[i].y * rs * g.y);
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[i].z * rs * g.z, v[i].w * rs * g.w);
      *reinterpret_cast<uint2*>(out + (size_t)r * D + 4 * c) = o;
    }
  }
}

// RMSNorm in fp32 with the GEMV path's operation order ((x * rs) * w, llama.py:157-158), output as three bf16 terms
// side by side: out3[r][0:D] | [D:2D] | [2D:3D]
__global__ void __launch_bounds__(512) rmsnorm_split3_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                             __nv_bfloat16* __restrict__ out3, int D, size_t ldh, float eps) {
  __shared__ float red[16];
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = h + (size_t)r * ldh;
  const int n4 = D >> 2;
  float4 v[2];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * blockDim.x;
    v[i] = c < n4 ? *reinterpret_cast<const float4*>(x + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float rs = rsqrtf(tot / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * blockDim.x;
    if (c < n4) {
      const float4 g = *reinterpret_cast<const float4*>(w + 4 * c);
      const float y[4] = {v[i].x * rs * g.x, v[i].y * rs * g.y, v[i].z * rs * g.z, v[i].w * rs * g.w};
      __nv_bfloat16 t[3][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split3(y[e], t[0][e], t[1][e], t[2][e]);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        *reinterpret_cast<uint2*>(out3 + (size_t)r * 3 * D + (size_t)k * D + 4 * c) = *reinterpret_cast<const uint2*>(t[k]);
    }
  }
}

cudaError_t launch_rmsnorm_split3(const float* h, const float* w, void* out3, int R, int D, size_t ldh, float eps, cudaStream_t st) {
  int threads = ((D / 4 + 1) / 2 + 31) / 32 * 32;  // two float4 per thread
  if (threads > 512) threads = 512;
  if (threads < 32) threads = 32;
  if (D % 4 != 0 || D / 4 > 2 * threads) return cudaErrorInvalidValue;
  rmsnorm_split3_kernel<<<R, threads, 0, st>>>(h, w, reinterpret_cast<__nv_bfloat16*>(out3), D, ldh, eps);
  return cudaGetLastError();
}

static cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t st, cudaLaunchAttribute* attr, int pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cfg;
}

cudaError_t launch_rmsnorm_bf16(const float* h, const float* w, void* out_bf16, int R, int D, size_t ldh, float eps, int pdl,
                                cudaStream_t st) {
  cudaLaunchAttribute attr[1];
  int threads = ((D / 4 + 1) / 2 + 31) / 32 * 32;  // two float4 per thread
  if (threads > 512) threads = 512;
  if (threads < 32) threads = 32;
  if (D % 4 != 0 || D / 4 > 2 * threads) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = pdl_config(dim3(R), dim3(threads), 0, st, attr, pdl);
  return cudaLaunchKernelEx(&cfg, rmsnorm_bf16_kernel, h, w, reinterpret_cast<__nv_bfloat16*>(out_bf16), R, D, ldh, eps, pdl);
}

__global__ void __launch_bounds__(128) attn_bf16_kernel(AttnBf16Args a) {
  __shared__ float qs[kHeadDim];
  __shared__ float sc[kMaxCtx];
  __shared__ float red[4];
  if (a.pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.pdl & 8) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  const int h = blockIdx.x, row = blockIdx.y;
  const int b = row / a.npos, j = row % a.npos;
  const int pos0 = a.state ? a.state->offset - a.npos : a.pos0;
  const int p = pos0 + j, nctx = p + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* kvp = reinterpret_cast<const __nv_bfloat16*>(a.kv.pages);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  if (tid < kHeadDim) qs[tid] = __bfloat162float(q[(size_t)row * a.d_model + h * kHeadDim + tid]);
  __syncthreads();

  // scores: 4 lanes per key position, 24 dims (48 bytes = 3 x 16B) each
  const int g = tid >> 2, t = tid & 3;
  float q24[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) q24[i] = qs[t * 24 + i];
  const int nround = (nctx + 31) & ~31;
  for (int jj = g; jj < nround; jj += 32) {
    float s = 0.f;
    if (jj < nctx) {
      const uint4* kr = reinterpret_cast<const uint4*>(kvp + a.kv.row(a.layer, 0, b, jj, h) + t * 24);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 kk = kr[c];
        const uint32_t w[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s = fmaf(q24[c * 8 + 2 * e], bf16_lo(w[e]), s);
          s = fmaf(q24[c * 8 + 2 * e + 1], bf16_hi(w[e]), s);
        }
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (t == 0 && jj < nctx) sc[jj] = s * a.scale;
  }
  __syncthreads();

  float m = -INFINITY;
  for (int jj = tid; jj < nctx; jj += 128) m = fmaxf(m, sc[jj]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int jj = tid; jj < nctx; jj += 128) {
    const float e = expf(sc[jj] - m);
    sc[jj] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];

  // P.V: 48 threads x 2 dims; the 128 threads split the positions in 2 halves (threads 48..95 take odd ones)
  __shared__ float part[2][kHeadDim];
  if (tid < 96) {
    const int half = tid / 48, d2 = (tid % 48) * 2;
    float a0 = 0.f, a1 = 0.f;
    for (int jj = half; jj < nctx; jj += 2) {
      const uint32_t vv = *reinterpret_cast<const uint32_t*>(kvp + a.kv.row(a.layer, 1, b, jj, h) + d2);
      a0 = fmaf(sc[jj], bf16_lo(vv), a0);
      a1 = fmaf(sc[jj], bf16_hi(vv), a1);
    }
    part[half][d2] = a0;
    part[half][d2 + 1] = a1;
  }
  __syncthreads();
  if (tid < kHeadDim) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out);
    o[(size_t)row * a.d_model + h * kHeadDim + tid] = __float2bfloat16_rn((part[0][tid] + part[1][tid]) / sum);
  }
}

cudaError_t launch_attn_bf16(const AttnBf16Args& a, int nhead, int rows, cudaStream_t st) {
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = pdl_config(dim3(nhead, rows), dim3(128), 0, st, attr, a.pdl);
  return cudaLaunchKernelEx(&cfg, attn_bf16_kernel, a);
}

}  // namespace vaura
