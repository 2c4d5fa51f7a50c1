// Segment-AVCLIP visual tower: MotionFormer `divided_224_16x4` + the per-frame spatial aggregation layer, i.e.
// what the reference's MotionFormer.forward computes in the shipped configuration
// (models/modules/feature_extractors/avclip/motionformer.py:252-342; motionformer_src/video_model_builder.py:174-274;
// motionformer_src/vit_helper.py:80-171, :392-472).  The dense contractions (tubelet embedding, q|k|v, projections, MLP) run
// on tcgen05 through launch_vit_linear (gemm_tcgen05.cu: persistent tile loop, bias / GELU / residual / position-embedding
// epilogues); this file holds everything between them: tubelet gathering, LayerNorm, the three attention patterns of a
// divided space-time block and the CLS-only attention of the aggregation layer.
//
// Activations of a chunk of S segments (T = 1 + t * n tokens each, D = embed_dim):
//   x    fp32 [S*T][D]      residual stream           xn  bf16 [S*t*(n+1)][D]  LayerNorm output (GEMM operand)
//   qkv  bf16 [S*t*(n+1)][3D]                          att bf16 [S*T][D]        attention output ("b n (h d)")
//   hid  bf16 [S*T][4D]     GELU(fc1) / tubelet matrix [S*t*n][C*2*16*16]
// Token order inside a segment: CLS, then (frame, row, column) - PatchEmbed3D flattens (t, h, w) (vit_helper.py:547-552).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace vaura {

constexpr int kVitDh = 64;  // head width of ViT-B/16 (768 / 12)

// ------------------------------------------------------------------------------------------------------------------------
// tubelets -> rows of the embedding GEMM: A[row][c * tub*ps*ps + dt * ps*ps + dy * ps + dx]  (the flattening of the Conv3d
// weight (D, C, tub, ps, ps), video_model_builder.py:185, vit_helper.py:536-541); one thread = 8 consecutive dx
// ------------------------------------------------------------------------------------------------------------------------
__global__ void vit_patchify_kernel(const float* __restrict__ frames, __nv_bfloat16* __restrict__ A, int S, int C, int T,
                                    int H, int W, int tub, int ps) {
  const int gh = H / ps, gw = W / ps, t = T / tub;
  const int K = C * tub * ps * ps, K8 = K / 8;
  const size_t total = (size_t)S * t * gh * gw * K8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K8) * 8;
    const size_t row = i / K8;
    const int px = (int)(row % gw), py = (int)((row / gw) % gh), tt = (int)((row / ((size_t)gw * gh)) % t);
    const int s = (int)(row / ((size_t)gw * gh * t));
    const int dx = k % ps, dy = (k / ps) % ps, dt = (k / (ps * ps)) % tub, c = k / (ps * ps * tub);
    const float* src = frames + ((((size_t)s * C + c) * T + (tt * tub + dt)) * H + (py * ps + dy)) * W + px * ps + dx;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
    uint4 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(a.x, a.y);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(a.z, a.w);
    *reinterpret_cast<__nv_bfloat162*>(&o.z) = __floats2bfloat162_rn(b.x, b.y);
    *reinterpret_cast<__nv_bfloat162*>(&o.w) = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(A + row * K + k) = o;
  }
}

// dst[i * row_stride][0..D) = src[0..D)   (CLS rows of the token matrix, start value of the aggregation residual)
__global__ void vit_broadcast_row_kernel(float* __restrict__ dst, const float* __restrict__ src, int D, int count, size_t row_stride) {
  const int i = blockIdx.x;
  if (i >= count) return;
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[(size_t)i * row_stride * D + c] = src[c];
}

// ------------------------------------------------------------------------------------------------------------------------
// LayerNorm (eps inside the square root, biased variance: torch.nn.LayerNorm), one warp per row, fp32 in, bf16 out
// ------------------------------------------------------------------------------------------------------------------------
template <int NV>  // NV float4 per lane: D = 128 * NV
__device__ __forceinline__ void ln_row(float4 (&v)[NV], const float* __restrict__ g, const float* __restrict__ b, float eps,
                                       int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)(128 * NV);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)(128 * NV) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    v[i].x = v[i].x * rstd * gg.x + bb.x; v[i].y = v[i].y * rstd * gg.y + bb.y;
    v[i].z = v[i].z * rstd * gg.z + bb.z; v[i].w = v[i].w * rstd * gg.w + bb.w;
  }
}
template <int NV>
__device__ __forceinline__ void ln_store_bf16(const float4 (&v)[NV], __nv_bfloat16* __restrict__ dst, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint2 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[i].x, v[i].y);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[i].z, v[i].w);
    *reinterpret_cast<uint2*>(dst + 4 * (lane + 32 * i)) = o;
  }
}

template <int NV>
__global__ void vit_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                     __nv_bfloat16* __restrict__ out, int rows, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int D = 128 * NV;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *(reinterpret_cast<const float4*>(x + (size_t)row * D) + lane + 32 * i);
  ln_row<NV>(v, g, b, eps, lane);
  ln_store_bf16<NV>(v, out + (size_t)row * D, lane);
}

// Rows of the aggregation layer's sequences (motionformer.py:316-342, :395-399, :449-470): sequence (s, f) = [agg CLS token,
// the n patch tokens of frame f after the tower's final LayerNorm]; what is stored is LayerNorm_agg1 of that row (the layer
// is pre-norm and only its keys / values and the CLS query are needed).
template <int NV>
__global__ void vit_final_norm_agg_kernel(const float* __restrict__ x, const float* __restrict__ agg_cls,
                                          const float* __restrict__ gf, const float* __restrict__ bf,
                                          const float* __restrict__ g1, const float* __restrict__ b1,
                                          __nv_bfloat16* __restrict__ out, int S, int t, int n, float eps) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= S * t * (n + 1)) return;
  constexpr int D = 128 * NV;
  const int j = r % (n + 1), seq = r / (n + 1), f = seq % t, s = seq / t;
  float4 v[NV];
  if (j == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(agg_cls) + lane + 32 * i);
  } else {
    const float* src = x + ((size_t)s * (1 + t * n) + 1 + f * n + (j - 1)) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *(reinterpret_cast<const float4*>(src) + lane + 32 * i);
    ln_row<NV>(v, gf, bf, eps, lane);
  }
  ln_row<NV>(v, g1, b1, eps, lane);
  ln_store_bf16<NV>(v, out + (size_t)r * D, lane);
}

// ------------------------------------------------------------------------------------------------------------------------
// attention helpers: 8 bf16 (one uint4) -> fp32
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ------------------------------------------------------------------------------------------------------------------------
// time attention of a divided block (vit_helper.py:100-171 with "b (f n) d -> (b n) f d"): the t tokens of one spatial
// location attend each other and the CLS token.  One warp per (segment, location, head); lane = (query frame, 16-dim slice).
// ------------------------------------------------------------------------------------------------------------------------
template <int TF>  // frames per segment after tubelet embedding (8)
__global__ void vit_time_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int n,
                                     int heads, float scale) {
  constexpr int SL = 32 / TF, DW = kVitDh / SL;  // slices per query, dims per slice (4 x 16 for TF = 8)
  static_assert(DW % 8 == 0, "a slice is whole 16-byte vectors");
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= S * n * heads) return;
  const int h = w % heads, loc = (w / heads) % n, s = w / (heads * n);
  const int D = heads * kVitDh, T = 1 + TF * n;
  const int qi = lane / SL, sl = lane % SL;
  const size_t ld = 3 * (size_t)D;
  const __nv_bfloat16* base = qkv + (size_t)s * T * ld + h * kVitDh + sl * DW;
  float q[DW];
  {
    const __nv_bfloat16* qp = base + (size_t)(1 + qi * n + loc) * ld;
#pragma unroll
    for (int c = 0; c < DW / 8; ++c) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(qp + 8 * c), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) q[8 * c + e] = f[e] * scale;
    }
  }
  float sc[TF + 1];
#pragma unroll
  for (int j = 0; j <= TF; ++j) {
    const __nv_bfloat16* kp = base + (size_t)(j == 0 ? 0 : 1 + (j - 1) * n + loc) * ld + D;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < DW / 8; ++c) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(kp + 8 * c), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(q[8 * c + e], f[e], d);
    }
#pragma unroll
    for (int o = SL / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    sc[j] = d;
  }
  float mx = sc[0];
#pragma unroll
  for (int j = 1; j <= TF; ++j) mx = fmaxf(mx, sc[j]);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j <= TF; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float acc[DW];
#pragma unroll
  for (int e = 0; e < DW; ++e) acc[e] = 0.f;
#pragma unroll
  for (int j = 0; j <= TF; ++j) {
    const __nv_bfloat16* vp = base + (size_t)(j == 0 ? 0 : 1 + (j - 1) * n + loc) * ld + 2 * D;
    const float p = sc[j] * inv;
#pragma unroll
    for (int c = 0; c < DW / 8; ++c) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(vp + 8 * c), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[8 * c + e] = fmaf(p, f[e], acc[8 * c + e]);
    }
  }
  __nv_bfloat16* op = out + ((size_t)s * T + 1 + qi * n + loc) * D + h * kVitDh + sl * DW;
#pragma unroll
  for (int c = 0; c < DW / 8; ++c) {
    uint4 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(acc[8 * c], acc[8 * c + 1]);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(acc[8 * c + 2], acc[8 * c + 3]);
    *reinterpret_cast<__nv_bfloat162*>(&o.z) = __floats2bfloat162_rn(acc[8 * c + 4], acc[8 * c + 5]);
    *reinterpret_cast<__nv_bfloat162*>(&o.w) = __floats2bfloat162_rn(acc[8 * c + 6], acc[8 * c + 7]);
    *reinterpret_cast<uint4*>(op + 8 * c) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// One query (row 0 of a sequence) against all `len` keys of that sequence: the CLS token of a divided block attends every
// token of its segment in both the time and the space attention (vit_helper.py:130-131); the aggregation layer only needs
// its CLS output row (motionformer.py:443-444).  One CTA per (head, sequence); out row = seq * out_rows_per_seq.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kClsThreads = 256;
__global__ void __launch_bounds__(kClsThreads)
vit_cls_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int len, int heads, float scale,
                    int out_rows_per_seq) {
  extern __shared__ float sm[];
  float* qs = sm;              // [64]
  float* red = sm + 64;        // [32]
  float* part = sm + 96;       // [32][64] partial outputs of the key lanes
  float* sc = sm + 96 + 2048;  // [len]
  const int h = blockIdx.x, seq = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = heads * kVitDh;
  const size_t ld = 3 * (size_t)D;
  const __nv_bfloat16* base = qkv + (size_t)seq * len * ld + h * kVitDh;
  if (tid < kVitDh) qs[tid] = __bfloat162float(base[tid]) * scale;
  __syncthreads();
  float q[kVitDh];
#pragma unroll
  for (int e = 0; e < kVitDh; e += 4) {
    const float4 t4 = *reinterpret_cast<const float4*>(qs + e);
    q[e] = t4.x; q[e + 1] = t4.y; q[e + 2] = t4.z; q[e + 3] = t4.w;
  }
  float mx = -INFINITY;
  for (int j = tid; j < len; j += kClsThreads) {
    const __nv_bfloat16* kp = base + (size_t)j * ld + D;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int c = 0; c < kVitDh / 8; ++c) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(kp + 8 * c), f);
#pragma unroll
      for (int e = 0; e < 8; e += 2) { d0 = fmaf(q[8 * c + e], f[e], d0); d1 = fmaf(q[8 * c + e + 1], f[e + 1], d1); }
    }
    const float d = d0 + d1;
    sc[j] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kClsThreads / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < len; j += kClsThreads) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < kClsThreads / 32; ++i) sum += red[i];
  // P.V: thread = (dim group g of 8, key lane kl of 32); eight neighbouring threads read one 128-byte value row
  const int g = tid & 7, kl = tid >> 3;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = kl; j < len; j += 32) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(base + (size_t)j * ld + 2 * D + 8 * g), f);
    const float p = sc[j];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, f[e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[kl * 64 + 8 * g + e] = acc[e];
  __syncthreads();
  if (tid < kVitDh) {
    float o = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) o += part[k * 64 + tid];
    out[(size_t)seq * out_rows_per_seq * D + h * kVitDh + tid] = __float2bfloat16_rn(o / sum);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// space attention of a divided block (vit_helper.py:100-171 with "b (f n) d -> (b f) n d"): the n patches of one frame attend
// each other and the CLS token.  One CTA per (head, frame, segment): K / V of the n + 1 keys staged in shared memory, each
// warp walks query pairs - scores with one lane per key, P.V with one lane per pair of output dims.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kSpThreads = 256, kSpKStride = 72;  // K rows padded to 144 bytes: conflict-free 16-byte reads by a quarter warp
__host__ __device__ inline size_t space_attn_smem(int n) {
  const int nk = n + 1, nkp = (nk + 31) & ~31;
  return (size_t)nk * kSpKStride * 2 + (size_t)nk * kVitDh * 2 + (size_t)(kSpThreads / 32) * 2 * nkp * 4 + 64;
}
__global__ void __launch_bounds__(kSpThreads)
vit_space_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int t, int n, int heads,
                      float scale) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int nk = n + 1, nkp = (nk + 31) & ~31;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smraw);            // [nk][72]
  __nv_bfloat16* Vs = Ks + (size_t)nk * kSpKStride;                       // [nk][64]
  float* Ps = reinterpret_cast<float*>(Vs + (size_t)nk * kVitDh);         // [warps][2][nkp]
  const int h = blockIdx.x, f = blockIdx.y, s = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = heads * kVitDh, T = 1 + t * n;
  const size_t ld = 3 * (size_t)D;
  const __nv_bfloat16* seg = qkv + (size_t)s * T * ld + h * kVitDh;
  auto key_row = [&](int j) { return j == 0 ? 0 : 1 + f * n + (j - 1); };
  for (int i = tid; i < nk * 8; i += kSpThreads) {
    const int j = i >> 3, c = i & 7;
    const __nv_bfloat16* src = seg + (size_t)key_row(j) * ld;
    *reinterpret_cast<uint4*>(Ks + (size_t)j * kSpKStride + 8 * c) = *reinterpret_cast<const uint4*>(src + D + 8 * c);
    *reinterpret_cast<uint4*>(Vs + (size_t)j * kVitDh + 8 * c) = *reinterpret_cast<const uint4*>(src + 2 * D + 8 * c);
  }
  __syncthreads();
  float* P0 = Ps + (size_t)warp * 2 * nkp;
  float* P1 = P0 + nkp;
  const int niter = (nk + 31) >> 5;
  for (int q0 = 2 * warp; q0 < n; q0 += 2 * (kSpThreads / 32)) {
    const bool two = q0 + 1 < n;
    float inv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float* P = u ? P1 : P0;
      if (u == 1 && !two) { inv[1] = 0.f; break; }
      const __nv_bfloat16* qp = seg + (size_t)(1 + f * n + q0 + u) * ld;
      float q[kVitDh];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float fq[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(qp + 8 * c)), fq);
#pragma unroll
        for (int e = 0; e < 8; ++e) q[8 * c + e] = fq[e] * scale;
      }
      float mx = -INFINITY;
      for (int it = 0; it < niter; ++it) {
        const int j = lane + 32 * it;
        float d = -INFINITY;
        if (j < nk) {
          float d0 = 0.f, d1 = 0.f;
          const __nv_bfloat16* kr = Ks + (size_t)j * kSpKStride;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float fk[8];
            unpack8(*reinterpret_cast<const uint4*>(kr + 8 * c), fk);
#pragma unroll
            for (int e = 0; e < 8; e += 2) { d0 = fmaf(q[8 * c + e], fk[e], d0); d1 = fmaf(q[8 * c + e + 1], fk[e + 1], d1); }
          }
          d = d0 + d1;
        }
        P[j] = d;
        mx = fmaxf(mx, d);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int it = 0; it < niter; ++it) {
        const int j = lane + 32 * it;
        const float e = j < nk ? __expf(P[j] - mx) : 0.f;
        P[j] = e;
        sum += e;
      }
      inv[u] = 1.f / warp_sum(sum);
    }
    __syncwarp();
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    const uint32_t* Vw = reinterpret_cast<const uint32_t*>(Vs) + lane;
#pragma unroll 4
    for (int j = 0; j < nk; ++j) {
      const uint32_t vv = Vw[j * (kVitDh / 2)];
      const float v0 = bf16_lo(vv), v1 = bf16_hi(vv);
      const float p0 = P0[j], p1 = two ? P1[j] : 0.f;
      a00 = fmaf(p0, v0, a00); a01 = fmaf(p0, v1, a01);
      a10 = fmaf(p1, v0, a10); a11 = fmaf(p1, v1, a11);
    }
    __nv_bfloat16* op = out + ((size_t)s * T + 1 + f * n + q0) * D + h * kVitDh + 2 * lane;
    *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(a00 * inv[0], a01 * inv[0]);
    if (two) *reinterpret_cast<__nv_bfloat162*>(op + D) = __floats2bfloat162_rn(a10 * inv[1], a11 * inv[1]);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// space attention on the tensor cores (mma.sync m16n8k16, bf16 in / fp32 accumulate): same CTA = (head, frame, segment) and the
// same result as vit_space_attn_kernel up to the bf16 rounding of the probabilities.  K [key][dim] and V^T [dim][key] of the
// n + 1 keys sit in shared memory (row strides chosen so that the B-fragment reads of a warp hit 32 different banks); a warp
// owns 16 queries at a time: S = Q K^T for all keys in registers (NKT x 2 accumulator tiles), softmax on the fragments, the
// probabilities re-used as the A operand of P V (the accumulator layout of two n8 tiles is the A layout of one k16 step).
// The FLOPs of this phase are 3 % of the tower's; it only has to stay off the critical path (SIMT: 53 % of the tower's time,
// profiles/r02_avclip_launches.summary.txt).
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kSpTcThreads = 128;
template <int NKT>  // key tiles of 16: (n + 1) <= 16 * NKT
struct SpTc {
  static constexpr int NKP = 16 * NKT;           // padded keys
  static constexpr int KST = 72;                 // K row stride (bf16): 36 words -> bank 4 g + t
  static constexpr int VST = NKP + 8;            // V^T row stride (bf16): (NKP + 8) / 2 words; 108 for NKT = 13 -> bank 12 g + t
  static constexpr size_t smem = (size_t)NKP * KST * 2 + (size_t)kVitDh * VST * 2;
};

template <int NKT>
__global__ void __launch_bounds__(kSpTcThreads)
vit_space_attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int t, int n, int heads,
                         float scale) {
  using G = SpTc<NKT>;
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smraw);       // [NKP][KST]
  __nv_bfloat16* Vt = Ks + (size_t)G::NKP * G::KST;                  // [64][VST]
  const int nk = n + 1;
  const int h = blockIdx.x, f = blockIdx.y, s = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = heads * kVitDh, T = 1 + t * n;
  const size_t ld = 3 * (size_t)D;
  const __nv_bfloat16* seg = qkv + (size_t)s * T * ld + h * kVitDh;
  for (int i = tid; i < G::NKP * 8; i += kSpTcThreads) {
    const int j = i >> 3, c = i & 7;
    uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
    if (j < nk) {
      const __nv_bfloat16* src = seg + (size_t)(j == 0 ? 0 : 1 + f * n + (j - 1)) * ld;
      kk = *reinterpret_cast<const uint4*>(src + D + 8 * c);
      vv = *reinterpret_cast<const uint4*>(src + 2 * D + 8 * c);
    }
    *reinterpret_cast<uint4*>(Ks + (size_t)j * G::KST + 8 * c) = kk;
    const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
#pragma unroll
    for (int e = 0; e < 8; ++e) Vt[(size_t)(8 * c + e) * G::VST + j] = ve[e];
  }
  __syncthreads();
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t* Kw = reinterpret_cast<const uint32_t*>(Ks);
  const uint32_t* Vw = reinterpret_cast<const uint32_t*>(Vt);
  const float sl2 = scale * 1.4426950408889634f;  // exp(x * scale) = exp2(x * scale * log2 e)
  for (int q0 = 16 * warp; q0 < n; q0 += 16 * (kSpTcThreads / 32)) {
    // Q fragments of the 16 queries q0 .. q0 + 15 (rows beyond n are clamped: computed, never stored)
    const int r0 = min(q0 + g, n - 1), r1 = min(q0 + g + 8, n - 1);
    const uint32_t* qa = reinterpret_cast<const uint32_t*>(seg + (size_t)(1 + f * n + r0) * ld);
    const uint32_t* qb = reinterpret_cast<const uint32_t*>(seg + (size_t)(1 + f * n + r1) * ld);
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      aq[ks][0] = __ldg(qa + 8 * ks + tq);
      aq[ks][1] = __ldg(qb + 8 * ks + tq);
      aq[ks][2] = __ldg(qa + 8 * ks + 4 + tq);
      aq[ks][3] = __ldg(qb + 8 * ks + 4 + tq);
    }
    float sacc[2 * NKT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * NKT; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
      const uint32_t* kr = Kw + (size_t)(8 * nt + g) * (G::KST / 2) + tq;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_bf16_16816(sacc[nt], aq[ks], kr[8 * ks], kr[8 * ks + 4]);
    }
    // softmax over the keys of rows g (values 0, 1 of a tile) and g + 8 (values 2, 3); keys >= nk are masked
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * NKT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = 8 * nt + 2 * tq + e < nk;
        if (!ok) { sacc[nt][e] = -INFINITY; sacc[nt][2 + e] = -INFINITY; }
        m0 = fmaxf(m0, sacc[nt][e]);
        m1 = fmaxf(m1, sacc[nt][2 + e]);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
    const float b0 = m0 * sl2, b1 = m1 * sl2;
#pragma unroll
    for (int nt = 0; nt < 2 * NKT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sacc[nt][e] = exp2f(fmaf(sacc[nt][e], sl2, -b0));
        sacc[nt][2 + e] = exp2f(fmaf(sacc[nt][2 + e], sl2, -b1));
        s0 += sacc[nt][e];
        s1 += sacc[nt][2 + e];
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    float oacc[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NKT; ++kt) {
      uint32_t ap[4];
      ap[0] = pack_bf16(sacc[2 * kt][0], sacc[2 * kt][1]);
      ap[1] = pack_bf16(sacc[2 * kt][2], sacc[2 * kt][3]);
      ap[2] = pack_bf16(sacc[2 * kt + 1][0], sacc[2 * kt + 1][1]);
      ap[3] = pack_bf16(sacc[2 * kt + 1][2], sacc[2 * kt + 1][3]);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const uint32_t* vr = Vw + (size_t)(8 * dt + g) * (G::VST / 2) + 8 * kt + tq;
        mma_bf16_16816(oacc[dt], ap, vr[0], vr[4]);
      }
    }
    const float i0 = 1.f / s0, i1 = 1.f / s1;
    __nv_bfloat16* o0 = out + ((size_t)s * T + 1 + f * n + q0 + g) * D + h * kVitDh + 2 * tq;
    __nv_bfloat16* o1 = o0 + (size_t)8 * D;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (q0 + g < n) *reinterpret_cast<__nv_bfloat162*>(o0 + 8 * dt) = __floats2bfloat162_rn(oacc[dt][0] * i0, oacc[dt][1] * i0);
      if (q0 + g + 8 < n) *reinterpret_cast<__nv_bfloat162*>(o1 + 8 * dt) = __floats2bfloat162_rn(oacc[dt][2] * i1, oacc[dt][3] * i1);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------------
static int device_slot() {  // shared-memory opt-in is a per-device function attribute
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}
cudaError_t launch_vit_patchify(const float* frames, void* A, int S, int C, int T, int H, int W, int tub, int ps, cudaStream_t st) {
  if (ps % 8 || W % 8 || H % ps || W % ps || T % tub) return cudaErrorInvalidValue;
  const size_t total = (size_t)S * (T / tub) * (H / ps) * (W / ps) * (C * tub * ps * ps / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  vit_patchify_kernel<<<blocks, 256, 0, st>>>(frames, reinterpret_cast<__nv_bfloat16*>(A), S, C, T, H, W, tub, ps);
  return cudaGetLastError();
}
cudaError_t launch_vit_broadcast_row(float* dst, const float* src, int D, int count, size_t row_stride, cudaStream_t st) {
  vit_broadcast_row_kernel<<<count, 256, 0, st>>>(dst, src, D, count, row_stride);
  return cudaGetLastError();
}
cudaError_t launch_vit_layernorm(const float* x, const float* g, const float* b, void* out, int rows, int D, float eps,
                                 cudaStream_t st) {
  const int blocks = (rows + 7) / 8;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (D == 768) vit_layernorm_kernel<6><<<blocks, 256, 0, st>>>(x, g, b, o, rows, eps);
  else if (D == 1024) vit_layernorm_kernel<8><<<blocks, 256, 0, st>>>(x, g, b, o, rows, eps);
  else if (D == 512) vit_layernorm_kernel<4><<<blocks, 256, 0, st>>>(x, g, b, o, rows, eps);
  else if (D == 256) vit_layernorm_kernel<2><<<blocks, 256, 0, st>>>(x, g, b, o, rows, eps);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_vit_final_norm_agg(const float* x, const float* agg_cls, const float* gf, const float* bf, const float* g1,
                                      const float* b1, void* out, int S, int t, int n, int D, float eps, cudaStream_t st) {
  const int rows = S * t * (n + 1), blocks = (rows + 7) / 8;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (D == 768) vit_final_norm_agg_kernel<6><<<blocks, 256, 0, st>>>(x, agg_cls, gf, bf, g1, b1, o, S, t, n, eps);
  else if (D == 1024) vit_final_norm_agg_kernel<8><<<blocks, 256, 0, st>>>(x, agg_cls, gf, bf, g1, b1, o, S, t, n, eps);
  else if (D == 512) vit_final_norm_agg_kernel<4><<<blocks, 256, 0, st>>>(x, agg_cls, gf, bf, g1, b1, o, S, t, n, eps);
  else if (D == 256) vit_final_norm_agg_kernel<2><<<blocks, 256, 0, st>>>(x, agg_cls, gf, bf, g1, b1, o, S, t, n, eps);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_vit_time_attn(const void* qkv, void* out, int S, int t, int n, int heads, cudaStream_t st) {
  if (t != 8) return cudaErrorInvalidValue;
  const int warps = S * n * heads, blocks = (warps + 3) / 4;
  vit_time_attn_kernel<8><<<blocks, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out),
                                                 S, n, heads, 0.125f);
  return cudaGetLastError();
}
cudaError_t launch_vit_cls_attn(const void* qkv, void* out, int seqs, int len, int heads, int out_rows_per_seq, cudaStream_t st) {
  const size_t smem = (96 + 2048 + (size_t)len) * 4;
  static bool attr[64] = {false};
  const int slot = device_slot();
  if (!attr[slot]) {
    cudaError_t e = cudaFuncSetAttribute(vit_cls_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr[slot] = true;
  }
  if (smem > 96 * 1024) return cudaErrorInvalidValue;
  vit_cls_attn_kernel<<<dim3(heads, seqs), kClsThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                    reinterpret_cast<__nv_bfloat16*>(out), len, heads, 0.125f,
                                                                    out_rows_per_seq);
  return cudaGetLastError();
}
cudaError_t launch_vit_space_attn(const void* qkv, void* out, int S, int t, int n, int heads, cudaStream_t st) {
  const bool simt = knobs().avclip_simt_attn;  // the SIMT kernel for every shape
  if (!simt && n + 1 <= 16 * 13 && n + 1 > 16 * 12) {  // 14 x 14 patches + CLS = 197 keys
    using G = SpTc<13>;
    static bool attr_tc[64] = {false};
    const int slot = device_slot();
    if (!attr_tc[slot]) {
      cudaError_t e = cudaFuncSetAttribute(vit_space_attn_tc_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem);
      if (e != cudaSuccess) return e;
      attr_tc[slot] = true;
    }
    vit_space_attn_tc_kernel<13><<<dim3(heads, t, S), kSpTcThreads, G::smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), t, n, heads, 0.125f);
    return cudaGetLastError();
  }
  const size_t smem = space_attn_smem(n);
  static bool attr[64] = {false};
  const int slot = device_slot();
  if (!attr[slot]) {
    cudaError_t e = cudaFuncSetAttribute(vit_space_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    attr[slot] = true;
  }
  if (smem > 100 * 1024) return cudaErrorInvalidValue;
  vit_space_attn_kernel<<<dim3(heads, t, S), kSpThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                     reinterpret_cast<__nv_bfloat16*>(out), t, n, heads, 0.125f);
  return cudaGetLastError();
}

}  // namespace vaura
