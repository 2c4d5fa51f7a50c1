// fp32-activation decode path (VAURA_PRECISION_FP32ACT): bf16 weights streamed once per step,
// fp32 activations / KV / accumulation on the CUDA cores.  This is the HBM-bound small-batch
// regime: arithmetic is free, bytes are not, and keeping activations in fp32 is
// what makes greedy token parity with the fp32 reference well-posed (SURVEY §7 "hard parts").
//
// Reference code replaced (paths relative to /root/reference):
//   embed_kernel        llama.py:455-472 (+ :60-73 folded tables, :555-586 repeat/pad)
//   gemv_kernel<QKV>    llama.py:153-158 (RMSNorm), :228 (wqkv), :633-650 (RoPE), KV append
//   attn_kernel         llama.py:246-255 (causal SDPA, scale 1/sqrt(96))
//   gemv_kernel<RESID>  llama.py:259 / :177 (wo, w2) + residual adds :279-283
//   gemv_kernel<SWIGLU> llama.py:176-177 (silu(w1 x) * w3 x)
//   gemv_kernel<STORE>  llama.py:503-504 (final norm + 9 heads)
#include "common.cuh"
#include "kernels.h"

namespace vaura {

// ------------------------------------------------------------------------------------------------
// embedding sum + channel concat
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_kernel(EmbedArgs a) {
  const int r = blockIdx.x;  // row = b * npos + j
  const int b = r / a.npos, j = r % a.npos;
  const int pos0 = a.state ? a.state->offset - a.npos : a.pos0;
  const int p = pos0 + j;
  const int bt = b % a.batch;  // CFG halves share the token sequence (vaura_model.py:795)
  const int C = a.cond_dim, D = a.d_model, TD = D - C;
  int vrow = p / a.atpvf;
  if (vrow > a.cond_tokens) vrow = a.cond_tokens;  // >= Tv -> empty_video_emb row (llama.py:569-572)
  const float* cr = a.cond_rows + ((size_t)b * (a.cond_tokens + 1) + vrow) * C;
  float* out = a.h + (size_t)r * D;
  __shared__ int tok[16];
  if (threadIdx.x < a.K) {
    int t = a.seq[((size_t)bt * a.K + threadIdx.x) * a.S + p];
    tok[threadIdx.x] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) out[i] = cr[i];
  for (int i = threadIdx.x; i < TD; i += blockDim.x) {
    float s = 0.f;  // python sum() starts from 0 and adds codebooks in order (llama.py:455-460)
    for (int k = 0; k < a.K; ++k) s += a.tables[((size_t)k * (a.vocab + 1) + tok[k]) * TD + i];
    out[C + i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// weight-streaming GEMV with fused RMSNorm prologue and fused epilogues
// ------------------------------------------------------------------------------------------------
// x permutation in shared memory: element k = 8c + j lives at (j < 4 ? 0 : K/2) + 4c + (j & 3), so a
// lane's two float4 reads of chunk c are each conflict-free across the warp.
template <int NB, int EPI, bool NORM>
__global__ void __launch_bounds__(256) gemv_kernel(GemvArgs a) {
  extern __shared__ float xs[];  // [NB][K]
  const int K = a.K, K8 = K >> 3, halfK = K >> 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.y * NB;
  const int nrows = min(NB, a.R - r0);

  // ---- stage x (optionally RMS-normalised): warp w stages rows w, w+8, ...
  for (int rr = warp; rr < NB; rr += 8) {
    float* xr = xs + (size_t)rr * K;
    if (rr < nrows) {
      const float* src = a.x + (size_t)(r0 + rr) * a.ldx;
      float ss = 0.f;
      for (int c = lane; c < K8; c += 32) {
        float4 lo = *reinterpret_cast<const float4*>(src + 8 * c);
        float4 hi = *reinterpret_cast<const float4*>(src + 8 * c + 4);
        if (NORM) {
          ss += lo.x * lo.x + lo.y * lo.y + lo.z * lo.z + lo.w * lo.w;
          ss += hi.x * hi.x + hi.y * hi.y + hi.z * hi.z + hi.w * hi.w;
        }
        *reinterpret_cast<float4*>(xr + 4 * c) = lo;
        *reinterpret_cast<float4*>(xr + halfK + 4 * c) = hi;
      }
      if (NORM) {
        ss = warp_sum(ss);
        const float rs = rsqrtf(ss / (float)K + a.eps);  // llama.py:153-154
        __syncwarp();
        for (int c = lane; c < K8; c += 32) {
          float4 lo = *reinterpret_cast<float4*>(xr + 4 * c);
          float4 hi = *reinterpret_cast<float4*>(xr + halfK + 4 * c);
          float4 wl = *reinterpret_cast<const float4*>(a.norm_w + 8 * c);
          float4 wh = *reinterpret_cast<const float4*>(a.norm_w + 8 * c + 4);
          // (x * rs) * w, the reference's order (llama.py:157-158)
          lo.x = lo.x * rs * wl.x; lo.y = lo.y * rs * wl.y; lo.z = lo.z * rs * wl.z; lo.w = lo.w * rs * wl.w;
          hi.x = hi.x * rs * wh.x; hi.y = hi.y * rs * wh.y; hi.z = hi.z * rs * wh.z; hi.w = hi.w * rs * wh.w;
          *reinterpret_cast<float4*>(xr + 4 * c) = lo;
          *reinterpret_cast<float4*>(xr + halfK + 4 * c) = hi;
        }
      }
    } else {
      for (int c = lane; c < K8; c += 32) {
        *reinterpret_cast<float4*>(xr + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(xr + halfK + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();

  const int pos0 = a.state ? a.state->offset - a.npos : a.pos0;
  const int npairs = a.N >> 1;
  for (int q = blockIdx.x * 8 + warp; q < npairs; q += gridDim.x * 8) {
    const uint4* w0 = reinterpret_cast<const uint4*>(a.W + (size_t)(2 * q) * K);
    const uint4* w1 = w0 + K8;
    float acc0[NB], acc1[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) acc0[r] = acc1[r] = 0.f;
#pragma unroll 2
    for (int c = lane; c < K8; c += 32) {
      const uint4 wa = ldg_stream16(w0 + c);
      const uint4 wb = ldg_stream16(w1 + c);
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        const float4 xl = *reinterpret_cast<const float4*>(xs + (size_t)r * K + 4 * c);
        const float4 xh = *reinterpret_cast<const float4*>(xs + (size_t)r * K + halfK + 4 * c);
        float s0 = acc0[r], s1 = acc1[r];
        s0 = fmaf(bf16_lo(wa.x), xl.x, s0); s0 = fmaf(bf16_hi(wa.x), xl.y, s0);
        s0 = fmaf(bf16_lo(wa.y), xl.z, s0); s0 = fmaf(bf16_hi(wa.y), xl.w, s0);
        s0 = fmaf(bf16_lo(wa.z), xh.x, s0); s0 = fmaf(bf16_hi(wa.z), xh.y, s0);
        s0 = fmaf(bf16_lo(wa.w), xh.z, s0); s0 = fmaf(bf16_hi(wa.w), xh.w, s0);
        s1 = fmaf(bf16_lo(wb.x), xl.x, s1); s1 = fmaf(bf16_hi(wb.x), xl.y, s1);
        s1 = fmaf(bf16_lo(wb.y), xl.z, s1); s1 = fmaf(bf16_hi(wb.y), xl.w, s1);
        s1 = fmaf(bf16_lo(wb.z), xh.x, s1); s1 = fmaf(bf16_hi(wb.z), xh.y, s1);
        s1 = fmaf(bf16_lo(wb.w), xh.z, s1); s1 = fmaf(bf16_hi(wb.w), xh.w, s1);
        acc0[r] = s0; acc1[r] = s1;
      }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) { acc0[r] = warp_sum(acc0[r]); acc1[r] = warp_sum(acc1[r]); }

    const int n = 2 * q;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (lane != r || r >= nrows) continue;
      const int row = r0 + r;
      const float y0 = acc0[r], y1 = acc1[r];
      if (EPI == EPI_STORE) {
        size_t o = (size_t)row * a.ldo + n;
        if (a.perm_S > 0) {
          const int b = row / a.perm_S, j = row % a.perm_S, kk = n / a.perm_V;
          o = (((size_t)b * (a.N / a.perm_V) + kk) * a.perm_S + j) * a.perm_V + (n % a.perm_V);
        }
        a.out[o] = y0;
        a.out[o + 1] = y1;
      } else if (EPI == EPI_RESID) {
        float* o = a.out + (size_t)row * a.ldo + n;  // in-place residual stream
        o[0] += y0;
        o[1] += y1;
      } else if (EPI == EPI_SWIGLU) {
        // rows 2j / 2j+1 are w1[j] / w3[j]: silu(w1 x) * (w3 x)  (llama.py:177)
        const float s = y0 / (1.f + expf(-y0));
        a.out[(size_t)row * a.ldo + q] = s * y1;
      } else {  // EPI_QKV: RoPE on q,k (adjacent pairs, llama.py:633-650) + KV append
        const int D = a.d_model;
        const int sec = n / D, within = n % D;
        const int hd = within / kHeadDim, e = within % kHeadDim;
        const int b = row / a.npos, j = row % a.npos;
        const int p = pos0 + j;
        if (sec == 2) {
          float* v = reinterpret_cast<float*>(a.kv.pages) + a.kv.row(a.layer, 1, b, p, hd) + e;
          v[0] = y0;
          v[1] = y1;
        } else {
          const float2 cs = *reinterpret_cast<const float2*>(a.rope + ((size_t)p * (kHeadDim / 2) + (e >> 1)) * 2);
          const float o0 = y0 * cs.x - y1 * cs.y;
          const float o1 = y1 * cs.x + y0 * cs.y;
          if (sec == 0) {
            a.out[(size_t)row * a.ldo + within] = o0;
            a.out[(size_t)row * a.ldo + within + 1] = o1;
          } else {
            float* kk = reinterpret_cast<float*>(a.kv.pages) + a.kv.row(a.layer, 0, b, p, hd) + e;
            kk[0] = o0;
            kk[1] = o1;
          }
        }
      }
    }
  }
}

constexpr int kGemvMaxSmem = 200 * 1024;

template <int NB, int EPI, bool NORM>
static cudaError_t launch_gemv_t(const GemvArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)NB * a.K * sizeof(float);
  if (smem > (size_t)kGemvMaxSmem) return cudaErrorInvalidValue;
  const int npairs = a.N / 2;
  int gx = (npairs + 7) / 8;
  const int cap = 148 * 8;
  if (gx > cap) gx = cap;
  dim3 grid(gx, (a.R + NB - 1) / NB);
  gemv_kernel<NB, EPI, NORM><<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

template <int EPI, bool NORM>
static cudaError_t init_gemv_epi() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(gemv_kernel<1, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem))) return e;
  if ((e = cudaFuncSetAttribute(gemv_kernel<2, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem))) return e;
  if ((e = cudaFuncSetAttribute(gemv_kernel<4, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem))) return e;
  return cudaFuncSetAttribute(gemv_kernel<8, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem);
}

// opt every instantiation into large dynamic shared memory once, outside any stream capture
cudaError_t init_decode_kernels() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e;
  if ((e = init_gemv_epi<EPI_QKV, true>())) return e;
  if ((e = init_gemv_epi<EPI_SWIGLU, true>())) return e;
  if ((e = init_gemv_epi<EPI_RESID, false>())) return e;
  if ((e = init_gemv_epi<EPI_STORE, true>())) return e;
  if ((e = init_gemv_epi<EPI_STORE, false>())) return e;
  done = true;
  return cudaSuccess;
}

template <int EPI, bool NORM>
static cudaError_t launch_gemv_nb(const GemvArgs& a, cudaStream_t st) {
  // rows per weight pass: as many as fit the register/smem budget (<= 8); K=4096 caps smem at NB<=8 (128 KB)
  if (a.R <= 1) return launch_gemv_t<1, EPI, NORM>(a, st);
  if (a.R <= 2) return launch_gemv_t<2, EPI, NORM>(a, st);
  if (a.R <= 4) return launch_gemv_t<4, EPI, NORM>(a, st);
  return launch_gemv_t<8, EPI, NORM>(a, st);
}

cudaError_t launch_gemv(int epi, bool norm, const GemvArgs& a, cudaStream_t st) {
  if (a.K % 8 != 0 || a.N % 2 != 0) return cudaErrorInvalidValue;
  switch (epi) {
    case EPI_QKV: return launch_gemv_nb<EPI_QKV, true>(a, st);
    case EPI_SWIGLU: return launch_gemv_nb<EPI_SWIGLU, true>(a, st);
    case EPI_RESID: return launch_gemv_nb<EPI_RESID, false>(a, st);
    case EPI_STORE: return norm ? launch_gemv_nb<EPI_STORE, true>(a, st) : launch_gemv_nb<EPI_STORE, false>(a, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_embed(const EmbedArgs& a, int rows, cudaStream_t st) {
  embed_kernel<<<rows, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// causal attention over the paged fp32 KV cache: one CTA per (head, query row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_kernel(AttnArgs a) {
  __shared__ float qs[kHeadDim];
  __shared__ float sc[kMaxCtx];
  __shared__ float red[4];
  const int h = blockIdx.x, row = blockIdx.y;
  const int b = row / a.npos, j = row % a.npos;
  const int pos0 = a.state ? a.state->offset - a.npos : a.pos0;
  const int p = pos0 + j, nctx = p + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* kvp = reinterpret_cast<const float*>(a.kv.pages);
  if (tid < kHeadDim) qs[tid] = a.q[(size_t)row * a.d_model + h * kHeadDim + tid];
  __syncthreads();

  // scores: 8 lanes per key position, 12 dims each
  const int g = tid >> 3, t = tid & 7;
  float q12[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) q12[i] = qs[t * 12 + i];
  const int nround = (nctx + 15) & ~15;  // every lane of a warp runs the same trip count (shuffles below)
  for (int jj = g; jj < nround; jj += 16) {
    float s = 0.f;
    if (jj < nctx) {
      const float4* kr = reinterpret_cast<const float4*>(kvp + a.kv.row(a.layer, 0, b, jj, h) + t * 12);
      const float4 k0 = kr[0], k1 = kr[1], k2 = kr[2];
      s = q12[0] * k0.x + q12[1] * k0.y + q12[2] * k0.z + q12[3] * k0.w + q12[4] * k1.x + q12[5] * k1.y +
          q12[6] * k1.z + q12[7] * k1.w + q12[8] * k2.x + q12[9] * k2.y + q12[10] * k2.z + q12[11] * k2.w;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
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

  if (tid < kHeadDim) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int jj = 0;
    for (; jj + 3 < nctx; jj += 4) {
      a0 = fmaf(sc[jj], kvp[a.kv.row(a.layer, 1, b, jj, h) + tid], a0);
      a1 = fmaf(sc[jj + 1], kvp[a.kv.row(a.layer, 1, b, jj + 1, h) + tid], a1);
      a2 = fmaf(sc[jj + 2], kvp[a.kv.row(a.layer, 1, b, jj + 2, h) + tid], a2);
      a3 = fmaf(sc[jj + 3], kvp[a.kv.row(a.layer, 1, b, jj + 3, h) + tid], a3);
    }
    for (; jj < nctx; ++jj) a0 = fmaf(sc[jj], kvp[a.kv.row(a.layer, 1, b, jj, h) + tid], a0);
    const float o = ((a0 + a1) + (a2 + a3)) / sum;
    a.out[(size_t)row * a.d_model + h * kHeadDim + tid] = o;
    if (a.out3 && a.out_terms == 1) {  // bf16-operand prefill: the row rounded once
      reinterpret_cast<__nv_bfloat16*>(a.out3)[(size_t)row * a.d_model + h * kHeadDim + tid] = __float2bfloat16_rn(o);
    } else if (a.out3) {  // the wo GEMM of the tensor-core prefill reads the row as three bf16 terms
      __nv_bfloat16 t1, t2, t3;
      split3(o, t1, t2, t3);
      __nv_bfloat16* o3 = reinterpret_cast<__nv_bfloat16*>(a.out3) + (size_t)row * 3 * a.d_model + h * kHeadDim + tid;
      o3[0] = t1; o3[a.d_model] = t2; o3[2 * a.d_model] = t3;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// causal attention of a multi-position pass (prompt prefill, teacher-forced forward) over the paged fp32 KV cache:
// one CTA per (head, 8 QW consecutive query positions of one sequence), QW = queries per warp.  Keys / values are staged 32
// positions at a time in shared memory and shared by the CTA's queries (attn_kernel re-reads the whole context per query:
// 166 x the K/V traffic for a 166-position prompt); flash-style online softmax in fp32 (llama.py:246-255, scale 1/sqrt(96)).
// Warp w owns queries QW w .. QW w + QW - 1: scores with lane = key (K rows padded to 97 floats: conflict-free), P.V with
// lane = output dims lane, lane + 32, lane + 64.
// The kernel is bound by its own instruction latency (ncu, 167 positions, QW = 4: 96 CTAs = two warps per scheduler on 96 SMs,
// issue slots 36 % busy, short-scoreboard and wait stalls, 53.6 k cycles = the CTA of the last query block), so a single
// prompt window takes fewer queries per warp - more CTAs, shorter chains, more warps per scheduler to hide the shared-memory
// latency - and the query blocks are issued last-first: the blocks that see the most keys start first on SMs of their own
// (25.8 -> 18.8 us per launch).
// ------------------------------------------------------------------------------------------------
constexpr int kPfK = 32, kPfThreads = 256;
template <int QW>
__global__ void __launch_bounds__(kPfThreads) attn_prefill_kernel(AttnArgs a) {
  constexpr int kPfQ = 8 * QW;
  __shared__ float qs[kPfQ][kHeadDim];
  __shared__ float ks[kPfK][kHeadDim + 1];
  __shared__ float vs[kPfK][kHeadDim];
  __shared__ float ps[kPfThreads / 32][kPfK][QW];
  const int h = blockIdx.x, j0 = ((int)gridDim.y - 1 - (int)blockIdx.y) * kPfQ, b = blockIdx.z;
  const int pos0 = a.state ? a.state->offset - a.npos : a.pos0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = min(kPfQ, a.npos - j0);
  const float* kvp = reinterpret_cast<const float*>(a.kv.pages);
  for (int i = tid; i < kPfQ * (kHeadDim / 4); i += kPfThreads) {
    const int qi = i / (kHeadDim / 4), c = i % (kHeadDim / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qi < nq) v = *reinterpret_cast<const float4*>(a.q + (size_t)(b * a.npos + j0 + qi) * a.d_model + h * kHeadDim + 4 * c);
    *reinterpret_cast<float4*>(&qs[qi][4 * c]) = v;
  }
  float m[QW], l[QW], acc[QW][3];
#pragma unroll
  for (int u = 0; u < QW; ++u) { m[u] = -INFINITY; l[u] = 0.f; acc[u][0] = acc[u][1] = acc[u][2] = 0.f; }
  const int last_pos = pos0 + j0 + nq - 1;  // last key any query of this block may see
  // key / value rows of a tile travel through registers: 8 threads per position, 3 float4 each; the rows of tile k0 + 32 are
  // requested before tile k0 is consumed, so their latency overlaps its arithmetic
  const int sr = tid >> 3, scol = tid & 7;
  float4 kreg[3], vreg[3];
  auto fetch = [&](int k0) {
    const int pos = k0 + sr;
    if (pos <= last_pos) {
      const float4* kr = reinterpret_cast<const float4*>(kvp + a.kv.row(a.layer, 0, b, pos, h));
      const float4* vr = reinterpret_cast<const float4*>(kvp + a.kv.row(a.layer, 1, b, pos, h));
#pragma unroll
      for (int i = 0; i < 3; ++i) { kreg[i] = kr[scol + 8 * i]; vreg[i] = vr[scol + 8 * i]; }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 <= last_pos; k0 += kPfK) {
    __syncthreads();  // previous tile consumed (and qs written, first iteration)
    if (k0 + sr <= last_pos) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float* kd = &ks[sr][4 * (scol + 8 * i)];
        kd[0] = kreg[i].x; kd[1] = kreg[i].y; kd[2] = kreg[i].z; kd[3] = kreg[i].w;
        *reinterpret_cast<float4*>(&vs[sr][4 * (scol + 8 * i)]) = vreg[i];
      }
    }
    if (k0 + kPfK <= last_pos) fetch(k0 + kPfK);
    __syncthreads();
    // scores of this warp's queries against key k0 + lane
    float sc[QW];
#pragma unroll
    for (int u = 0; u < QW; ++u) sc[u] = 0.f;
    const int kpos = k0 + lane;
#pragma unroll 4
    for (int d4 = 0; d4 < kHeadDim / 4; ++d4) {
      const float k0v = ks[lane][4 * d4], k1v = ks[lane][4 * d4 + 1], k2v = ks[lane][4 * d4 + 2], k3v = ks[lane][4 * d4 + 3];
#pragma unroll
      for (int u = 0; u < QW; ++u) {
        const float4 qv = *reinterpret_cast<const float4*>(&qs[QW * warp + u][4 * d4]);
        sc[u] = fmaf(qv.x, k0v, sc[u]); sc[u] = fmaf(qv.y, k1v, sc[u]);
        sc[u] = fmaf(qv.z, k2v, sc[u]); sc[u] = fmaf(qv.w, k3v, sc[u]);
      }
    }
    float corr[QW];
#pragma unroll
    for (int u = 0; u < QW; ++u) {
      const int qpos = pos0 + j0 + QW * warp + u;
      const bool ok = kpos <= qpos && QW * warp + u < nq;
      const float sv = ok ? sc[u] * a.scale : -INFINITY;
      const float mn = fmaxf(m[u], warp_max(sv));
      const float pe = ok ? expf(sv - mn) : 0.f;           // mn is finite whenever ok (the key itself is visible)
      corr[u] = m[u] == -INFINITY ? 0.f : expf(m[u] - mn);  // no key seen yet: nothing to rescale
      if (mn == -INFINITY) corr[u] = 0.f;
      l[u] = l[u] * corr[u] + warp_sum(pe);
      m[u] = mn;
      ps[warp][lane][u] = pe;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < QW; ++u) { acc[u][0] *= corr[u]; acc[u][1] *= corr[u]; acc[u][2] *= corr[u]; }
    const int nk = min(kPfK, last_pos - k0 + 1);
    for (int j = 0; j < nk; ++j) {
      const float v0 = vs[j][lane], v1 = vs[j][lane + 32], v2 = vs[j][lane + 64];
      float pw[QW];
      if constexpr (QW == 4) {
        const float4 pv = *reinterpret_cast<const float4*>(&ps[warp][j][0]);
        pw[0] = pv.x; pw[1] = pv.y; pw[2] = pv.z; pw[3] = pv.w;
      } else if constexpr (QW == 2) {
        const float2 pv = *reinterpret_cast<const float2*>(&ps[warp][j][0]);
        pw[0] = pv.x; pw[1] = pv.y;
      } else {
        pw[0] = ps[warp][j][0];
      }
#pragma unroll
      for (int u = 0; u < QW; ++u) {
        acc[u][0] = fmaf(pw[u], v0, acc[u][0]); acc[u][1] = fmaf(pw[u], v1, acc[u][1]); acc[u][2] = fmaf(pw[u], v2, acc[u][2]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int u = 0; u < QW; ++u) {
    const int qi = QW * warp + u;
    if (qi >= nq) continue;
    const size_t row = (size_t)b * a.npos + j0 + qi;
    const float inv = 1.f / l[u];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float o = acc[u][i] * inv;
      const int col = h * kHeadDim + lane + 32 * i;
      a.out[row * a.d_model + col] = o;
      if (a.out3 && a.out_terms == 1) {
        reinterpret_cast<__nv_bfloat16*>(a.out3)[row * a.d_model + col] = __float2bfloat16_rn(o);
      } else if (a.out3) {
        __nv_bfloat16 t1, t2, t3;
        split3(o, t1, t2, t3);
        __nv_bfloat16* o3 = reinterpret_cast<__nv_bfloat16*>(a.out3) + row * 3 * a.d_model + col;
        o3[0] = t1; o3[a.d_model] = t2; o3[2 * a.d_model] = t3;
      }
    }
  }
}

cudaError_t launch_attn(const AttnArgs& a, int nhead, int rows, cudaStream_t st) {
  if (a.npos >= 16 && rows % a.npos == 0) {  // multi-position pass: keys / values shared by the 8 QW queries of a CTA
    const int seqs = rows / a.npos;
    auto ctas = [&](int qw) { return (long long)nhead * ((a.npos + 8 * qw - 1) / (8 * qw)) * seqs; };
    // four queries per warp (fewest instructions in total) when that still gives every SM two CTAs; a single prompt window
    // does not: two per warp.  Measured at 167 positions, one sequence: 23.9 / 18.8 / 19.5 us per launch with 4 / 2 / 1 queries
    // per warp (25.8 us before the blocks were issued last-first)
    int qw = ctas(4) >= 296 ? 4 : 2;
    if (knobs().prefill_attn_qw == 1 || knobs().prefill_attn_qw == 2 || knobs().prefill_attn_qw == 4) qw = knobs().prefill_attn_qw;
    const dim3 grid(nhead, (a.npos + 8 * qw - 1) / (8 * qw), seqs);
    if (qw == 4) attn_prefill_kernel<4><<<grid, kPfThreads, 0, st>>>(a);
    else if (qw == 2) attn_prefill_kernel<2><<<grid, kPfThreads, 0, st>>>(a);
    else attn_prefill_kernel<1><<<grid, kPfThreads, 0, st>>>(a);
    return cudaGetLastError();
  }
  attn_kernel<<<dim3(nhead, rows), 128, 0, st>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conditioning MLP (once per clip): rows_out[r][t] = fc2(gelu_tanh(fc1(feats[r][t])))  llama.py:79-92
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cond_project_kernel(const float* __restrict__ feats, const float* __restrict__ fc1,
                                                           const float* __restrict__ fc2, const float* __restrict__ empty,
                                                           float* __restrict__ out, int tv, int cin, int C) {
  extern __shared__ float sm[];  // x[cin] + hid[C]
  float* x = sm;
  float* hid = sm + cin;
  const int r = blockIdx.y, t = blockIdx.x;
  float* o = out + ((size_t)r * (tv + 1) + t) * C;
  if (t == tv) {  // appended empty_video_emb row (llama.py:336-338, :569-572)
    for (int i = threadIdx.x; i < C; i += blockDim.x) o[i] = empty[i];
    return;
  }
  const float* f = feats + ((size_t)r * tv + t) * cin;
  for (int i = threadIdx.x; i < cin; i += blockDim.x) x[i] = f[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n = warp; n < C; n += 8) {
    float s = 0.f;
    for (int k = lane; k < cin; k += 32) s = fmaf(fc1[(size_t)n * cin + k], x[k], s);
    s = warp_sum(s);
    if (lane == 0) {
      // GELU tanh approximation (llama.py:85)
      const float c0 = 0.7978845608028654f, c1 = 0.044715f;
      hid[n] = 0.5f * s * (1.f + tanhf(c0 * (s + c1 * s * s * s)));
    }
  }
  __syncthreads();
  for (int n = warp; n < C; n += 8) {
    float s = 0.f;
    for (int k = lane; k < C; k += 32) s = fmaf(fc2[(size_t)n * C + k], hid[k], s);
    s = warp_sum(s);
    if (lane == 0) o[n] = s;
  }
}

// Same MLP, kCondRows feature rows per cluster of kCondSplit CTAs: every weight row read from L2 is multiplied with all of the
// cluster's rows (the one-row-per-CTA version above re-streams fc1 + fc2 = 2.5 MB per row: 493 us for one clip, 818 us for 64),
// and the output features of both layers are split over the cluster's CTAs so that one clip still spreads over 32 SMs; the
// hidden slices meet in every CTA's shared memory through DSMEM stores and one cluster barrier.  Same arithmetic per output (one
// fmaf chain per lane over k = lane, lane + 32, ..., then the warp butterfly): bit-identical results.
constexpr int kCondRows = 8, kCondSplit = 8;
__global__ void __cluster_dims__(kCondSplit, 1, 1) __launch_bounds__(256)
cond_project_rows_kernel(const float* __restrict__ feats, const float* __restrict__ fc1, const float* __restrict__ fc2,
                         const float* __restrict__ empty, float* __restrict__ out, int rows, int tv, int cin, int C) {
  extern __shared__ float sm[];  // x[kCondRows][cin] + hid[kCondRows][C]
  float* x = sm;
  float* hid = sm + kCondRows * cin;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tile = blockIdx.x / kCondSplit, m0 = tile * kCondRows, M = rows * tv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npc = C / kCondSplit, n_lo = (int)rank * npc;  // output features of this CTA in both layers
  if (rank == 0 && tile < rows)  // appended empty_video_emb row of clip row `tile` (llama.py:336-338, :569-572)
    for (int i = threadIdx.x; i < C; i += blockDim.x) out[((size_t)tile * (tv + 1) + tv) * C + i] = empty[i];
  for (int i = threadIdx.x; i < kCondRows * cin; i += blockDim.x) {
    const int j = i / cin, m = m0 + j;
    x[i] = m < M ? feats[(size_t)m * cin + (i - j * cin)] : 0.f;
  }
  __syncthreads();
  const uint32_t hid_base = (uint32_t)__cvta_generic_to_shared(hid);
  for (int n = n_lo + warp; n < n_lo + npc; n += 8) {
    float s[kCondRows];
#pragma unroll
    for (int j = 0; j < kCondRows; ++j) s[j] = 0.f;
    for (int k = lane; k < cin; k += 32) {
      const float w = __ldg(fc1 + (size_t)n * cin + k);
#pragma unroll
      for (int j = 0; j < kCondRows; ++j) s[j] = fmaf(w, x[j * cin + k], s[j]);
    }
#pragma unroll
    for (int j = 0; j < kCondRows; ++j) s[j] = warp_sum(s[j]);
    // lane = (row j, destination CTA d): every CTA of the cluster gets the hidden value of (row j, feature n)
    const int j = lane & (kCondRows - 1), d = lane / kCondRows;
    float v = s[0];
#pragma unroll
    for (int jj = 1; jj < kCondRows; ++jj) v = j == jj ? s[jj] : v;
    const float c0 = 0.7978845608028654f, c1 = 0.044715f;  // GELU tanh approximation (llama.py:85)
    const float h = 0.5f * v * (1.f + tanhf(c0 * (v + c1 * v * v * v)));
    for (int dd = d; dd < kCondSplit; dd += 32 / kCondRows) {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(hid_base + (uint32_t)(j * C + n) * 4u), "r"((uint32_t)dd));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(h) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  for (int n = n_lo + warp; n < n_lo + npc; n += 8) {
    float s[kCondRows];
#pragma unroll
    for (int j = 0; j < kCondRows; ++j) s[j] = 0.f;
    for (int k = lane; k < C; k += 32) {
      const float w = __ldg(fc2 + (size_t)n * C + k);
#pragma unroll
      for (int j = 0; j < kCondRows; ++j) s[j] = fmaf(w, hid[j * C + k], s[j]);
    }
#pragma unroll
    for (int j = 0; j < kCondRows; ++j) s[j] = warp_sum(s[j]);
    if (lane < kCondRows && m0 + lane < M) {
      float v = s[0];
#pragma unroll
      for (int j = 1; j < kCondRows; ++j) v = lane == j ? s[j] : v;
      const int m = m0 + lane;
      out[((size_t)(m / tv) * (tv + 1) + m % tv) * C + n] = v;
    }
  }
}

cudaError_t launch_cond_project(const float* feats, const float* fc1, const float* fc2, const float* empty, float* out,
                                int rows, int tv, int cin, int C, cudaStream_t st) {
  const size_t smem = (size_t)kCondRows * (cin + C) * sizeof(float);
  const int tiles = (rows * tv + kCondRows - 1) / kCondRows;
  if (smem <= 48 * 1024 && tiles >= rows && C % (kCondSplit * 8) == 0) {
    cond_project_rows_kernel<<<tiles * kCondSplit, 256, smem, st>>>(feats, fc1, fc2, empty, out, rows, tv, cin, C);
    return cudaGetLastError();
  }
  cond_project_kernel<<<dim3(tv + 1, rows), 256, (cin + C) * sizeof(float), st>>>(feats, fc1, fc2, empty, out, tv, cin, C);
  return cudaGetLastError();
}

}  // namespace vaura
