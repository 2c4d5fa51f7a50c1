// probe: latency of a cluster-wide exchange (DSMEM push + barrier) with and without concurrent cp.async.bulk traffic
// into the same SMs.  Variants: software barrier (remote mbarrier arrive) and hardware barrier.cluster.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int CL = 4, SLOT = 12288, NSLOT = 12, ITERS = 200;
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par, bool cluster) {
  uint32_t ok;
  if (cluster) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  else asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
// mode bit0: bulk traffic on, bit1: hardware cluster barrier, bit2: no data push (barrier only)
__global__ void __launch_bounds__(416, 1) k(const uint8_t* src, size_t per_cta, int mode, long long* out, int pace, int nsl, unsigned* gbar) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSLOT * SLOT);  // full[NSLOT], cbar
  float* recv = reinterpret_cast<float*>(smem + NSLOT * SLOT + 256);  // [4][288]
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t full0 = s_u32(bars), cbar = full0 + NSLOT * 8, xbar = cbar + 8;
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * i));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cbar), "r"(CL));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  __shared__ volatile int stop;
  if (tid == 0) stop = 0;
  const bool hw = mode & 2;
  if (warp == 12 && !hw) {
    if ((mode & 1) && (tid & 31) == 0) {  // stream bulk copies round-robin over the slots until told to stop
      const uint8_t* p = src + (size_t)blockIdx.x * per_cta;
      size_t off = 0;
      uint32_t ph[NSLOT] = {0};
      long long next = 0;
      for (int it = 0; !stop; ++it) {
        const int s = it % nsl;
        if (it >= nsl) { while (!try_wait(full0 + 8 * s, ph[s], false)) {} ph[s] ^= 1; }
        if (pace) { long long now = clock64(); while (now < next) now = clock64(); next = now + pace; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"(SLOT) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(smem + s * SLOT)), "l"(p + off), "r"(SLOT), "r"(full0 + 8 * s) : "memory");
        off += SLOT; if (off + SLOT > per_cta) off = 0;
      }
      // drain
      for (int s = 0; s < NSLOT; ++s) { /* outstanding copies complete before exit: wait on each armed barrier */ }
    }
  } else if (warp < 12 || hw) {
    uint32_t phase = 0;
    long long tsum = 0, tmax = 0;
    for (int it = 0; it < ITERS; ++it) {
      // spacing between exchanges
      long long t0 = clock64();
      while (clock64() - t0 < 3000) {}
      if (hw) __syncthreads(); else asm volatile("bar.sync 1, 384;" ::: "memory");
      t0 = clock64();
      if (mode & 64) {
        const uint32_t xb = xbar + 8 * (it & 1);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xb), "r"(288 * 4) : "memory");
        if (tid < 288) {
          const uint32_t dst = mapa(s_u32(recv + rank * 288 + tid), tid / 72);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(dst), "r"(it), "r"(mapa(xb, tid / 72)) : "memory");
        }
        while (!try_wait(xb, (it >> 1) & 1, false)) {}
        if (tid < 288 && reinterpret_cast<volatile int*>(recv)[(tid / 72) * 288 + 72 * rank + tid % 72] != it) asm volatile("trap;");
      } else if (mode & 128) {
        asm volatile("bar.sync 1, 384;" ::: "memory");
        if (tid == 0) {
          unsigned* ctr = gbar + (blockIdx.x / CL) * 32;
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          const unsigned target = (unsigned)(it + 1) * CL;
          unsigned v;
          do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while ((int)(v - target) < 0);
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");
      } else
      if (!(mode & 4) && !(mode & 8) && tid < 288) {
        const uint32_t dst = mapa(s_u32(recv + rank * 288 + tid), tid / 72);
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"((float)it) : "memory");
      }
      if (mode & (64 | 128)) {
      } else if (hw) {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      } else if (mode & 8) {
        asm volatile("bar.sync 1, 384;" ::: "memory");
        if (tid == 0) {
          unsigned* ctr = gbar + (blockIdx.x / CL) * 32;
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          const unsigned target = (unsigned)(it + 1) * CL;
          unsigned v;
          do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while ((int)(v - target) < 0);
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");
      } else {
        asm volatile("bar.sync 1, 384;" ::: "memory");
        if (tid < CL) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(cbar, tid)) : "memory");
        if (mode & 16) {
          if (tid == 0) while (!try_wait(cbar, phase, true)) {}
          asm volatile("bar.sync 1, 384;" ::: "memory");
        } else if (mode & 32) {
          if ((tid & 31) == 0) while (!try_wait(cbar, phase, true)) {}
          __syncwarp();
        } else {
          while (!try_wait(cbar, phase, true)) {}
        }
        phase ^= 1;
      }
      const long long dt = clock64() - t0;
      tsum += dt; if (dt > tmax) tmax = dt;
    }
    if (tid == 0) { out[blockIdx.x * 2] = tsum / ITERS; out[blockIdx.x * 2 + 1] = tmax; stop = 1; }
  }
  __syncthreads();
  // let outstanding bulk copies land before the CTA exits
  { long long t0 = clock64(); while (clock64() - t0 < 20000) {} }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  const size_t per_cta = 8u << 20;
  uint8_t* src; long long* out;
  cudaMalloc(&src, per_cta * 128); cudaMemset(src, 1, per_cta * 128);
  cudaMalloc(&out, 128 * 2 * sizeof(long long));
  const int smem = NSLOT * SLOT + 256 + 4 * 288 * 4;  // bars: full[12] cbar xbar[2] = 120 B
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned* gbar; cudaMalloc(&gbar, 32 * 32 * 4);
  struct Cfg { int pace, mode, nsl; };
  const Cfg cfgs[] = {{0, 4, 12}, {0, 5, 12}, {0, 64, 12}, {0, 65, 12}, {0, 128, 12}, {0, 129, 12}, {0, 5, 6}, {0, 5, 3}, {0, 5, 1}, {1000, 5, 12}, {2000, 5, 12}, {4000, 5, 12}, {0, 12, 12}, {0, 13, 12}, {500, 13, 12}, {0, 13, 3}};
  for (const Cfg& cf : cfgs) {
    const int pace = cf.pace, mode = cf.mode;
    cudaMemset(gbar, 0, 32 * 32 * 4);
    cudaLaunchConfig_t c{}; c.gridDim = dim3(128); c.blockDim = dim3(416); c.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    c.attrs = a; c.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&c, k, (const uint8_t*)src, per_cta, mode, out, pace, cf.nsl, gbar);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long avg = 0, mx = 0; for (int i = 0; i < 128; ++i) { avg += h[2 * i]; if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; }
    printf("pace %4d slots in flight %2d mode %2d (%s%s): avg %lld cycles, max %lld cycles  [%s %s]\n", pace, cf.nsl, mode, (mode & 1) ? "bulk traffic, " : "idle, ",
           (mode & 8) ? "global-memory barrier" : (mode & 64) ? "st.async exchange" : (mode & 128) ? "relaxed global barrier" : "dsmem sw barrier", avg / 128, mx, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  return 0;
}
