// Achievable HBM read rate of per-warp bulk copies (cp.async.bulk global -> shared) as the decode kernels issue them:
// 148 CTAs x W warps, every warp streams its own sequence of CHUNK-byte pieces through a ring of SLOTS slots and does
// nothing else.  Reports TB/s for a cold 1.5 GB buffer (>> L2) per chunk size / warp count / ring depth, and for a
// short burst (50 MB, the K/V bytes of one layer at 64 rows) including the start-up latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_read_rate bulk_read_rate.cu && ./bulk_read_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CHUNK, int SLOTS>
__global__ void __launch_bounds__(320, 1) reader(const uint8_t* src, size_t chunks_per_warp, size_t warp_stride_chunks, int warps) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024 + (size_t)warp * SLOTS * CHUNK;
  if (threadIdx.x == 0)
    for (int i = 0; i < 10 * SLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (warp >= warps) return;
  uint64_t* bar = bars + warp * SLOTS;
  // chunk c of this warp: interleaved across all warps of the grid like (row, head) items are
  const size_t w = (size_t)blockIdx.x * warps + warp;
  const uint8_t* base = src + w * warp_stride_chunks * CHUNK;
  size_t issued = 0;
  unsigned acc = 0;
  for (size_t c = 0; c < chunks_per_warp; ++c) {
    while (issued < chunks_per_warp && issued < c + SLOTS) {
      if (lane == 0) {
        const uint32_t b = smem_u32(&bar[issued % SLOTS]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + (issued % SLOTS) * CHUNK)),
                     "l"(base + issued * CHUNK), "r"(CHUNK), "r"(b)
                     : "memory");
      }
      ++issued;
    }
    const uint32_t b = smem_u32(&bar[c % SLOTS]);
    const uint32_t parity = (uint32_t)(c / SLOTS) & 1u;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(b), "r"(parity)
                   : "memory");
    acc += ring[(c % SLOTS) * CHUNK + lane * 16];
    __syncwarp();
  }
  if (acc == 0xffffffffu) printf("never\n");
}

// every CTA re-reads the same `region` bytes (L2 hits after the first pass), like the activation matrix of a GEMM phase
template <int CHUNK, int SLOTS>
__global__ void __launch_bounds__(320, 1) reader_same(const uint8_t* src, size_t chunks_per_warp, int passes, int warps) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024 + (size_t)warp * SLOTS * CHUNK;
  if (threadIdx.x == 0)
    for (int i = 0; i < 10 * SLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (warp >= warps) return;
  uint64_t* bar = bars + warp * SLOTS;
  const uint8_t* base = src + (size_t)warp * chunks_per_warp * CHUNK;  // the same for every CTA
  const size_t total = chunks_per_warp * passes;
  // start at a CTA-dependent offset so that the CTAs do not walk the region in lock step
  const size_t start = ((size_t)blockIdx.x * 7919) % chunks_per_warp;
  size_t issued = 0;
  unsigned acc = 0;
  for (size_t c = 0; c < total; ++c) {
    while (issued < total && issued < c + SLOTS) {
      if (lane == 0) {
        const uint32_t b = smem_u32(&bar[issued % SLOTS]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + (issued % SLOTS) * CHUNK)),
                     "l"(base + ((issued + start) % chunks_per_warp) * CHUNK), "r"(CHUNK), "r"(b)
                     : "memory");
      }
      ++issued;
    }
    const uint32_t b = smem_u32(&bar[c % SLOTS]);
    const uint32_t parity = (uint32_t)(c / SLOTS) & 1u;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(b), "r"(parity)
                   : "memory");
    acc += ring[(c % SLOTS) * CHUNK + lane * 16];
    __syncwarp();
  }
  if (acc == 0xffffffffu) printf("never\n");
}

template <int CHUNK, int SLOTS>
static void run_l2(const uint8_t* buf, size_t region, int passes, int warps, const char* what) {
  const int smem = 1024 + warps * SLOTS * CHUNK;
  cudaFuncSetAttribute(reader_same<CHUNK, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // warp w of a CTA reads region / warps bytes, `passes` times over (chunks_per_warp wraps by launching passes times)
  const size_t cpw = region / CHUNK / warps;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  reader_same<CHUNK, SLOTS><<<148, 320, smem>>>(buf, cpw, passes, warps);
  cudaEventRecord(e0);
  reader_same<CHUNK, SLOTS><<<148, 320, smem>>>(buf, cpw, passes, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)cpw * CHUNK * warps * 148 * passes;
  printf("%-28s chunk %5d B  slots %2d  warps %2d : %8.1f us  %6.2f TB/s delivered to the SMs (region %.1f MB x %d passes)\n", what, CHUNK,
         SLOTS, warps, ms * 1e3, bytes / (ms * 1e-3) / 1e12, region / 1e6, passes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
}

template <int CHUNK, int SLOTS>
static void run(const uint8_t* buf, size_t total_bytes, int warps, const char* what) {
  const int smem = 1024 + warps * SLOTS * CHUNK;
  cudaFuncSetAttribute(reader<CHUNK, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const size_t nw = (size_t)148 * warps;
  const size_t cpw = total_bytes / CHUNK / nw;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    reader<CHUNK, SLOTS><<<148, 320, smem>>>(buf, cpw, cpw, warps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double bytes = (double)cpw * CHUNK * nw;
  printf("%-28s chunk %5d B  slots %2d  warps %2d : %8.1f us  %6.2f TB/s  (%.0f MB)\n", what, CHUNK, SLOTS, warps, best * 1e3,
         bytes / (best * 1e-3) / 1e12, bytes / 1e6);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
}

int main() {
  const size_t big = (size_t)1536 << 20;
  uint8_t* buf;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  // long stream (start-up amortised)
  run<3072, 6>(buf, big, 7, "stream 1.5 GB");
  run<3072, 6>(buf, big, 10, "stream 1.5 GB");
  run<3072, 5>(buf, big, 10, "stream 1.5 GB");
  run<6144, 3>(buf, big, 10, "stream 1.5 GB");
  run<12288, 1>(buf, big, 10, "stream 1.5 GB");
  run<12288, 1>(buf, big, 7, "stream 1.5 GB");
  run<12288, 2>(buf, big, 7, "stream 1.5 GB");
  // one layer's worth of K/V at 64 rows, context 128 / 256: includes the launch + first-byte latency (~3 us of the time)
  run<3072, 6>(buf, (size_t)50 << 20, 7, "burst 50 MB");
  run<3072, 6>(buf, (size_t)50 << 20, 10, "burst 50 MB");
  run<3072, 6>(buf, (size_t)100 << 20, 7, "burst 100 MB");
  run<12288, 1>(buf, (size_t)100 << 20, 10, "burst 100 MB");
  // L2-resident operand re-read by every SM (what the activation matrix of a GEMM phase is)
  run_l2<8192, 2>(buf, (size_t)196608 * 10, 40, 10, "L2-resident, all SMs same");
  run_l2<8192, 2>(buf, (size_t)16 << 20, 4, 10, "L2-resident, all SMs same");
  run_l2<3072, 6>(buf, (size_t)16 << 20, 4, 10, "L2-resident, all SMs same");
  run_l2<12288, 1>(buf, (size_t)16 << 20, 4, 10, "L2-resident, all SMs same");
  return 0;
}
