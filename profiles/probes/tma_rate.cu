// Delivery rate of an L2-resident operand into the shared memory of all 148 SMs, as a GEMM phase of the fused decode step
// reads its activation matrix: one thread per CTA streams the same [64 x 1536] bf16 matrix PASSES times through a ring of
// STAGES stages, either as tensor-map copies (64 x 64 boxes, 128B swizzle; NBLK K blocks per copy) or as plain 1-D bulk
// copies of the same number of bytes.  Reports bytes delivered per second (all SMs) and per SM per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_parity(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}

constexpr int ROWS = 64, K = 1536, KB = K / 64;  // 24 K blocks of [64 rows x 128 B]

// MODE 0: tensor map, NBLK blocks per copy; MODE 1: 1-D bulk copy of NBLK * 8 KB from a pre-tiled image
template <int MODE, int NBLK, int STAGES>
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* tiled, int passes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * NBLK * 8192);
  constexpr uint32_t BYTES = NBLK * 8192;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int total = passes * (KB / NBLK);
    int issued = 0;
    unsigned acc = 0;
    for (int c = 0; c < total; ++c) {
      while (issued < total && issued < c + STAGES) {
        const int s = issued % STAGES, kb = ((issued + blockIdx.x) % (KB / NBLK)) * NBLK;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(BYTES) : "memory");
        if (MODE == 0)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                           smem_u32(smem + s * BYTES)),
                       "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(0), "r"(kb)
                       : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem + s * BYTES)),
                       "l"(tiled + (size_t)kb * 8192), "r"(BYTES), "r"(smem_u32(&bars[s]))
                       : "memory");
        ++issued;
      }
      wait_parity(&bars[c % STAGES], (uint32_t)(c / STAGES) & 1u);
      acc += smem[(c % STAGES) * BYTES];
    }
    if (acc == 0xffffffffu) printf("never\n");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int NBLK, int STAGES>
static void run(const CUtensorMap& map, const uint8_t* buf, const char* what) {
  const int smem = STAGES * NBLK * 8192 + 1024 + 256;
  cudaFuncSetAttribute(stream_kernel<MODE, NBLK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int passes = 40;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  stream_kernel<MODE, NBLK, STAGES><<<148, 128, smem>>>(map, buf, passes);
  cudaEventRecord(e0);
  stream_kernel<MODE, NBLK, STAGES><<<148, 128, smem>>>(map, buf, passes);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 148.0 * passes * ROWS * K * 2;
  cudaError_t e = cudaGetLastError();
  printf("%-34s %d blocks/copy, %d stages (%3d KB in flight): %7.1f us  %6.2f TB/s  %5.1f GB/s per SM%s\n", what, NBLK, STAGES,
         STAGES * NBLK * 8, ms * 1e3, bytes / (ms * 1e-3) / 1e12, bytes / 148 / (ms * 1e-3) / 1e9,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  uint8_t* buf;
  cudaMalloc(&buf, (size_t)ROWS * K * 2);
  cudaMemset(buf, 1, (size_t)ROWS * K * 2);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  auto make = [&](int nblk) {
    CUtensorMap m;
    cuuint64_t dims[3] = {64, ROWS, KB};
    cuuint64_t strides[2] = {K * 2, 128};
    cuuint32_t box[3] = {64, ROWS, (cuuint32_t)nblk};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
  };
  const CUtensorMap m1 = make(1), m4 = make(4);
  run<0, 1, 12>(m1, buf, "tensor map, 64x64 boxes");
  run<0, 1, 24>(m1, buf, "tensor map, 64x64 boxes");
  run<0, 4, 3>(m4, buf, "tensor map, 4 boxes per copy");
  run<0, 4, 6>(m4, buf, "tensor map, 4 boxes per copy");
  run<1, 1, 12>(m1, buf, "1-D bulk, 8 KB");
  run<1, 1, 24>(m1, buf, "1-D bulk, 8 KB");
  run<1, 4, 3>(m4, buf, "1-D bulk, 32 KB");
  run<1, 4, 6>(m4, buf, "1-D bulk, 32 KB");
  return 0;
}
