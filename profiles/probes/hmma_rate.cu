// probe: issue rate of legacy mma.sync.m16n8k16 (bf16, fp32 accumulate) on sm_100a, per SM, for 4..16 warps
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* cyc) {
  float c[8][4] = {};
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 12, 16, 32}) {
    const int iters = 2000;
    k<<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double mma = (double)iters * 8 * warps;
    printf("%2d warps/SM: %.2f cycles per mma.m16n8k16 per SM -> %.0f MAC/clk/SM (dense bf16; tcgen05 peak is ~4000 at 2.25 PF)\n", warps, h / mma, mma * 2048 / h);
  }
  return 0;
}
