// CFG combine -> temperature -> softmax -> top-k / top-p filter -> Philox inverse-CDF draw (or argmax)
// -> delay-pattern mask-fix -> prompt-preserving write-back.  One warp per (clip, codebook) row of
// V = 1024 logits; every selection is done with warp shuffles / votes on registers.
//
// Reference code replaced (paths relative to /root/reference):
//   CFG combine            models/vaura_model.py:810-813   (uncond + (cond - uncond) * scale)
//   sampling switch        models/vaura_model.py:816-825
//   sample_top_k           utils/utils.py:163-177  (keep probs >= k-th largest, ties kept, renormalise)
//   sample_top_p           utils/utils.py:180-196  (keep sorted prefix while cumsum - p_i <= top_p)
//   multinomial            utils/utils.py:139-160  (torch.multinomial; ours: Philox4x32-10 + inverse CDF in
//                                                   vocabulary order, statistically equivalent)
//   mask-fix / write-back  models/vaura_model.py:536-544
#include "sampling.cuh"

namespace vaura {

__global__ void __launch_bounds__(128) sample_kernel(SampleArgs a) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);  // (b, k)
  const int offset = a.state ? a.state->offset : a.offset;
  if (row < a.B * a.K) sample_row(a, row / a.K, row % a.K, lane, offset);
  // last CTA to finish advances the device-resident loop state
  if (a.state) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int prev = atomicAdd(&a.state->done, 1);
      if (prev == (int)gridDim.x - 1) {
        a.state->done = 0;
        a.state->offset = offset + 1;
        __threadfence();
      }
    }
  }
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st) {
  if (a.V != 32 * kEPT) return cudaErrorInvalidValue;
  const int rows = a.B * a.K;
  sample_kernel<<<(rows + 3) / 4, 128, 0, st>>>(a);
  return cudaGetLastError();
}

__global__ void set_state_kernel(StepState* s, int offset) {
  s->offset = offset;
  s->done = 0;
  s->epoch = 0;
  s->barrier = 0;
  s->tiles_done = 0;
}

cudaError_t launch_set_state(StepState* s, int offset, cudaStream_t st) {
  set_state_kernel<<<1, 1, 0, st>>>(s, offset);
  return cudaGetLastError();
}

}  // namespace vaura
