// group.cu — device-side hand-shake of the process-per-GPU frame ring (api.cu: rtb_group_*).
//
// The ranks of one job render row bands of the same frame and store them straight into a frame buffer of rank 0 (the resolve
// kernel's stores travel over NVLink through a CUDA-IPC mapping).  What is left to coordinate is ordering:
//   stored[rank][buffer] = k + 1   "this rank's bands of frame k are in the buffer"      (written by every rank, read by rank 0)
//   read_done[buffer]    = k + 1   "rank 0 has copied frame k out of the buffer"          (written by rank 0, read by every rank)
// Both live in rank 0's allocation behind the frames.  A rank posts a flag with a one-thread kernel ordered behind its resolve
// kernels, and waits with a one-warp kernel ordered in front of the work that depends on it, so no host thread, no barrier and
// no collective sits between frames: every process only enqueues.  Waits poll with volatile loads (over NVLink for the peers),
// back off with nanosleep, and give up after `timeout_ns` — a lost peer then surfaces as an error instead of a hung GPU.
#include "kernels.hpp"

namespace rtb {
namespace {

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void k_group_post(volatile uint32_t* flag, uint32_t value) {
  __threadfence_system();  // the resolve kernels before this launch have completed; make their stores visible system-wide first
  *flag = value;
  __threadfence_system();
}

// Lane i waits until flags[i * stride] >= target (sequence numbers: compared as a signed difference).
__global__ void k_group_wait(const volatile uint32_t* flags, int n, int stride, uint32_t target, uint32_t* error, unsigned long long timeout_ns) {
  const int i = (int)threadIdx.x;
  if (i >= n) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned backoff = 64;
  while ((int32_t)(flags[(size_t)i * (size_t)stride] - target) < 0) {
    __nanosleep(backoff);
    if (backoff < 2048) backoff *= 2;
    if (global_timer_ns() - t0 > timeout_ns) { atomicExch(error, 1u); return; }
  }
}

}  // namespace

void launch_group_post(uint32_t* flag, uint32_t value, cudaStream_t st) { k_group_post<<<1, 1, 0, st>>>(flag, value); }
void launch_group_wait(const uint32_t* flags, int n, int stride, uint32_t target, uint32_t* error, unsigned long long timeout_ns, cudaStream_t st) {
  k_group_wait<<<1, 32, 0, st>>>(flags, n, stride, target, error, timeout_ns);
}

}  // namespace rtb
