// sb_xchg.cu — the per-round MIN exchange over NVLink peer memory.
//
// The only communication the path has is one MIN of a packed 64-bit key per round:
// (fp32 makespan bits << 32) | global candidate id.  Through NCCL that is a latency-bound 8-byte
// all-reduce (~15-30 us per round next to a 0.45 ms evaluation kernel).  Here every rank owns a
// small mailbox in its HBM, mapped into every peer's address space with CUDA IPC.  At the end of a
// round a rank publishes {key, round number} in ITS OWN mailbox (two local stores, release at
// system scope — fused into the tail of the evaluation kernel, which therefore pays no NVLink
// latency); a one-warp kernel on every rank then has lane r LOAD rank r's mailbox straight over
// NVLink/NVSwitch (acquire at system scope) until the round number matches, and folds the MIN with
// warp shuffles.  No NCCL kernel, no host round trip, one NVLink round trip of latency.
//
// Two-deep (parity) slots suffice: a rank publishes round r+2 only after its fold of round r+1,
// which needs every peer's round r+1, which each peer publishes after ITS fold of round r (stream
// order) — i.e. after it finished reading our round-r slot.
#include "sb_internal.h"

namespace sb {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// publish this rank's key for round `seq` in its own mailbox: slot[parity] = {key, seq}
__global__ void k_xchg_post(XchgDev x, const unsigned long long* key, unsigned long long seq) {
  if (threadIdx.x != 0) return;
  unsigned long long* slot = x.local + (seq & 1ull) * 2;
  st_relaxed_sys(slot, *key);
  st_release_sys(slot + 1, seq);  // the key is visible before the round number
}

// lane r reads rank r's mailbox over NVLink until it shows round `seq`; the warp folds the MIN into
// *out (and into *fold, typically the local best key, so the running best becomes the global one).
__global__ void k_xchg_reduce(XchgDev x, unsigned long long seq, unsigned long long* out, unsigned long long* fold,
                              int* error) {
  const int r = threadIdx.x;
  unsigned long long k = ~0ull;
  bool ok = true;
  if (r < x.world) {
    const unsigned long long* slot = x.peer[r] + (seq & 1ull) * 2;
    unsigned spins = 0;
    while (ld_acquire_sys(slot + 1) < seq) {
      if (++spins > (1u << 22)) {  // seconds (each look is an NVLink round trip): a peer died or never posted
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    if (ok) k = ld_relaxed_sys(slot);
  }
  const unsigned all_ok = __all_sync(0xffffffffu, ok);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, d);
    k = o < k ? o : k;
  }
  if (threadIdx.x == 0) {
    if (!all_ok) {
      *error = 1;
      *out = ~0ull;  // never leave a stale key behind a timed-out wait
    } else {
      *out = k;
      if (fold != nullptr && k < *fold) *fold = k;
    }
  }
}

// post + fold in one launch (the in-process multi-device search, sb_api.cu): thread 0 publishes this device's
// key, then lane r waits for device r's mailbox to show round `seq`; out[0] = MIN over all devices (~0 on a
// timed-out wait), out[1] = 1 if a wait timed out.
__global__ void k_xchg_post_reduce(XchgDev x, const unsigned long long* key, unsigned long long seq,
                                   unsigned long long* out) {
  const int r = threadIdx.x;
  if (r == 0) {
    unsigned long long* slot = x.local + (seq & 1ull) * 2;
    st_relaxed_sys(slot, *key);
    st_release_sys(slot + 1, seq);
  }
  __syncwarp();
  unsigned long long k = ~0ull;
  bool ok = true;
  if (r < x.world) {
    const unsigned long long* slot = x.peer[r] + (seq & 1ull) * 2;
    unsigned spins = 0;
    while (ld_acquire_sys(slot + 1) < seq) {
      if (++spins > (1u << 22)) {
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    if (ok) k = ld_relaxed_sys(slot);
  }
  const unsigned all_ok = __all_sync(0xffffffffu, ok);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, d);
    k = o < k ? o : k;
  }
  if (threadIdx.x == 0) {
    out[0] = all_ok ? k : ~0ull;
    out[1] = all_ok ? 0ull : 1ull;
  }
}

cudaError_t xchg_post_reduce_launch(const XchgDev& x, const unsigned long long* key, unsigned long long seq,
                                    unsigned long long* out, cudaStream_t st) {
  k_xchg_post_reduce<<<1, 32, 0, st>>>(x, key, seq, out);
  return cudaGetLastError();
}

cudaError_t xchg_post_launch(const XchgDev& x, const unsigned long long* key, unsigned long long seq, cudaStream_t st) {
  k_xchg_post<<<1, 32, 0, st>>>(x, key, seq);
  return cudaGetLastError();
}

cudaError_t xchg_reduce_launch(const XchgDev& x, unsigned long long seq, unsigned long long* out,
                               unsigned long long* fold, int* error, cudaStream_t st) {
  k_xchg_reduce<<<1, 32, 0, st>>>(x, seq, out, fold, error);
  return cudaGetLastError();
}

}  // namespace sb
