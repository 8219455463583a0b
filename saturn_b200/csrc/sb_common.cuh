// sb_common.cuh — shared device helpers for the SPASE candidate evaluator (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/saturn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "saturn_b200 kernels target sm_100a (B200) only"
#endif

namespace sb {

constexpr int kSlots = SB_NSLOT;
constexpr int kWarp = 32;
constexpr int kMaxNodes = SB_MAX_NODES;
constexpr float kSentinel = 1.0e6f;  // reference's "unprofiled" runtime, PerformanceEvaluator.py:99

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

// ---------------------------------------------------------------- mbarrier + TMA bulk copy (1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// thread-block cluster helpers (k_search_pos with the table split over a CTA pair)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// every thread of every CTA of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster window address of `addr` (a shared::cta address of this CTA) in the CTA with rank `rank`
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the inits visible to the async proxy (the TMA unit) before any bulk copy signals them
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a barrier that never completes is a bug (wrong byte count / misaligned copy);
// trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// global -> shared bulk copy through the TMA unit; completion is signalled on `bar` as `bytes`
// of transaction count.  dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------- the list-scheduling step
// State: the 8 slot ready-times kept SORTED ascending in registers (f[0] <= ... <= f[7]).  Which
// physical slot holds which time does not influence any start time or the makespan (ties are
// between equal values), so the hot kernel evolves the sorted multiset only; the slot-exact
// variant lives in k_eval_full.
//
// A job with k = km1 + 1 GPUs starts at s = f[km1] (the k-th smallest), and the k smallest
// entries are replaced by v = s + hold.  With sh[i] = f[i + k] (+inf past the end) the new sorted
// state is  f'[i] = max(f[i], min(v, sh[i]))  — every surviving element below v moves down k
// places, the k copies of v follow, larger elements stay.  sh (and s, as element -1 of the same
// window) is produced by a 3-stage barrel shifter on the bits of km1: no dynamic register
// indexing, no divergence.
//
// Instruction mix, shaped by the ncu captures in profiles/r01_summary.md — the kernel is bound by
// instruction issue, and before that by the ALU pipe, so work is spread over the pipes at the
// lowest instruction count found (54 per step):
//   stage "by 4", lower half : 4 FSEL (ALU); predicates come from the bit inside the asm so that
//                              ptxas derives all three stage predicates with ONE R2P of the opt byte
//   stage "by 4", upper half : x = f + m with m = bit ? +inf : -0.0  (1 FSEL + 4 FADD, FMA pipe;
//                              f + -0.0 == f exactly, f + inf == inf) — doubles as the copy that
//                              keeps f intact for the final max
//   stages "by 2", "by 1"    : in-place predicated moves written as `@p mad.lo dst, src, one, 0`
//                              with `one` a run-time 1, so ptxas cannot fold them into SEL: they
//                              issue as predicated IMAD on the FMA pipe; +inf padding as predicated
//                              `add dst, dst, +inf`.  Ascending order reads only unmodified sources.
//   merge                    : 7 FMNMX (min) + 8 FMNMX (max) on the ALU pipe.
__device__ __forceinline__ void pmov_fma(float& dst, float src, int bit, int one) {
  int d = __float_as_int(dst);
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p mad.lo.s32 %0, %1, %3, 0;\n\t}"
      : "+r"(d)
      : "r"(__float_as_int(src)), "r"(bit), "r"(one));
  dst = __int_as_float(d);
}
// dst = +inf under the predicate, as `@p add.f32 dst, dst, +inf`: it depends on dst, so ptxas cannot
// hoist it into a SEL of a loop-invariant.
__device__ __forceinline__ void pinf_fma(float& dst, int bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p add.f32 %0, %0, 0f7F800000;\n\t}" : "+f"(dst) : "r"(bit));
}
// bit ? a : b with the predicate formed inside the asm (FSEL)
__device__ __forceinline__ float psel(float a, float b, int bit) {
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"(bit));
  return r;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // one FMNMX3
  return d;
}

// kTrackMk: fold this job's completion (s + rt) into mk.  Needed with integer starts (the slot
// state holds s + ceil(rt), not the completion) and with several nodes (no single f[7] at the end).
// `ph` pairs the completions of two consecutive steps into one 3-input max: 0 parks this step's
// completion in `pend`, 1 folds max(mk, pend, completion); callers with unrolled loops pass t & 1 (a
// compile-time constant after unrolling), others pass -1 for the plain 2-input max.  A parked value
// that is never folded is picked up by the final max(mk, pend) (LaneState::result).
template <bool kIntegerStarts, bool kTrackMk = kIntegerStarts>
__device__ __forceinline__ void ls_step(float (&f)[8], float& mk, float& pend, float rt, int km1, int one, int ph) {
  const float INF = inf_f();
  const int b2 = km1 & 4, b1 = km1 & 2, b0 = km1 & 1;
  // stage "shift by 4"
  float x0 = psel(f[4], f[0], b2), x1 = psel(f[5], f[1], b2), x2 = psel(f[6], f[2], b2), x3 = psel(f[7], f[3], b2);
  const float m2 = psel(INF, -0.0f, b2);
  float x4 = f[4] + m2, x5 = f[5] + m2, x6 = f[6] + m2, x7 = f[7] + m2;
  // stage "shift by 2"
  pmov_fma(x0, x2, b1, one); pmov_fma(x1, x3, b1, one); pmov_fma(x2, x4, b1, one); pmov_fma(x3, x5, b1, one);
  pmov_fma(x4, x6, b1, one); pmov_fma(x5, x7, b1, one); pinf_fma(x6, b1); pinf_fma(x7, b1);
  // stage "shift by 1"
  pmov_fma(x0, x1, b0, one); pmov_fma(x1, x2, b0, one); pmov_fma(x2, x3, b0, one); pmov_fma(x3, x4, b0, one);
  pmov_fma(x4, x5, b0, one); pmov_fma(x5, x6, b0, one); pmov_fma(x6, x7, b0, one); pinf_fma(x7, b0);
  const float s = x0;  // = f[km1]
  float v;
  if (kIntegerStarts) {
    // every entry of f is an integer here, so s is; the slot is usable again at s + ceil(rt)
    v = s + ceilf(rt);
  } else {
    v = s + rt;
  }
  if (kTrackMk) {
    const float e = kIntegerStarts ? s + rt : v;
    if (ph < 0) mk = fmaxf(mk, e);
    else if (ph == 0) pend = e;
    else mk = fmax3(mk, pend, e);
  }
  f[0] = fmaxf(f[0], fminf(v, x1));
  f[1] = fmaxf(f[1], fminf(v, x2));
  f[2] = fmaxf(f[2], fminf(v, x3));
  f[3] = fmaxf(f[3], fminf(v, x4));
  f[4] = fmaxf(f[4], fminf(v, x5));
  f[5] = fmaxf(f[5], fminf(v, x6));
  f[6] = fmaxf(f[6], fminf(v, x7));
  f[7] = fmaxf(f[7], v);
}

__device__ __forceinline__ unsigned long long pack_key(float mk, uint32_t id) {
  return (static_cast<unsigned long long>(__float_as_uint(mk)) << 32) | id;
}

// counter-based RNG: one 64-bit mix per draw, keyed by (seed, stream id, counter)
__device__ __host__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __host__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t stream, uint64_t ctr) {
  return mix64(mix64(seed ^ (stream * 0xd1342543de82ef95ull)) + ctr * 0x2545f4914f6cdd1dull);
}

}  // namespace sb
