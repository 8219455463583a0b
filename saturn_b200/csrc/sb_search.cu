// sb_search.cu — the search that replaces `prob.solve(solver)` (saturn/solver/milp.py:321-327).
//
// The reference minimises makespan by branch-and-bound over the MILP of milp.py:96-319 with a
// wall-clock limit and a warm start (milp.py:103-104,151-155,197-202,323-325).  Here a population
// of candidates (one per "chain") lives in HBM and each round does, entirely on the device:
//   propose  : copy the chain's current candidate and apply one move (swap two priorities /
//              re-insert a job elsewhere in the order / change one job's option);
//   evaluate : the k_eval_tiles kernel of sb_eval.cu over all proposals (the measured hot path),
//              folding (makespan, global chain id) into a 64-bit arg-min key;
//   keep     : if the key improved, save that proposal's encoding as the incumbent;
//   accept   : Metropolis rule per chain at the round's temperature.
// Random numbers are counter-based (seed, global chain id, round), so a run is reproducible and
// independent of how chains are sharded across GPUs.
#include "sb_search.h"

namespace sb {

__device__ __forceinline__ uint32_t bounded(uint64_t r, uint32_t n) {  // uniform in [0, n)
  return static_cast<uint32_t>((static_cast<uint64_t>(static_cast<uint32_t>(r >> 32)) * n) >> 32);
}

template <int PB>
__device__ __forceinline__ int prio_ld(const uint8_t* row, int i) {
  return PB == 1 ? row[i] : reinterpret_cast<const uint16_t*>(row)[i];
}
template <int PB>
__device__ __forceinline__ void prio_st(uint8_t* row, int i, int v) {
  if (PB == 1) row[i] = static_cast<uint8_t>(v);
  else reinterpret_cast<uint16_t*>(row)[i] = static_cast<uint16_t>(v);
}

// ---- initial population: random valid options, random permutation (Fisher-Yates); thread per chain
template <int PB>
__global__ void k_init_population(SearchDev s) {
  const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= s.chains) return;
  uint8_t* orow = s.cur_o + c * s.stride_o;
  uint8_t* prow = s.cur_p + c * s.stride_p;
  const uint64_t gid = s.chain_base + static_cast<uint64_t>(c);
  for (int j = 0; j < s.J; ++j) {
    const uint64_t r = rng_u64(s.seed, gid, 0x100000000ull + j);
    uint8_t ob = s.vopt[j * kSlots + bounded(r, s.nvalid[j])];
    if (s.nodes > 1) ob = static_cast<uint8_t>((ob & 7) | (bounded(rng_u64(s.seed, gid, 0x300000000ull + j), s.nodes) << 3));
    orow[j] = ob;
    prio_st<PB>(prow, j, j);
  }
  for (int i = s.J - 1; i > 0; --i) {
    const uint64_t r = rng_u64(s.seed, gid, 0x200000000ull + i);
    const int k = bounded(r, i + 1);
    const int a = prio_ld<PB>(prow, i), b = prio_ld<PB>(prow, k);
    prio_st<PB>(prow, i, b);
    prio_st<PB>(prow, k, a);
  }
}

__device__ __forceinline__ void copy_row16(uint8_t* dst, const uint8_t* src, int bytes, int lane) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = lane; i * 16 < bytes; i += 32) d4[i] = s4[i];
}

// ---- propose: warp per chain
template <int PB>
__global__ void k_propose(SearchDev s, int round) {
  const int lane = threadIdx.x & 31;
  const long long c = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (c >= s.chains) return;
  const uint8_t* co = s.cur_o + c * s.stride_o;
  const uint8_t* cp = s.cur_p + c * s.stride_p;
  uint8_t* po = s.prop_o + c * s.stride_o;
  uint8_t* pp = s.prop_p + c * s.stride_p;
  copy_row16(po, co, static_cast<int>(s.stride_o), lane);
  copy_row16(pp, cp, static_cast<int>(s.stride_p), lane);
  __syncwarp();
  const uint64_t gid = s.chain_base + static_cast<uint64_t>(c);
  const uint64_t r0 = rng_u64(s.seed, gid, 4ull * round + 0);
  const uint64_t r1 = rng_u64(s.seed, gid, 4ull * round + 1);
  const uint64_t r2 = rng_u64(s.seed, gid, 4ull * round + 2);
  const uint32_t kind = bounded(r0, 100);
  const int J = s.J;
  if (s.nodes > 1 && kind >= 85) {
    // move one job to another node (milp.py:117-137: exactly one node per task)
    if (lane == 0) {
      const int j = bounded(r1, J);
      const uint8_t curv = co[j];
      int nn = bounded(r2, s.nodes - 1);
      if (nn >= (curv >> 3)) ++nn;
      po[j] = static_cast<uint8_t>((curv & 7) | (nn << 3));
    }
    return;
  }
  if (kind < 30) {
    // change one job's option (keeping its node)
    const int j = bounded(r1, J);
    const int n = s.nvalid[j];
    if (n > 1) {
      if (lane == 0) {
        int pick = bounded(r2, n - 1);
        const uint8_t curv = co[j];
        const uint8_t node_bits = s.nodes > 1 ? (curv & 0xf8) : 0;
        const uint8_t cur_opt = s.nodes > 1 ? (curv & 7) : curv;
        uint8_t nv = s.vopt[j * kSlots + pick];
        if (nv == cur_opt) nv = s.vopt[j * kSlots + n - 1];
        po[j] = nv | node_bits;
      }
      return;
    }
  }
  const int a = bounded(r1, J);
  int b = bounded(r2, J - 1);
  if (b >= a) ++b;
  if (J < 2) return;
  if (kind < 65) {
    // swap two priorities
    if (lane == 0) {
      const int va = prio_ld<PB>(cp, a), vb = prio_ld<PB>(cp, b);
      prio_st<PB>(pp, a, vb);
      prio_st<PB>(pp, b, va);
    }
  } else {
    // take the job at position a and re-insert it at position b
    if (a < b) {
      for (int i = a + lane; i < b; i += 32) prio_st<PB>(pp, i, prio_ld<PB>(cp, i + 1));
    } else {
      for (int i = b + 1 + lane; i <= a; i += 32) prio_st<PB>(pp, i, prio_ld<PB>(cp, i - 1));
    }
    if (lane == 0) prio_st<PB>(pp, b, prio_ld<PB>(cp, a));
  }
}

// ---- keep the incumbent encoding when the key improved (one warp)
__global__ void k_keep_best(SearchDev s, const uint8_t* rows_o, const uint8_t* rows_p) {
  const int lane = threadIdx.x & 31;
  const unsigned long long key = s.keys[0];
  if (key >= s.keys[1]) return;
  const uint64_t id = key & 0xffffffffull;
  const uint64_t base = s.chain_base & 0xffffffffull;
  const long long c = static_cast<long long>((id - base) & 0xffffffffull);
  if (c < 0 || c >= s.chains) return;
  copy_row16(s.best_o, rows_o + c * s.stride_o, static_cast<int>(s.stride_o), lane);
  copy_row16(s.best_p, rows_p + c * s.stride_p, static_cast<int>(s.stride_p), lane);
  __syncwarp();
  if (lane == 0) s.keys[1] = key;
}

// ---- Metropolis acceptance: warp per chain
__global__ void k_accept(SearchDev s, int round, float temperature) {
  const int lane = threadIdx.x & 31;
  const long long c = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (c >= s.chains) return;
  const float pm = s.prop_mk[c], cm = s.cur_mk[c];
  bool acc = pm <= cm;
  if (!acc && temperature > 0.f && isfinite(pm)) {
    const uint64_t gid = s.chain_base + static_cast<uint64_t>(c);
    const uint64_t r = rng_u64(s.seed, gid, 4ull * round + 3);
    const float u = (static_cast<uint32_t>(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    acc = u < __expf(-(pm - cm) / temperature);
  }
  if (!acc) return;
  copy_row16(s.cur_o + c * s.stride_o, s.prop_o + c * s.stride_o, static_cast<int>(s.stride_o), lane);
  copy_row16(s.cur_p + c * s.stride_p, s.prop_p + c * s.stride_p, static_cast<int>(s.stride_p), lane);
  if (lane == 0) s.cur_mk[c] = pm;
}

// ---- overwrite chains [first, first+copies) with one candidate (warp per chain)
__global__ void k_inject(SearchDev s, const uint8_t* cand_o, const uint8_t* cand_p, long long first, int copies) {
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= copies) return;
  const long long c = first + w;
  copy_row16(s.cur_o + c * s.stride_o, cand_o, static_cast<int>(s.stride_o), lane);
  copy_row16(s.cur_p + c * s.stride_p, cand_p, static_cast<int>(s.stride_p), lane);
}

// ---- tournament resampling ("go with the winners"): every chain draws a random rival and, if the
// rival's current makespan is strictly better, continues from a copy of the rival's candidate.
// Two passes through the proposal buffers (then the host swaps the buffer roles), so no chain is
// read while it is being overwritten.  Warp per chain.
template <int PB>
__global__ void k_resample(SearchDev s, int round) {
  const int lane = threadIdx.x & 31;
  const long long c = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (c >= s.chains) return;
  const uint64_t gid = s.chain_base + static_cast<uint64_t>(c);
  const uint64_t r = rng_u64(s.seed ^ 0x7e57a11ull, gid, static_cast<uint64_t>(round));
  const long long d = bounded(r, static_cast<uint32_t>(s.chains));
  const long long src = (s.cur_mk[d] < s.cur_mk[c]) ? d : c;
  copy_row16(s.prop_o + c * s.stride_o, s.cur_o + src * s.stride_o, static_cast<int>(s.stride_o), lane);
  copy_row16(s.prop_p + c * s.stride_p, s.cur_p + src * s.stride_p, static_cast<int>(s.stride_p), lane);
  if (lane == 0) s.prop_mk[c] = s.cur_mk[src];
}

// ------------------------------------------------------------------------------------------ host
static int warp_grid(long long warps, int threads) {
  long long blocks = (warps * 32 + threads - 1) / threads;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

cudaError_t search_init_population(const SearchDev& s, cudaStream_t st) {
  const int threads = 128;
  const int grid = static_cast<int>((s.chains + threads - 1) / threads);
  if (s.pb == 1) k_init_population<1><<<grid, threads, 0, st>>>(s);
  else k_init_population<2><<<grid, threads, 0, st>>>(s);
  return cudaGetLastError();
}

cudaError_t search_propose(const SearchDev& s, int round, cudaStream_t st) {
  const int threads = 256;
  const int grid = warp_grid(s.chains, threads);
  if (s.pb == 1) k_propose<1><<<grid, threads, 0, st>>>(s, round);
  else k_propose<2><<<grid, threads, 0, st>>>(s, round);
  return cudaGetLastError();
}

cudaError_t search_keep_best(const SearchDev& s, bool from_cur, cudaStream_t st) {
  k_keep_best<<<1, 32, 0, st>>>(s, from_cur ? s.cur_o : s.prop_o, from_cur ? s.cur_p : s.prop_p);
  return cudaGetLastError();
}

cudaError_t search_accept(const SearchDev& s, int round, float temperature, cudaStream_t st) {
  const int threads = 256;
  k_accept<<<warp_grid(s.chains, threads), threads, 0, st>>>(s, round, temperature);
  return cudaGetLastError();
}

cudaError_t search_resample(const SearchDev& s, int round, cudaStream_t st) {
  const int threads = 256;
  if (s.pb == 1) k_resample<1><<<warp_grid(s.chains, threads), threads, 0, st>>>(s, round);
  else k_resample<2><<<warp_grid(s.chains, threads), threads, 0, st>>>(s, round);
  return cudaGetLastError();
}

cudaError_t search_inject(const SearchDev& s, const uint8_t* cand_o, const uint8_t* cand_p, long long first,
                          int copies, cudaStream_t st) {
  const int threads = 256;
  k_inject<<<warp_grid(copies, threads), threads, 0, st>>>(s, cand_o, cand_p, first, copies);
  return cudaGetLastError();
}

}  // namespace sb
