// sb_internal.h — host-side structures shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sb_common.cuh"

namespace sb {

struct Device {
  int ordinal = 0;
  int sm_count = 0;
  size_t smem_optin = 0;  // max dynamic shared memory per CTA (227 KB on B200)
};

// peer-memory MIN exchange (sb_xchg.cu)
constexpr int kMaxRanks = 16;
struct XchgDev {
  int rank = 0, world = 1;
  unsigned long long* local = nullptr;            // our mailbox: [2 parities][2] u64 = {key, round}
  unsigned long long* peer[kMaxRanks] = {nullptr};  // every rank's mailbox mapped here (peer[rank] == local)
};
struct XchgPost {
  XchgDev x;
  unsigned long long seq = 0;
  unsigned* counter = nullptr;  // CTA completion counter (zero between launches); null = no fused post
  int fold_prev = 0;            // prologue: fold the peers' keys of round seq-1 into best_key (pipelined exchange)
  int* error = nullptr;         // set to 1 if a peer's round never shows up
};
cudaError_t xchg_post_launch(const XchgDev& x, const unsigned long long* key, unsigned long long seq, cudaStream_t st);
cudaError_t xchg_reduce_launch(const XchgDev& x, unsigned long long seq, unsigned long long* out,
                               unsigned long long* fold, int* error, cudaStream_t st);
cudaError_t xchg_post_reduce_launch(const XchgDev& x, const unsigned long long* key, unsigned long long seq,
                                    unsigned long long* out /*[2]: key, error*/, cudaStream_t st);

struct TilePlan {
  int warps = 0;
  int row_o = 0, row_p = 0, copy_o = 0, copy_p = 0;
  size_t smem = 0;
};

struct EvalCall {
  const float* tab = nullptr;  // canonical table actually used (full or reduced)
  int J = 0, SG = 0;
  const uint8_t* opt = nullptr;
  const uint8_t* prio = nullptr;
  long long B = 0;
  long long stride_o = 0, stride_p = 0;  // bytes
  unsigned flags = 0;
  int nodes = 1;  // > 1: multi-node (reduced table, opt = (node << 3) | (k - 1))
  float* out = nullptr;
  unsigned long long* best_key = nullptr;
  uint32_t id_base = 0;
  int force_generic = 0;
  XchgPost xp;  // fused post of best_key at the end of the tile kernel (tile paths only)
};

// what the fused search round needs besides an EvalCall (see k_eval_tiles<..., SEARCH = true>)
constexpr int kMaxFusedRounds = 16;
constexpr int kSnapPos = 32;  // schedule positions per window / between state snapshots (incremental search rounds)
struct SearchFuse {
  float* cur_mk = nullptr;        // [chains] makespan of each chain's current candidate
  uint8_t* cur_o = nullptr;       // writable views of the rows the EvalCall reads
  uint8_t* cur_p = nullptr;
  const uint8_t* vopt = nullptr;  // [J][8] proposable opt bytes
  const int* nvalid = nullptr;    // [J]
  uint64_t seed = 0, chain_base = 0;
  int round = 0;    // first round of this launch (RNG counters are keyed by the round number)
  int nrounds = 1;  // rounds run back to back inside one launch, the rows staying on chip (<= kMaxFusedRounds)
  int nodes = 1;
  float temperature[kMaxFusedRounds] = {};  // per round of this launch
  // Tournament resampling inside the tile kernel: before every round r with (r - 1) % resample_every == 0
  // (r > 1) each lane takes over the rows of a random lane of its warp if that lane's candidate is better.
  // `deal` changes which chains share a warp: 0 = chain = tile * 32 + lane, 1 = chain = lane * ntiles + tile.
  int resample_every = 0;
  int deal = 0;
  // Incremental re-evaluation (snap != nullptr; one node only).  Every round a WARP draws one window of
  // kSnapPos schedule positions and its 32 chains make their moves inside that window (windowed moves:
  // `win` = 1), so that all of them can resume the list schedule from the same place: the sorted slot
  // state + running makespan of a chain's CURRENT candidate is snapshotted at every window boundary
  // ([tile][boundary][2 buffers][9 words][32 lanes] floats in HBM/L2, coalesced, valid for one launch), a
  // proposal is scored from the snapshot in front of its window and writes its own boundary states into
  // the other buffer; accepting the move flips which buffer is current (a per-lane bit per boundary), so a
  // rejected move costs nothing to undo.  An unmodified pass at the start of the launch fills the buffers.
  int win = 0;                    // 1: moves are drawn inside a per-warp window (also without snapshots)
  int win_bias = 0;               // 0: windows uniform; 1: P(w) ~ w + 1 (draw_window)
  float* snap = nullptr;          // scratch, (ntiles * nbound * 2 * 9 * 32) floats; nullptr = score every proposal from position 0
  unsigned long long* verify_bad = nullptr;  // test hook: also score from position 0 and count differing results here
  // keep-best in the kernel's tail (KeepBest::counter != nullptr): the CTA that finishes last copies the
  // incumbent's rows when the population's best key improved — saves the separate one-warp launch per round
  struct KeepBest {
    unsigned* counter = nullptr;         // zero between launches
    unsigned long long* keys = nullptr;  // [0] best key of the population, [1] key of the saved encoding
    uint8_t *best_o = nullptr, *best_p = nullptr;
    long long chains = 0, stride_o = 0, stride_p = 0;
  } keep;
};

int plan_tiles(const Device& dev, int J, int SG, int pb, bool stream, int nodes, TilePlan* tp, bool tab_global = false);
cudaError_t eval_launch(const Device& dev, const EvalCall& c, cudaStream_t st, int* path_used);
cudaError_t eval_alt_launch(const Device& dev, const EvalCall& c, cudaStream_t st);  // sb_eval_alt.cu
int search_round_mode(const Device& dev, int J, int SG, int nodes);
cudaError_t search_round_launch(const Device& dev, const EvalCall& c, const SearchFuse& sf, cudaStream_t st);
cudaError_t eval_full_launch(const Device& dev, const EvalCall& c, float* start, uint32_t* slotmask, cudaStream_t st);
cudaError_t validate_launch(const Device& dev, const EvalCall& c, unsigned long long* bad, cudaStream_t st,
                            bool by_pos = false);

// table construction (sb_table.cu)
cudaError_t build_table_launch(const float* T, int J, int S, int G, uint64_t gcount_packed, float* tab, float* tmin,
                               uint8_t* args, cudaStream_t st);
// valid (non-dominated, non-sentinel) option lists for the search: vopt[J][8] opt bytes, nvalid[J]
cudaError_t build_valid_launch(const float* tmin, const uint8_t* args, int J, int reduced, float sentinel, uint8_t* vopt,
                               int* nvalid, cudaStream_t st);

}  // namespace sb
