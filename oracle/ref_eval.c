/* CPU ORACLE (test infrastructure — NOT product code).
 *
 * Plain-C restatement of oracle/ref_eval.py:list_schedule, batched over
 * candidates with OpenMP.  It exists so that parity tests and bench.py's
 * cpu_baseline / --impl reference legs can evaluate 1e5..1e6 candidates in
 * seconds.  The product path never links or calls this file.
 *
 * Semantics restated (file:line under /root/reference):
 *   gang of k GPUs on one 8-GPU node             saturn/solver/milp.py:62, 209-227
 *   one shared Integer start per task            milp.py:139-149, 233-256
 *   no two tasks overlap on a GPU                milp.py:277-319
 *   makespan >= start + runtime                  milp.py:162-177
 *
 * Encodings (include/saturn_b200.h): tab[J][S][8] runtimes (+inf = absent),
 * opt[j] = (s << 3) | (k - 1), prio[i] = job scheduled i-th.
 * Rule: each job takes the k slots with smallest (ready, slot) — ties to the
 * lowest slot — starts at the largest ready time among them, and holds them
 * until start + ceil(rt) (integer_starts) or start + rt.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -ffp-contract=off oracle/ref_eval.c -o oracle/libref_eval.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NSLOT_MAX 8
#define NODES_MAX 8   /* multi-node: opt byte = (node << 3) | (k - 1), reduced table (S = 1) */

#define DEFINE_EVAL(NAME, REAL, CEIL)                                                         \
  int NAME(const REAL* tab, int J, int S, const uint8_t* opt, const void* prio,              \
           int prio_bytes, int64_t B, int integer_starts, int nslot, int nodes, REAL* makespan, \
           REAL* start_out, uint32_t* mask_out, int nthreads) {                              \
    if (J <= 0 || S <= 0 || nslot < 1 || nslot > NSLOT_MAX) return -1;                       \
    if (nodes < 1 || nodes > NODES_MAX || (nodes > 1 && S != 1)) return -3;                  \
    if (prio_bytes != 1 && prio_bytes != 2) return -2;                                       \
    if (nthreads > 0) {                                                                      \
      _Pragma("omp parallel for schedule(dynamic, 64) num_threads(nthreads)")                \
      for (int64_t b = 0; b < B; ++b) {                                                      \
        NAME##_one(tab, J, S, opt + (size_t)b * J, (const uint8_t*)prio +                    \
                   (size_t)b * J * prio_bytes, prio_bytes, integer_starts, nslot, nodes,     \
                   makespan + b, start_out ? start_out + (size_t)b * J : 0,                  \
                   mask_out ? mask_out + (size_t)b * J : 0);                                 \
      }                                                                                      \
    } else {                                                                                 \
      for (int64_t b = 0; b < B; ++b)                                                        \
        NAME##_one(tab, J, S, opt + (size_t)b * J, (const uint8_t*)prio +                    \
                   (size_t)b * J * prio_bytes, prio_bytes, integer_starts, nslot, nodes,     \
                   makespan + b, start_out ? start_out + (size_t)b * J : 0,                  \
                   mask_out ? mask_out + (size_t)b * J : 0);                                 \
    }                                                                                        \
    return 0;                                                                                \
  }

#define DEFINE_ONE(NAME, REAL, CEIL)                                                          \
  static void NAME##_one(const REAL* tab, int J, int S, const uint8_t* opt,                  \
                         const uint8_t* prio, int prio_bytes, int integer_starts,            \
                         int nslot, int nodes, REAL* makespan, REAL* start_out,              \
                         uint32_t* mask_out) {                                               \
    REAL ready_all[NODES_MAX * NSLOT_MAX];                                                   \
    int order[NSLOT_MAX];                                                                    \
    REAL mk = 0;                                                                             \
    int bad = 0;                                                                             \
    (void)S;                                                                                 \
    for (int g = 0; g < NODES_MAX * NSLOT_MAX; ++g) ready_all[g] = 0;                        \
    for (int i = 0; i < J; ++i) {                                                            \
      int j = prio_bytes == 1 ? prio[i] : ((const uint16_t*)prio)[i];                        \
      int o = opt[j];                                                                        \
      int k = (o & 7) + 1;                                                                   \
      int node = nodes > 1 ? (o >> 3) : 0;                                                   \
      REAL rt = tab[(size_t)j * S * 8 + (nodes > 1 ? (o & 7) : o)];                          \
      if (k > nslot || node >= nodes) { bad = 1; break; }                                    \
      REAL* ready = ready_all + node * NSLOT_MAX;                                            \
      /* order slots by (ready, slot): insertion sort, stable => ties keep slot order */     \
      for (int g = 0; g < nslot; ++g) {                                                      \
        int p = g;                                                                           \
        while (p > 0 && ready[order[p - 1]] > ready[g]) { order[p] = order[p - 1]; --p; }    \
        order[p] = g;                                                                        \
      }                                                                                      \
      REAL s = ready[order[k - 1]];                                                          \
      REAL hold = (integer_starts && isfinite(rt)) ? CEIL(rt) : rt;                          \
      REAL nxt = s + hold;                                                                   \
      uint32_t m = 0;                                                                        \
      for (int q = 0; q < k; ++q) { ready[order[q]] = nxt; m |= 1u << order[q]; }            \
      if (start_out) start_out[j] = s;                                                       \
      if (mask_out) mask_out[j] = ((uint32_t)node << 16) | m;                                \
      REAL c = s + rt;                                                                       \
      if (c > mk) mk = c;                                                                    \
    }                                                                                        \
    *makespan = bad ? (REAL)INFINITY : mk;                                                   \
  }

DEFINE_ONE(ref_eval_f32, float, ceilf)
DEFINE_EVAL(ref_eval_f32, float, ceilf)
DEFINE_ONE(ref_eval_f64, double, ceil)
DEFINE_EVAL(ref_eval_f64, double, ceil)

int ref_eval_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
