"""`saturn.solver` drop-in: solve() and convert_into_comprehensible().

Same call shapes, argument meaning and return structure as the reference
(saturn/solver/milp.py:23 and :448, re-exported by saturn/solver/__init__.py:1-2), so
`saturn.orchestrate` (orchestrator.py:21-23,55-61,69-75) and user code keep working:

    sta, tga, bss, bna, boa, makespan = solve(task_list, presolved, gurobi=..., threads=...,
                                               interval=..., timeout=...)
    node_per_task, task_dependency_dict, start_times = convert_into_comprehensible(
        task_list, bss, boa, tga, bna, sta)

Instead of building the MILP of milp.py:96-319 and shelling out to Gurobi/CBC (milp.py:321-327),
solve() uploads the (gpu_count, runtime) table the MILP would have been built from
(milp.py:77-81), searches list-schedule candidates on the GPU (saturn_b200.search) and emits the
winner in exactly the nested-list layout milp.py:330-352,445 returns.  `gurobi` and `threads` are
accepted and ignored.  `timeout` bounds the wall-clock of the search, as it bounded the MILP.

There is no CPU fallback: without the CUDA library / a GPU this raises.
"""
from __future__ import annotations

import math
import os
import time
from collections import defaultdict
from typing import List, Optional, Sequence

import numpy as np

NSLOT = 8            # GPUs per node, reference milp.py:62 (DEBUG = True -> 8 per node)
REPLAN_THRESHOLD = 500.0   # milp.py:363


class SolverError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------ inputs
def gpu_time_tuples_of(task_list) -> List[List[tuple]]:
    """[(gpu_count, runtime), ...] per task in dict-insertion order — milp.py:77-81."""
    out = []
    for task in task_list:
        out.append([(g_count, strat.runtime) for g_count, strat in task.strategies.items()])
    return out


def build_table(task_list):
    """Task.strategies -> (T[J][1][8] fp32 by gpu_count column, usable[J][8], optindex[J][8]).

    optindex[j][k-1] = position of the gpu_count-k option in task j's strategies dict (the index
    `bss` is expressed in, milp.py:96-111,477-486), -1 if the task has no such option.
    Options that cannot fit one node (gpu_count > 8, milp.py:62,209-227) are dropped; options whose
    executor is None are the profiler's sentinels (PerformanceEvaluator.py:99,106) and are kept in
    the table but never proposed unless a task has nothing else.
    """
    J = len(task_list)
    # one pass over the Python objects collecting plain lists (comprehensions: this loop is the host-side cost
    # of a solve at J = 256), the filtering and the arithmetic are vectorised below
    counts = [len(task.strategies) for task in task_list]
    if 0 in counts:
        j = counts.index(0)
        raise SolverError("task %r has no strategies; run the trial runner first" % getattr(task_list[j], "name", j))
    keys = [g for task in task_list for g in task.strategies]
    strats = [st for task in task_list for st in task.strategies.values()]
    rts = [st.runtime for st in strats]
    jj_all = np.repeat(np.arange(J), counts)
    oo_all = np.concatenate([np.arange(c) for c in counts]) if J else np.zeros(0, dtype=np.int64)
    uu_all = np.fromiter((getattr(st, "executor", True) is not None for st in strats), dtype=bool, count=len(strats))
    ok_key = np.fromiter((isinstance(g, (int, np.integer)) and 1 <= g <= NSLOT for g in keys), dtype=bool, count=len(keys))
    r_all = np.array([np.nan if r is None else r for r in rts], dtype=np.float64)
    r_all = np.where((r_all < 0) & (r_all >= -1e-6), 0.0, r_all)   # forecast's in-place decrements (executor.py:166-168) can undershoot by an ulp
    ok = ok_key & np.isfinite(r_all) & (r_all >= 0)
    jj, oo, rr, uu = jj_all[ok], oo_all[ok], r_all[ok], uu_all[ok]
    kk = np.array([int(g) - 1 for g, k in zip(keys, ok) if k], dtype=np.int64)
    T = np.full((J, 1, NSLOT), np.inf, dtype=np.float32)
    optindex = np.full((J, NSLOT), -1, dtype=np.int64)
    usable = np.zeros((J, NSLOT), dtype=bool)
    if len(jj):
        r64 = np.asarray(rr, dtype=np.float64)
        v = r64.astype(np.float32)                                   # smallest fp32 >= rt: the device's start + ceil(rt)
        low = v.astype(np.float64) < r64
        v[low] = np.nextafter(v[low], np.float32(np.inf))
        # dict keys are unique, so a (task, gpu_count) cell is written at most once; keep the first of the
        # smallest anyway (e.g. the keys 2 and numpy.int64(2) of a hand-built dict)
        order = np.lexsort((oo, v, kk, jj))
        first = np.ones(len(order), dtype=bool)
        first[1:] = (jj[order][1:] != jj[order][:-1]) | (kk[order][1:] != kk[order][:-1])
        sel = order[first]
        T[jj[sel], 0, kk[sel]] = v[sel]
        optindex[jj[sel], kk[sel]] = oo[sel]
        usable[jj[sel], kk[sel]] = uu[sel]
    none = ~np.isfinite(T[:, 0, :]).any(axis=1)
    if none.any():
        j = int(np.argmax(none))
        raise SolverError("task %r has no option that fits a node of %d GPUs" % (getattr(task_list[j], "name", j), NSLOT))
    return T, usable, optindex


# ------------------------------------------------------------------------------------------ outputs
def plan_to_arrays(n_options: Sequence[int], opt_index: Sequence[int], start: Sequence[float],
                   slotmask: Sequence[int], position: Sequence[int], nodes: int = 1,
                   node_of: Optional[Sequence[int]] = None):
    """(start, GPU mask, chosen option, schedule position) per task -> the reference's arrays.

    Layout and meaning follow milp.py:330-352: sta[N][G][J] start times (0 where the task does
    not run), tga[J][N][G] occupancy, bss[J][S_t] one-hot option, bna[J][N] one-hot node,
    boa[a][b] == 1 iff task a is ordered before task b (milp.py:292-319,510); the diagonal is
    None exactly as the reference leaves it (those variables never enter a constraint).
    All entries are plain Python floats.
    """
    J = len(n_options)
    start_a = np.asarray(start, dtype=np.float64)
    mask_a = np.asarray(slotmask, dtype=np.int64)
    node_a = np.zeros(J, dtype=np.int64) if node_of is None else np.asarray(node_of, dtype=np.int64)
    occ = ((mask_a[:, None] >> np.arange(NSLOT)[None, :]) & 1).astype(bool)        # [J][8]
    tga_a = np.zeros((J, nodes, NSLOT))
    sta_a = np.zeros((nodes, NSLOT, J))
    jj, gg = np.nonzero(occ)
    tga_a[jj, node_a[jj], gg] = 1.0
    sta_a[node_a[jj], gg, jj] = start_a[jj]
    bna_a = np.zeros((J, nodes))
    bna_a[np.arange(J), node_a] = 1.0
    sta, tga, bna = sta_a.tolist(), tga_a.tolist(), bna_a.tolist()
    bss = [[0.0] * int(n_options[t]) for t in range(J)]
    for t in range(J):
        bss[t][int(opt_index[t])] = 1.0
    pos = np.asarray(position)
    boa = (pos[:, None] < pos[None, :]).astype(np.float64).tolist()
    for t in range(J):
        boa[t][t] = None
    return sta, tga, bss, bna, boa


def candidate_from_arrays(task_list, presolved, nodes: int = 1):
    """Warm start: turn a previous plan (the `presolved` tuple) back into a candidate
    (reduced opt bytes, priority order) — the role of setInitialValue at milp.py:103-104,151-155,197-202."""
    if presolved is None:
        return None
    try:
        sta, tga, bss, bna, boa, _mk = presolved
        J = len(task_list)
        if sta is None or len(tga) != J or len(bss) != J:
            return None
        opt = np.zeros(J, dtype=np.uint8)
        starts = np.zeros(J)
        for t, task in enumerate(task_list):
            keys = list(task.strategies.keys())
            if len(bss[t]) != len(keys):
                return None
            k = int(keys[int(np.argmax(bss[t]))])
            if not 1 <= k <= NSLOT:
                return None
            n = int(np.argmax(bna[t]))
            if n >= nodes:
                return None
            opt[t] = (k - 1) | ((n << 3) if nodes > 1 else 0)
            gl = [g for g, v in enumerate(tga[t][n]) if v is not None and round(v) == 1]
            starts[t] = sta[n][gl[0]][t] if gl else 0.0
        order = np.argsort(starts, kind="stable")
        return opt, order
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ solve
_ENGINE = None
_MULTI: dict = {}
last_stats: dict = {}
FP32_EXACT_HORIZON = float(1 << 24)   # integer seconds are exact in the kernels' fp32 state below this


def _engine(devices=None):
    """The process-wide engine: one device (default), or — `devices` = N or a list of ordinals, else
    SATURN_B200_DEVICES — one handle per device of this process (engine.MultiEngine)."""
    global _ENGINE
    if devices is None:
        env = os.environ.get("SATURN_B200_DEVICES", "")
        if env:
            devices = [int(x) for x in env.split(",")] if "," in env else int(env)
    if devices is not None and not isinstance(devices, int):
        devices = tuple(int(d) for d in devices)
        if len(devices) == 1:
            devices = None if devices[0] == 0 else devices
    if devices is None or devices == 1:
        if _ENGINE is None:
            from .engine import Engine
            _ENGINE = Engine()
        return _ENGINE
    if devices not in _MULTI:
        from .engine import MultiEngine
        _MULTI[devices] = MultiEngine(devices)
    return _MULTI[devices]


def _check_horizon(T, found_makespan=None):
    """The kernels keep schedule times in fp32: `start + ceil(rt)` is exact only below 2^24 s (194 days).
    Before the search: a table whose area lower bound (sum_j min_k k * rt_jk / 8) already reaches that bound
    cannot have an exactly representable plan and is refused.  After the search (`found_makespan`, the
    device's value): every time inside the winning schedule is <= its makespan, and fp32 addition rounds
    monotonically, so a makespan below 2^24 proves that all of its starts were computed exactly; anything
    else is refused instead of returned with silently rounded starts (rescale to coarser time units)."""
    if found_makespan is not None:
        if not float(found_makespan) < FP32_EXACT_HORIZON:
            raise SolverError("best plan found has makespan %.6g s >= 2^24 s: its start times are not exact in "
                              "fp32; express runtimes in coarser units (e.g. minutes)" % float(found_makespan))
        return
    k = np.arange(1, T.shape[-1] + 1, dtype=np.float64)
    area = np.where(np.isfinite(T), T.astype(np.float64) * k, np.inf).reshape(T.shape[0], -1).min(axis=1)
    lower = float(area.sum()) / NSLOT
    if lower >= FP32_EXACT_HORIZON:
        raise SolverError("area lower bound of the makespan is %.3g s >= 2^24 s: schedule times are not exact in "
                          "fp32 at that horizon; express runtimes in coarser units (e.g. minutes) or drop sentinel "
                          "options" % lower)

def _default_nodes() -> int:
    env = os.environ.get("SATURN_B200_NODES")
    if env:
        return max(1, int(env))
    ray = __import__("sys").modules.get("ray")      # never import Ray just to ask
    try:
        if ray is not None and ray.is_initialized():
            return max(1, len(ray.nodes()))
    except Exception:
        pass
    return 1


def solve(task_list, presolved=None, gurobi=True, threads=max(1, (os.cpu_count() or 4) // 4), interval=1000,
          timeout=500, *, chains: Optional[int] = None, rounds: Optional[int] = None, seed: int = 0,
          integer_starts: bool = True, engine=None, hysteresis: Optional[bool] = None,
          nodes: Optional[int] = None, devices=None):
    """Drop-in for saturn.solver.solve (milp.py:23).

    Returns (sta, tga, bss, bna, boa, makespan) — milp.py:445 — with a real float makespan
    (the reference returns None on a cold start, milp.py:394-399; callers only thread it back in
    as `presolved`).  Keyword-only extras tune the GPU search; environment overrides:
    SATURN_B200_CHAINS, SATURN_B200_ROUNDS, SATURN_B200_BUDGET_S, SATURN_B200_HYSTERESIS.

    Devices.  `devices=N` (or a list of CUDA ordinals, or SATURN_B200_DEVICES) shards the search population
    over N GPUs of this process — one handle per device, one MIN of a uint64 per group of rounds over NVLink
    peer memory (sb_search_run_multi); `chains` is per device.  The reference calls solve() from a single
    process (orchestrator.py:21-23,55,69), so this is how that call site uses the whole node.

    Nodes.  The reference plans over len(ray.nodes()) nodes of 8 GPUs each (milp.py:58-62); here the
    node count is the `nodes` keyword, else SATURN_B200_NODES, else an initialised Ray's node count,
    else 1.  A task runs on exactly one node (milp.py:117-137).

    Re-planning policy.  As shipped, the reference ALWAYS adopts the fresh plan: its comparator
    (milp.py:383-442) keys on `saved_makespan`, which stays None from the cold start on
    (milp.py:394-399 never assigns it), so the swap / keep branches are unreachable.
    That observable behaviour is the default here.  `hysteresis=True` (or SATURN_B200_HYSTERESIS=1)
    enables the documented intent instead: keep the current plan, shifted by one interval, unless
    the new one is better by more than interval + 500 s (milp.py:363,377,429-442).
    """
    from .search import run_search
    t_wall = time.perf_counter()
    task_list = list(task_list)
    J = len(task_list)
    if J == 0:
        return [[[] for _ in range(NSLOT)]], [], [], [], [], 0.0
    eng = engine if engine is not None else _engine(devices)
    T, usable, optindex = build_table(task_list)
    # sentinel cells (executor None) must never be proposed: they are removed from the device table
    # unless the task has nothing else; every remaining finite cell is usable (sentinel = +inf)
    Tdev = T.copy()
    for j in range(J):
        if usable[j].any():
            Tdev[j, 0, ~usable[j]] = np.inf
    _check_horizon(Tdev)
    if nodes is None:
        nodes = _default_nodes()
    nodes = int(nodes)
    eng.set_table(Tdev, list(range(1, NSLOT + 1)), sentinel=float("inf"), nodes=nodes)
    if chains is None:
        chains = int(os.environ.get("SATURN_B200_CHAINS", 0))
        if chains <= 0:
            # about 131072 chains, rounded to whole waves of the round kernel (no partially filled last wave)
            wave = eng.search_wave(reduced=True) if hasattr(eng, "search_wave") else 0
            chains = max(1, round((1 << 17) / wave)) * wave if wave > 0 else 1 << 17
    if rounds is None:
        rounds = int(os.environ.get("SATURN_B200_ROUNDS", 400))
    budget = float(os.environ.get("SATURN_B200_BUDGET_S", 20.0))
    try:
        budget = min(budget, float(timeout))
    except (TypeError, ValueError):
        pass
    warm = candidate_from_arrays(task_list, presolved, nodes)
    res = run_search(eng, chains=chains, rounds=rounds, seed=seed, integer_starts=integer_starts, reduced=True,
                     time_budget_s=budget, patience=max(40, rounds // 4), warm=warm)
    _check_horizon(Tdev, res.makespan)
    dec = eng.decode(res.opt, res.prio, integer_starts=integer_starts, reduced=True)
    gpus = dec["gpus"].astype(np.int64)
    chosen = optindex[np.arange(J), gpus - 1]
    if (chosen < 0).any():
        raise SolverError("search returned an option a task does not have")
    position = np.empty(J, dtype=np.int64)
    position[res.prio.astype(np.int64)] = np.arange(J)
    n_options = [len(t.strategies) for t in task_list]
    prop = plan_to_arrays(n_options, chosen, dec["start"], dec["slotmask"], position, nodes=nodes,
                          node_of=dec["node"])
    # the makespan the caller sees is recomputed in float64 from the emitted plan and the tasks'
    # own (un-rounded) runtimes: max_t start_t + runtime_t  (milp.py:170-177)
    rts = [list(t.strategies.values())[int(chosen[i])].runtime for i, t in enumerate(task_list)]
    prop_makespan = max(float(dec["start"][i]) + float(rts[i]) for i in range(J))

    global last_stats
    last_stats = {"candidates": res.evaluated, "rounds": res.rounds, "search_wall_s": res.wall_s,
                  "device_makespan": res.makespan, "makespan": prop_makespan, "J": J, "chains": chains,
                  "nodes": nodes, "devices": len(getattr(eng, "engines", [eng])),
                  "total_wall_s": None, "adopted": True}

    # ---- introspection hysteresis (opt-in): the documented intent of milp.py:363-442
    out = prop + (prop_makespan,)
    if hysteresis is None:
        hysteresis = os.environ.get("SATURN_B200_HYSTERESIS", "0") not in ("", "0", "false", "False")
    if presolved is not None and hysteresis:
        p_sta, p_tga, p_bss, p_bna, p_boa, saved = presolved
        same_tasks = p_tga is not None and len(p_tga) == J
        if saved is not None and same_tasks:
            try:
                itv = float(interval)
            except (TypeError, ValueError):
                itv = 1000.0
            if not (prop_makespan < float(saved) - itv - REPLAN_THRESHOLD):
                # keep the current plan, shifted by one interval (milp.py:429-442)
                kept_sta = [[[max(float(v) - itv, 0.0) for v in g] for g in n] for n in p_sta]
                out = (kept_sta, p_tga, p_bss, p_bna, p_boa, float(saved) - itv)
                last_stats["adopted"] = False
    last_stats["total_wall_s"] = time.perf_counter() - t_wall
    return out


# ------------------------------------------------------------------------------------------ dense T
NOT_PROFILED = 1.0e6   # PerformanceEvaluator.py:99  (gpu count outside the task's gpu_range)
FAILED = 1.0e8         # PerformanceEvaluator.py:106 (every executor failed at this gpu count)


def table_from_trials(n_tasks: int, n_executors: int, gpu_ranges, flat_results, max_gpus: int = NSLOT):
    """The trial runner's raw results as the dense tensor the GPU path ingests.

    `flat_results` is the list PerformanceEvaluator.search collects (PerformanceEvaluator.py:78-93): one
    `(params, runtime)` per (task, g in the task's gpu_range, executor) in that nesting order, runtime
    already scaled to the whole job (`:24-26`), `params is None` for a failed trial.  `gpu_ranges[t]` is
    the task's gpu_range (None = 1..max_gpus).  Returns
        T[J][S][G] fp32  runtime of task j under executor s on g+1 GPUs; the reference's sentinels where
                         there is no measurement: 1e6 not profiled (`:99`), 1e8 failed (`:106`)
        mask[J][S][G]    True where a trial succeeded
        params[J][S][G]  the executor's tuned parameters (object array, None elsewhere)
    Instead of collapsing the executor axis into task.strategies on the host (`:101-115`) the solver is
    given all of it: `solve_table(T, mask)`; `strategies_from_table` gives the dict view."""
    G = int(max_gpus)
    T = np.full((n_tasks, n_executors, G), NOT_PROFILED, dtype=np.float32)
    mask = np.zeros((n_tasks, n_executors, G), dtype=bool)
    params = np.empty((n_tasks, n_executors, G), dtype=object)
    it = iter(flat_results)
    for t in range(n_tasks):
        rng_t = gpu_ranges[t] if gpu_ranges is not None and gpu_ranges[t] is not None else range(1, G + 1)
        for g in rng_t:
            for e in range(n_executors):
                prm, runtime = next(it)
                if not 1 <= int(g) <= G:
                    continue
                if prm is not None and runtime is not None:
                    v = np.float32(runtime)
                    if float(v) < float(runtime):                     # round up, as build_table does
                        v = np.nextafter(v, np.float32(np.inf))
                    T[t, e, int(g) - 1] = v
                    mask[t, e, int(g) - 1] = True
                    params[t, e, int(g) - 1] = prm
                else:
                    T[t, e, int(g) - 1] = FAILED
    return T, mask, params


def strategies_from_table(T, mask, executors=None, params=None, gcount=None):
    """The compatibility view: what PerformanceEvaluator.py:96-115 would have attached to each task —
    per task an ordered dict {g: Strategy(executor, g, params, runtime)} over ALL gpu counts, the fastest
    executor per g (first minimum wins, strict `<` at `:110`), Strategy(None, g, None, 1e6) where nothing was
    profiled and Strategy(None, g, None, 1e8) where every executor failed."""
    from .representations import Strategy
    T = np.asarray(T)
    mask = np.asarray(mask, dtype=bool)
    J, S, G = T.shape
    gcount = list(range(1, G + 1)) if gcount is None else [int(g) for g in gcount]
    out = []
    for j in range(J):
        d = {}
        for gi, g in enumerate(gcount):
            ok = mask[j, :, gi]
            if ok.any():
                col = np.where(ok, T[j, :, gi], np.inf)
                e = int(np.argmin(col))                                   # first minimum
                ex = executors[e] if executors is not None else e
                d[g] = Strategy(ex, g, params[j, e, gi] if params is not None else None, float(T[j, e, gi]))
            else:
                failed = bool((T[j, :, gi] >= FAILED).any())
                d[g] = Strategy(None, g, None, FAILED if failed else NOT_PROFILED)
        out.append(d)
    return out


def solve_table(T, mask=None, gcount=None, presolved=None, interval=1000, timeout=500, *,
                chains: Optional[int] = None, rounds: Optional[int] = None, seed: int = 0,
                integer_starts: bool = True, engine=None, nodes: Optional[int] = None, devices=None):
    """solve() on the dense profiler tensor T[J][S][G] (+ mask of usable cells, + gcount[G] GPU counts).

    The table goes to the device un-reduced (sb_set_table: min over strategies with the first-minimum rule
    and its arg-min on the device, PerformanceEvaluator.py:101-115); the search runs on the reduced view
    (a slower strategy at the same GPU count is dominated).  Returns the reference's 6-tuple
    (sta, tga, bss, bna, boa, makespan) — bss[t] is one-hot over the G gpu-count columns, the option order
    task.strategies has after profiling — plus strategy[J], the index of the winning strategy (executor) of
    each task's chosen cell.  For the same seed and population the plan equals solve() on the
    `strategies_from_table` view."""
    from .search import run_search
    T = np.ascontiguousarray(T, dtype=np.float32)
    if T.ndim != 3:
        raise SolverError("T must be [J][S][G]")
    J, S, G = T.shape
    if J == 0:
        return [[[] for _ in range(NSLOT)]], [], [], [], [], 0.0, np.zeros(0, dtype=np.int64)
    gcount = list(range(1, G + 1)) if gcount is None else [int(g) for g in gcount]
    if len(gcount) != G or len(set(gcount)) != G or not all(1 <= g <= NSLOT for g in gcount):
        raise SolverError("gcount must hold %d distinct GPU counts in 1..%d" % (G, NSLOT))
    usable = np.isfinite(T) & (T >= 0) if mask is None else (np.asarray(mask, dtype=bool) & np.isfinite(T))
    if mask is None:
        usable &= T < NOT_PROFILED
    Tdev = np.where(usable, T, np.inf).astype(np.float32)
    for j in np.nonzero(~usable.reshape(J, -1).any(axis=1))[0]:
        Tdev[j] = np.where(np.isfinite(T[j]), T[j], np.inf)      # nothing usable: the sentinels are all it has
    if not np.isfinite(Tdev.reshape(J, -1)).any(axis=1).all():
        raise SolverError("a task has no finite cell in T")
    _check_horizon(Tdev)
    eng = engine if engine is not None else _engine(devices)
    nodes = int(_default_nodes() if nodes is None else nodes)
    eng.set_table(Tdev, gcount, sentinel=float("inf"), nodes=nodes)
    if chains is None:
        chains = int(os.environ.get("SATURN_B200_CHAINS", 0))
        if chains <= 0:
            wave = eng.search_wave(reduced=True)
            chains = max(1, round((1 << 17) / wave)) * wave if wave > 0 else 1 << 17
    if rounds is None:
        rounds = int(os.environ.get("SATURN_B200_ROUNDS", 400))
    budget = min(float(os.environ.get("SATURN_B200_BUDGET_S", 20.0)), float(timeout))
    col_of_k = {g: gi for gi, g in enumerate(gcount)}
    warm = None
    if presolved is not None:
        class _Opt:                      # the option list a task has in this view: every gpu-count column
            strategies = {g: None for g in gcount}
        warm = candidate_from_arrays([_Opt] * J, presolved, nodes)
    res = run_search(eng, chains=chains, rounds=rounds, seed=seed, integer_starts=integer_starts, reduced=True,
                     time_budget_s=budget, patience=max(40, rounds // 4), warm=warm)
    _check_horizon(Tdev, res.makespan)
    dec = eng.decode(res.opt, res.prio, integer_starts=integer_starts, reduced=True)
    gpus = dec["gpus"].astype(np.int64)
    chosen = np.array([col_of_k[int(k)] for k in gpus], dtype=np.int64)
    position = np.empty(J, dtype=np.int64)
    position[res.prio.astype(np.int64)] = np.arange(J)
    arrays = plan_to_arrays([G] * J, chosen, dec["start"], dec["slotmask"], position, nodes=nodes, node_of=dec["node"])
    strategy = dec["strategy"].astype(np.int64)
    makespan = max(float(dec["start"][j]) + float(T[j, strategy[j], chosen[j]]) for j in range(J))
    global last_stats
    last_stats = {"candidates": res.evaluated, "rounds": res.rounds, "search_wall_s": res.wall_s,
                  "device_makespan": res.makespan, "makespan": makespan, "J": J, "chains": chains, "nodes": nodes,
                  "devices": len(getattr(eng, "engines", [eng])), "total_wall_s": None, "adopted": True}
    return arrays + (makespan, strategy)


# ------------------------------------------------------------------------------------------ decode
def convert_into_comprehensible(task_list, bss, boa, tga, bna, sta):
    """Drop-in for saturn.solver.convert_into_comprehensible (milp.py:448-513).

    Returns (node_per_task: dict Task -> int, task_dependency_dict: defaultdict Task -> [Task],
    start_time_per_task: list[float]) and, as the reference does, records the chosen option on
    each task via task.select_strategy (milp.py:475-486).  The O(J^2 * G) Python triple loop of
    milp.py:492-511 is replaced by bit-mask arithmetic; results are identical.
    """
    task_list = list(task_list)
    J = len(task_list)
    node_per_task = {}
    nodes = np.zeros(J, dtype=np.int64)
    for idx, task in enumerate(task_list):
        n = np.argmax(bna[idx])
        node_per_task[task] = n
        nodes[idx] = int(n)
    for idx, task in enumerate(task_list):
        want = int(np.argmax(bss[idx]))
        for ctr, strat in enumerate(task.strategies.values()):
            if ctr == want:
                task.select_strategy(strat)
                break
    # occupancy of each task on its node: round(tga[t][n][g]) == 1  (milp.py:490-497), as one array operation
    tga_a = np.asarray(tga, dtype=np.float64).reshape(J, -1, NSLOT)
    occ = np.rint(tga_a[np.arange(J), nodes, :]) == 1                                # [J][8]
    if not occ.any(axis=1).all():
        raise SolverError("task %d occupies no GPU in tga" % int(np.argmin(occ.any(axis=1))))
    first = np.argmax(occ, axis=1)
    masks = (occ.astype(np.int64) << np.arange(NSLOT)[None, :]).sum(axis=1)
    start_time_per_task = [sta[int(nodes[idx])][int(first[idx])][idx] for idx in range(J)]
    task_dependency_dict = defaultdict(list)
    if J > 1:
        # before[p][t]: round(boa[p][t]) == 1; the diagonal (None in the reference, milp.py:263-270) never counts
        b = np.array(boa, dtype=object)
        b[np.equal(b, None)] = 0.0
        before = np.rint(b.astype(np.float64)) == 1
        np.fill_diagonal(before, False)
        share = ((masks[:, None] & masks[None, :]) != 0) & (nodes[:, None] == nodes[None, :])
        dep = before & share                       # dep[p][t]: p must finish before t launches
        for idx in np.nonzero(dep.any(axis=0))[0]:
            task_dependency_dict[task_list[int(idx)]] = [task_list[int(p)] for p in np.nonzero(dep[:, idx])[0]]
    return node_per_task, task_dependency_dict, start_time_per_task
