"""Search driver: rounds of device-side Metropolis search with one MIN all-reduce per round.

This is the host loop around the C ABI's sb_search_* calls.  With `torch.distributed`
initialised (one process per GPU, NCCL) the population is sharded by global chain id: rank r
owns chains [r*chains, (r+1)*chains); after every round the ranks exchange ONE packed
uint64 — (fp32 makespan bits << 32) | global chain id — with all_reduce(MIN), which is an
arg-min because non-negative floats order like their bit patterns.  The winning
encoding is broadcast from its owner when the search ends (and when elites are re-seeded).
There is no reference counterpart: the reference solver is a single-process CPU MILP.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch

from .engine import Engine


@dataclass
class SearchResult:
    opt: np.ndarray
    prio: np.ndarray
    makespan: float
    evaluated: int          # candidates scored by all ranks
    rounds: int
    wall_s: float
    history: List[Tuple[float, int, float]] = field(default_factory=list)  # (wall s, evaluated, best makespan)
    owner_rank: int = 0


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def key_makespan(key: int) -> float:
    return float(np.array([(key >> 32) & 0xffffffff], dtype=np.uint32).view(np.float32)[0])


def lpt_seeds(tmin: np.ndarray, sentinel: float = 1.0e6, nodes: int = 1):
    """Heuristic warm candidates in the reduced encoding (opt byte = k-1, plus node << 3 when there
    are several nodes), longest-processing-time order: (a) every job on its fastest option,
    (b) every job on its least GPU-seconds option, (c) in between.  Nodes are filled greedily by
    accumulated GPU-seconds."""
    J = tmin.shape[0]
    usable = np.where(tmin < sentinel, tmin, np.inf)
    if not np.isfinite(usable).any(axis=1).all():
        usable = np.where(np.isfinite(tmin), tmin, np.inf)
    k = np.arange(1, 9, dtype=np.float64)[None, :]
    seeds = []
    for area_weight in (0.0, 1.0, 0.5):
        cost = usable.astype(np.float64) * (k ** area_weight)
        col = np.argmin(cost, axis=1)
        rt = usable[np.arange(J), col]
        order = np.argsort(-rt * (col + 1) ** 0.5, kind="stable")
        ob = col.astype(np.uint8)
        if nodes > 1:
            load = np.zeros(nodes)
            ob = ob.copy()
            for j in order:
                n = int(np.argmin(load))
                load[n] += float(rt[j]) * (int(col[j]) + 1)
                ob[j] |= n << 3
        seeds.append((ob, order))
    return seeds


def run_search(engine: Engine, chains: int = 1 << 16, rounds: int = 200, seed: int = 0,
               integer_starts: bool = True, reduced: bool = False, time_budget_s: Optional[float] = None,
               patience: Optional[int] = None, t_start: float = 5e-4, t_end: float = 1e-6,
               warm: Optional[Tuple[np.ndarray, np.ndarray]] = None, use_dist: bool = True,
               target_makespan: Optional[float] = None, reseed_every: int = 0, resample_every: Optional[int] = None,
               record_history: bool = False, heuristic_seeds: bool = True,
               exchange_every: int = 16, _no_fused: bool = False, _python_driver: bool = False,
               _extra_flags: int = 0) -> SearchResult:
    """Run the search on `engine` (table already set).  Returns the best candidate found by any rank.

    `rounds` device rounds are issued in groups of `exchange_every` (tournament resampling every
    `resample_every` rounds inside a group is only another launch); after each group the ranks exchange
    their best key (one MIN) and the stopping rules are evaluated, so the host synchronises once per
    group rather than once per round."""
    if resample_every is None:
        resample_every = -1          # the library's choice: 2 inside the tile kernel, 4 where it costs a copy
    dist = _dist() if use_dist else None
    if dist is None and not reseed_every and not _python_driver and hasattr(engine, "search_run"):
        # one process, one GPU: the same loop runs inside the library (sb_search_run)
        r = engine.search_run(chains, rounds, seed=seed, integer_starts=integer_starts, reduced=reduced,
                              t_start=t_start, t_end=t_end, warm=warm, resample_every=resample_every,
                              sync_every=exchange_every, patience=patience or 0, time_budget_s=time_budget_s or 0.0,
                              target_makespan=target_makespan or 0.0, heuristic_seeds=heuristic_seeds,
                              record_history=record_history, _no_fused=_no_fused, _extra_flags=_extra_flags)
        return SearchResult(opt=r["opt"], prio=r["prio"], makespan=r["makespan"], evaluated=r["evaluated"],
                            rounds=r["rounds"], wall_s=r["wall_s"], history=r["history"], owner_rank=0)
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    J = engine.J
    pdt = np.uint8 if J <= 256 else np.uint16
    t0 = time.perf_counter()
    engine.search_init(chains, seed=seed, chain_base=rank * chains, integer_starts=integer_starts,
                       reduced=reduced, t_start=t_start, t_end=t_end, total_rounds=max(rounds, 1), warm=warm,
                       resample_every=resample_every, **({"_no_fused": True} if _no_fused else {}),
                       **({"_extra_flags": _extra_flags} if _extra_flags else {}))
    if heuristic_seeds:
        # every rank plants the longest-processing-time seeds in an eighth of its population each;
        # the rest stays random (diversity), tournament resampling then concentrates the population
        tmin, args = engine.reduced_table()
        per = max(1, chains // 8)
        nodes = getattr(engine, "nodes", 1)
        for i, (col, order) in enumerate(lpt_seeds(tmin, nodes=nodes)):
            opt = col if reduced else ((args[np.arange(J), col & 7].astype(np.uint8) << 3) | col)
            first = min(i * per, max(0, chains - per))
            engine.search_inject(opt.astype(np.uint8), order.astype(pdt), copies=min(per, chains), first=first)
    local_key = engine.search_best_key()          # aliases device memory
    KEY_MAX = 0x7fffffffffffffff
    gkey = torch.full((1,), KEY_MAX, dtype=torch.int64, device=engine.device)
    history: List[Tuple[float, int, float]] = []
    best_seen = None
    stale = 0
    done_rounds = 0
    stop = torch.zeros(1, dtype=torch.int32, device=engine.device)

    if dist and not getattr(engine, "has_xchg", False) and hasattr(engine, "xchg_init") \
            and dist.get_backend() == "nccl":
        engine.xchg_init(dist)              # NVLink peer-memory mailboxes; collectively False -> NCCL on all ranks
    use_xchg = bool(dist) and getattr(engine, "has_xchg", False)

    def exchange() -> int:
        if use_xchg:
            # NVLink peer-memory MIN: every rank publishes its key in its own mailbox, a one-warp kernel
            # folds all mailboxes.  A wait that times out (a peer seconds late: a dead rank) leaves the
            # preset maximum in gkey and raises here — a loud failure, never a stale key as the owner id.
            gkey.fill_(KEY_MAX)
            engine.xchg_post(local_key)
            engine.xchg_reduce(gkey)
            engine.xchg_check()             # synchronises; raises SaturnB200Error on a timed-out wait
            k = int(gkey.item())
            if k == KEY_MAX:
                raise RuntimeError("peer exchange produced no key")
            return k
        engine.sync()
        gkey.copy_(local_key)
        if dist:
            dist.all_reduce(gkey, op=dist.ReduceOp.MIN)
        return int(gkey.item())

    key = exchange()
    best_seen = key
    if record_history:
        history.append((time.perf_counter() - t0, chains * world, key_makespan(key)))
    exchange_every = max(1, int(exchange_every))
    while done_rounds < rounds:
        # one group of rounds, no host synchronisation; the library resamples on its own cadence
        step = min(exchange_every, rounds - done_rounds)
        engine.search_round(step)
        done_rounds += step
        r = done_rounds - 1
        key = exchange()
        if key < best_seen:
            best_seen = key
            stale = 0
        else:
            stale += step
        if record_history:
            history.append((time.perf_counter() - t0, chains * world * (done_rounds + 1), key_makespan(key)))
        # stopping decisions must be identical on every rank: derive them from rank 0's clock
        want_stop = False
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            want_stop = True
        if patience is not None and stale >= patience:
            want_stop = True
        if target_makespan is not None and key_makespan(best_seen) <= target_makespan:
            want_stop = True
        if dist:
            stop.fill_(1 if want_stop else 0)
            dist.broadcast(stop, src=0)
            want_stop = bool(stop.item())
        if want_stop:
            break
        if reseed_every and (r + 1) % reseed_every == 0 and r + 1 < rounds:
            opt, prio = _gather_best(engine, best_seen, chains, dist, rank)
            engine.search_inject(opt, prio, copies=max(1, chains // 64), first=-1)
    opt, prio = _gather_best(engine, best_seen, chains, dist, rank)
    ev, _ = engine.search_stats()
    total_ev = ev
    if dist:
        evt = torch.tensor([ev], dtype=torch.int64, device=engine.device)
        dist.all_reduce(evt, op=dist.ReduceOp.SUM)
        total_ev = int(evt.item())
    return SearchResult(opt=opt, prio=prio, makespan=key_makespan(best_seen), evaluated=total_ev,
                        rounds=done_rounds, wall_s=time.perf_counter() - t0, history=history,
                        owner_rank=int((best_seen & 0xffffffff) // chains) if world > 1 else 0)


def _gather_best(engine: Engine, key: int, chains: int, dist, rank: int):
    """Fetch the encoding that produced `key` from the rank that owns it."""
    J = engine.J
    opt, prio, _mk, lkey = engine.search_best()
    if dist is None:
        return opt, prio
    owner = int((key & 0xffffffff) // chains)
    dev = engine.device
    pdt = torch.uint8 if J <= 256 else torch.int32
    o = torch.from_numpy(opt.copy()).to(dev)
    p = torch.from_numpy(prio.astype(np.int32) if J > 256 else prio.copy()).to(dev).to(pdt)
    dist.broadcast(o, src=owner)
    dist.broadcast(p, src=owner)
    return o.cpu().numpy().astype(np.uint8), p.cpu().numpy().astype(np.uint8 if J <= 256 else np.uint16)
