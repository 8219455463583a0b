"""CPU, world_size 2, gloo: the N>1 host logic of the search driver — sharding by global chain id,
the per-round MIN all-reduce of the packed (makespan bits << 32 | id) key, ownership and broadcast of
the winning encoding.  A numpy stand-in (scored by the oracle) replaces the CUDA engine; the real
engine's kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_eval as R


class FakeEngine:
    """Implements the Engine search surface run_search uses, on the CPU."""

    def __init__(self, tab, valid):
        self.device = torch.device("cpu")
        self.tab, self.valid = tab, valid
        self.J = tab.shape[0]
        self._key = torch.full((1,), 2 ** 63 - 1, dtype=torch.int64)

    def reduced_table(self):
        return R.reduce_table(self.tab.astype(np.float32))

    def sync(self):
        pass

    def _score(self, opt, prio):
        return R.list_schedule_batch(self.tab, opt, prio, True, np.float32)

    def _fold(self, mk, first_id):
        for b, m in enumerate(mk):
            key = (int(np.float32(m).view(np.uint32)) << 32) | (first_id + b)
            if key < int(self._key.item()):
                self._key.fill_(key)
                self._best = (self.opt[b].copy(), self.prio[b].copy()) if first_id == self.base else self._inj

    def search_init(self, chains, seed=0, chain_base=0, **kw):
        self.chains, self.base = chains, chain_base
        self.opt, self.prio = R.synth_candidates(self.J, chains, self.valid, seed=seed + chain_base)
        self.rng = np.random.default_rng(seed + chain_base)
        self.mk = self._score(self.opt, self.prio)
        self._inj = None
        self._fold(self.mk, chain_base)
        self.evaluated = chains

    def search_resample(self):
        pass

    def search_inject(self, opt, prio, copies=1, first=-1):
        self._inj = (np.asarray(opt).copy(), np.asarray(prio).copy())
        mk = self._score(opt[None, :], prio[None, :])
        if first < 0:
            first = self.chains - copies
        key = (int(np.float32(mk[0]).view(np.uint32)) << 32) | (self.base + first)
        if key < int(self._key.item()):
            self._key.fill_(key)
            self._best = self._inj
        self.evaluated += copies

    def search_best_key(self):
        return self._key

    def search_round(self, n=1):
        for _ in range(n):
            for b in range(self.chains):
                i, j = self.rng.integers(0, self.J, size=2)
                p = self.prio[b].copy()
                p[i], p[j] = p[j], p[i]
                m = self._score(self.opt[b:b + 1], p[None, :])[0]
                if m <= self.mk[b]:
                    self.prio[b], self.mk[b] = p, m
            self._fold(self.mk, self.base)
            self.evaluated += self.chains

    def search_best(self):
        k = int(self._key.item())
        return self._best[0], self._best[1], 0.0, k

    def search_stats(self):
        return self.evaluated, 0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from saturn_b200.search import key_makespan, run_search
        T, valid = R.synth_table(12, 2, 8, seed=1)
        tab = R.canon_table(T, range(1, 9))
        eng = FakeEngine(tab, valid)
        res = run_search(eng, chains=16, rounds=5, seed=3, use_dist=True)
        # every rank ends with the SAME incumbent and it evaluates to the agreed makespan
        mk = float(R.list_schedule(tab, res.opt, res.prio, True, np.float32)[0])
        local_best = key_makespan(int(eng.search_best_key().item()))
        q.put((rank, res.makespan, mk, res.opt.tolist(), res.prio.tolist(), res.evaluated, local_best, res.owner_rank))
    except Exception as e:          # surface worker failures instead of a queue timeout
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_search_exchange():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=240) for _ in range(2))
    assert all(len(o) == 8 for o in out), out
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mk0, ev0, o0, p0, n0, lb0, own0), (r1, mk1, ev1, o1, p1, n1, lb1, own1) = out
    assert mk0 == mk1 == ev0 == ev1                  # agreed, and the broadcast encoding really scores it
    assert o0 == o1 and p0 == p1
    assert mk0 == min(lb0, lb1)                      # the all-reduce(MIN) picked the better rank
    assert own0 == own1 and own0 == (0 if lb0 <= lb1 else 1)
    assert n0 == n1 and n0 >= 2 * 16 * 6             # whole-job candidate count over both ranks
