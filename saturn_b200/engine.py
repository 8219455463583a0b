"""Host-side wrapper of the C ABI (include/saturn_b200.h) over torch tensors.

PyTorch is used only as plumbing here: device memory (`torch.Tensor.data_ptr()`), the current
CUDA stream, and `torch.distributed` for the per-round MIN all-reduce.  All arithmetic runs in
the hand-written kernels of saturn_b200/csrc.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FLAG_INTEGER_STARTS, FLAG_REDUCED, SaturnB200Error, SearchParams, check

NSLOT = 8


def _flags(integer_starts: bool, reduced: bool) -> int:
    return (FLAG_INTEGER_STARTS if integer_starts else 0) | (FLAG_REDUCED if reduced else 0)


class Engine:
    """One solver handle bound to one CUDA device."""

    def __init__(self, device: int | torch.device | None = None, stream: Optional[torch.cuda.Stream] = None):
        self._h = C.c_void_p()
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise SaturnB200Error("no CUDA device is visible; saturn_b200 has no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.device = dev
        self._lib = lib
        if stream is None:
            stream = torch.cuda.current_stream(dev)
        self._stream = stream
        sptr = C.c_void_p(stream.cuda_stream) if stream.cuda_stream else None
        check(lib.sb_create(dev.index or 0, sptr, C.byref(self._h)))
        self.J = 0
        self.S = 0
        self.G = 0
        self.gcount = None
        self.nodes = 1

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if self._h:
            self._lib.sb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self._lib.sb_sync(self._h))

    # ------------------------------------------------------------------ table
    def set_table(self, T, gcount: Optional[Sequence[int]] = None, sentinel: Optional[float] = None, nodes: int = 1):
        """T[J][S][G] fp32 (numpy array or torch tensor, host or device); gcount[G] GPU counts.
        `sentinel`: cells at/above it are never proposed by the search (default 1e6).
        `nodes` > 1: candidates are evaluated with reduced=True and opt = (node << 3) | (k - 1)."""
        if sentinel is not None:
            check(self._lib.sb_set_sentinel(self._h, C.c_float(sentinel)))
        if isinstance(T, torch.Tensor):
            Tt = T.detach().to(torch.float32).contiguous()
            J, S, G = Tt.shape
            ptr = Tt.data_ptr()
            keep = Tt
        else:
            Ta = np.ascontiguousarray(T, dtype=np.float32)
            J, S, G = Ta.shape
            ptr = Ta.ctypes.data
            keep = Ta
        if gcount is None:
            gcount = list(range(1, G + 1))
        gc = np.ascontiguousarray(gcount, dtype=np.uint8)
        if gc.shape != (G,):
            raise ValueError("gcount must have %d entries" % G)
        check(self._lib.sb_set_table(self._h, C.c_void_p(ptr), C.c_void_p(gc.ctypes.data), J, S, G, int(nodes)))
        del keep
        self.J, self.S, self.G = int(J), int(S), int(G)
        self.gcount = [int(x) for x in gc]
        self.nodes = int(nodes)
        return self

    def reduced_table(self) -> Tuple[np.ndarray, np.ndarray]:
        tmin = np.empty((self.J, NSLOT), dtype=np.float32)
        args = np.empty((self.J, NSLOT), dtype=np.uint8)
        check(self._lib.sb_get_reduced(self._h, C.c_void_p(tmin.ctypes.data), C.c_void_p(args.ctypes.data)))
        return tmin, args

    @property
    def prio_dtype(self):
        return torch.uint8 if self.J <= 256 else torch.uint16

    # ------------------------------------------------------------------ evaluation
    def _check_cands(self, opt: torch.Tensor, prio: torch.Tensor, on_device: bool):
        if opt.dtype != torch.uint8:
            raise TypeError("opt must be uint8")
        if prio.dtype != self.prio_dtype:
            raise TypeError("prio must be %s for J=%d" % (self.prio_dtype, self.J))
        if opt.dim() != 2 or prio.dim() != 2 or opt.shape != prio.shape:
            raise ValueError("opt / prio must both be [B][J]")
        if opt.shape[1] != self.J:
            raise ValueError("candidates have %d jobs, table has %d" % (opt.shape[1], self.J))
        if opt.stride(1) != 1 or prio.stride(1) != 1 or (opt.shape[0] > 1 and opt.stride(0) != prio.stride(0)):
            raise ValueError("opt / prio rows must be contiguous with the same row stride")
        if on_device and (opt.device != self.device or prio.device != self.device):
            raise ValueError("candidates must live on %s" % self.device)
        if not on_device and (opt.is_cuda or prio.is_cuda):
            raise ValueError("host evaluation takes CPU tensors")
        B = opt.shape[0]
        stride = opt.stride(0) if B > 1 else max(opt.stride(0), self.J)
        return B, stride

    def eval(self, opt: torch.Tensor, prio: torch.Tensor, integer_starts: bool = True, reduced: bool = False,
             out: Optional[torch.Tensor] = None, best_key: Optional[torch.Tensor] = None, id_base: int = 0,
             _force_generic: bool = False, _no_stream: bool = False, post_key: bool = False, fold_prev: bool = False,
             by_position: bool = False, _plain_addr: bool = False, alt_shape: bool = False,
             _table_home: int = 0, _reorder: Optional[bool] = None) -> torch.Tensor:
        """Makespan of every candidate (device tensors).  Asynchronous on the handle's stream.
        by_position: opt[b][i] is the option of the job scheduled i-th (see `opt_by_position`).
        alt_shape: the alternate warp-shuffle kernel (SB_FLAG_ALT_WARPSCAN; a measurement, not a fast path).
        Test hooks: _table_home 2 / 1 puts the position-major kernel's table in a CTA pair's shared memory / in
        global memory whatever its size; _reorder True / False forces / forbids the route that re-orders
        job-indexed opt rows on the device (path 9)."""
        B, stride = self._check_cands(opt, prio, True)
        if out is None:
            out = torch.empty(B, dtype=torch.float32, device=self.device)
        fl = _flags(integer_starts, reduced) | (_lib._FLAG_FORCE_GENERIC if _force_generic else 0) | (
            0x40000000 if _no_stream else 0) | (0x02000000 if _plain_addr else 0) | (
            _lib.FLAG_POST_KEY if post_key else 0) | (
            _lib.FLAG_FOLD_PREV if (post_key and fold_prev) else 0) | (
            _lib.FLAG_OPT_BY_POSITION if by_position else 0) | (_lib.FLAG_ALT_WARPSCAN if alt_shape else 0) | (
            {0: 0, 1: 0x00800000, 2: 0x00400000}[_table_home]) | (
            0 if _reorder is None else (0x00200000 if _reorder else 0x00100000))
        kp = C.c_void_p(best_key.data_ptr()) if best_key is not None else None
        check(self._lib.sb_eval(self._h, C.c_void_p(opt.data_ptr()), C.c_void_p(prio.data_ptr()), B, stride, fl,
                                C.c_void_p(out.data_ptr()), kp, id_base & 0xffffffff))
        return out

    def last_eval_path(self) -> int:
        return int(self._lib.sb_last_eval_path(self._h))

    def validate(self, opt: torch.Tensor, prio: torch.Tensor, reduced: bool = False) -> int:
        B, stride = self._check_cands(opt, prio, True)
        bad = C.c_int64(0)
        check(self._lib.sb_validate(self._h, C.c_void_p(opt.data_ptr()), C.c_void_p(prio.data_ptr()), B, stride,
                                    _flags(False, reduced), C.byref(bad)))
        return int(bad.value)

    def eval_host(self, opt: torch.Tensor, prio: torch.Tensor, integer_starts: bool = True, reduced: bool = False,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same through HOST tensors (pinned for full PCIe speed): H2D + kernel + D2H, synchronous."""
        B, stride = self._check_cands(opt, prio, False)
        if out is None:
            out = torch.empty(B, dtype=torch.float32, pin_memory=True)
        check(self._lib.sb_eval_host(self._h, C.c_void_p(opt.data_ptr()), C.c_void_p(prio.data_ptr()), B, stride,
                                     _flags(integer_starts, reduced), C.c_void_p(out.data_ptr())))
        return out

    def eval_full(self, opt: torch.Tensor, prio: torch.Tensor, integer_starts: bool = True, reduced: bool = False):
        """(makespan[B], start[B][J], slotmask[B][J]) — slot-exact plan of every candidate."""
        B, stride = self._check_cands(opt, prio, True)
        mk = torch.empty(B, dtype=torch.float32, device=self.device)
        start = torch.empty((B, self.J), dtype=torch.float32, device=self.device)
        mask = torch.empty((B, self.J), dtype=torch.int32, device=self.device)
        check(self._lib.sb_eval_full(self._h, C.c_void_p(opt.data_ptr()), C.c_void_p(prio.data_ptr()), B, stride,
                                     _flags(integer_starts, reduced), C.c_void_p(mk.data_ptr()),
                                     C.c_void_p(start.data_ptr()), C.c_void_p(mask.data_ptr())))
        return mk, start, mask

    def decode(self, opt: np.ndarray, prio: np.ndarray, integer_starts: bool = True, reduced: bool = False):
        """One candidate (host arrays) -> dict(start, slotmask, strategy, gpus, makespan)."""
        J = self.J
        opt = np.ascontiguousarray(opt, dtype=np.uint8)
        prio = np.ascontiguousarray(prio, dtype=np.uint8 if J <= 256 else np.uint16)
        if opt.shape != (J,) or prio.shape != (J,):
            raise ValueError("opt / prio must have J=%d entries" % J)
        start = np.empty(J, dtype=np.float32)
        mask = np.empty(J, dtype=np.uint32)
        strat = np.empty(J, dtype=np.uint8)
        gpus = np.empty(J, dtype=np.uint8)
        node = np.empty(J, dtype=np.uint8)
        mk = C.c_float(0)
        check(self._lib.sb_decode(self._h, C.c_void_p(opt.ctypes.data), C.c_void_p(prio.ctypes.data),
                                  _flags(integer_starts, reduced), C.c_void_p(start.ctypes.data),
                                  C.c_void_p(mask.ctypes.data), C.c_void_p(strat.ctypes.data),
                                  C.c_void_p(gpus.ctypes.data), C.c_void_p(node.ctypes.data), C.byref(mk)))
        return {"start": start, "slotmask": mask, "strategy": strat, "gpus": gpus, "node": node,
                "makespan": float(mk.value)}

    # ------------------------------------------------------------------ multi-GPU exchange (NVLink peer memory)
    def xchg_init(self, dist) -> bool:
        """Set up the peer-memory MIN exchange over the ranks of an initialised torch.distributed
        group (one process per GPU of one node).  The 64-byte CUDA IPC handles are all-gathered with
        the group itself; returns False (and leaves the engine on the NCCL path) if a peer mapping
        cannot be opened."""
        rank, world = dist.get_rank(), dist.get_world_size()
        hdl = np.zeros(_lib.IPC_HANDLE_BYTES, dtype=np.uint8)
        # success or failure is decided COLLECTIVELY: a rank whose local step fails still takes part in the
        # all_gather / all_reduce below, so no rank is left waiting in a collective and all ranks end up
        # with the same answer (some posting to mailboxes while others call NCCL would deadlock)
        ok_local = self._lib.sb_xchg_create(self._h, rank, world, C.c_void_p(hdl.ctypes.data)) == 0
        mine = torch.from_numpy(hdl).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        okt = torch.tensor([1 if ok_local else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)           # every rank created its mailbox?
        if bool(okt.item()):
            flat = torch.stack(allh).cpu().numpy().copy()
            ok_local = self._lib.sb_xchg_connect(self._h, C.c_void_p(flat.ctypes.data)) == 0
            okt.fill_(1 if ok_local else 0)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)       # every rank mapped every peer?
        self._xchg = bool(okt.item())
        return self._xchg

    @property
    def has_xchg(self) -> bool:
        return getattr(self, "_xchg", False)

    def xchg_post(self, key: torch.Tensor):
        check(self._lib.sb_xchg_post(self._h, C.c_void_p(key.data_ptr())))

    def xchg_reduce(self, out: torch.Tensor, fold: Optional[torch.Tensor] = None):
        """out[0] = MIN over all ranks of the keys posted this round; `fold` is MIN-ed in place."""
        check(self._lib.sb_xchg_reduce(self._h, C.c_void_p(out.data_ptr()),
                                       C.c_void_p(fold.data_ptr()) if fold is not None else None))

    def xchg_check(self):
        check(self._lib.sb_xchg_check(self._h))

    # ------------------------------------------------------------------ search
    def search_init(self, chains: int, seed: int = 0, chain_base: int = 0, integer_starts: bool = True,
                    reduced: bool = False, t_start: float = 0.02, t_end: float = 1e-4, total_rounds: int = 200,
                    warm: Optional[Tuple[np.ndarray, np.ndarray]] = None, resample_every: int = 0,
                    _no_fused: bool = False, _extra_flags: int = 0):
        """resample_every > 0: search_round resamples the population by tournament on that cadence itself.
        _extra_flags: test hooks of sb_search_params.flags (see sb_search_verify_count in the header)."""
        p = SearchParams(seed=seed, chains=chains, chain_base=chain_base, resample_every=int(resample_every or 0),
                         flags=_flags(integer_starts, reduced) | (0x20000000 if _no_fused else 0) | int(_extra_flags),
                         t_start=t_start, t_end=t_end, total_rounds=total_rounds)
        wo = wp = None
        keep = None
        if warm is not None:
            o = np.ascontiguousarray(warm[0], dtype=np.uint8)
            pr = np.ascontiguousarray(warm[1], dtype=np.uint8 if self.J <= 256 else np.uint16)
            keep = (o, pr)
            wo, wp = C.c_void_p(o.ctypes.data), C.c_void_p(pr.ctypes.data)
        check(self._lib.sb_search_init(self._h, C.byref(p), wo, wp))
        del keep
        self._search_chains = chains
        self._search_base = chain_base

    def search_seed_lpt(self):
        """Plant the three longest-processing-time seeds into an eighth of the population each."""
        check(self._lib.sb_search_seed_lpt(self._h))

    def search_run(self, chains: int, rounds: int, seed: int = 0, chain_base: int = 0, integer_starts: bool = True,
                   reduced: bool = False, t_start: float = 5e-4, t_end: float = 1e-6,
                   warm: Optional[Tuple[np.ndarray, np.ndarray]] = None, resample_every: int = -1, sync_every: int = 16,
                   patience: int = 0, time_budget_s: float = 0.0, target_makespan: float = 0.0,
                   heuristic_seeds: bool = True, record_history: bool = False, _no_fused: bool = False,
                   _extra_flags: int = 0):
        """The whole single-GPU search in one C call (sb_search_run).  Returns a dict: opt, prio, makespan, key,
        evaluated, rounds, stop_reason, wall_s, history [(wall s, evaluated, makespan)]."""
        return _search_run(self._lib, [self._h], self.J, chains, rounds, seed, chain_base, integer_starts, reduced,
                           t_start, t_end, warm, resample_every, sync_every, patience, time_budget_s, target_makespan,
                           heuristic_seeds, record_history, _no_fused, _extra_flags)

    def search_wave(self, reduced: bool = False) -> int:
        """Chains that fill the device exactly once with the round kernel of the current table; populations
        that are whole multiples of it leave no partially filled last wave."""
        n = C.c_int64(0)
        check(self._lib.sb_search_wave(self._h, _flags(False, reduced), C.byref(n)))
        return int(n.value)

    def search_is_fused(self) -> bool:
        return bool(self._lib.sb_search_is_fused(self._h))

    def search_round(self, rounds: int = 1):
        check(self._lib.sb_search_round(self._h, rounds))

    def search_best_key(self) -> torch.Tensor:
        """A 1-element int64 tensor aliasing the device-resident best key (makespan bits << 32 | id)."""
        ptr = C.c_void_p()
        check(self._lib.sb_search_best_key_ptr(self._h, C.byref(ptr)))
        return _alias_int64(ptr.value, self.device)

    def search_best(self):
        J = self.J
        opt = np.empty(J, dtype=np.uint8)
        prio = np.empty(J, dtype=np.uint8 if J <= 256 else np.uint16)
        mk = C.c_float(0)
        key = C.c_uint64(0)
        check(self._lib.sb_search_best(self._h, C.c_void_p(opt.ctypes.data), C.c_void_p(prio.ctypes.data),
                                       C.byref(mk), C.byref(key)))
        return opt, prio, float(mk.value), int(key.value)

    def search_inject(self, opt: np.ndarray, prio: np.ndarray, copies: int = 1, first: int = -1):
        o = np.ascontiguousarray(opt, dtype=np.uint8)
        p = np.ascontiguousarray(prio, dtype=np.uint8 if self.J <= 256 else np.uint16)
        check(self._lib.sb_search_inject(self._h, C.c_void_p(o.ctypes.data), C.c_void_p(p.ctypes.data), first,
                                         copies))

    def search_resample(self):
        check(self._lib.sb_search_resample(self._h))

    def search_verify_count(self) -> int:
        """Incremental scores that differed from a from-scratch score (test hook, needs _extra_flags 0x08000000)."""
        n = C.c_uint64(0)
        check(self._lib.sb_search_verify_count(self._h, C.byref(n)))
        return int(n.value)

    def search_validate(self) -> int:
        """Chains of the current population whose rows are not a permutation + existing table cells."""
        bad = C.c_int64(0)
        check(self._lib.sb_search_validate(self._h, C.byref(bad)))
        return int(bad.value)

    def search_stats(self):
        ev, rd = C.c_int64(0), C.c_int64(0)
        check(self._lib.sb_search_stats(self._h, C.byref(ev), C.byref(rd)))
        return int(ev.value), int(rd.value)


def _search_run(lib, handles, J, chains, rounds, seed, chain_base, integer_starts, reduced, t_start, t_end, warm,
                resample_every, sync_every, patience, time_budget_s, target_makespan, heuristic_seeds,
                record_history, _no_fused, _extra_flags=0):
    """sb_search_run (one handle) / sb_search_run_multi (one handle per device of this process)."""
    pdt = np.uint8 if J <= 256 else np.uint16
    p = SearchParams(seed=seed, chains=chains, chain_base=chain_base,
                     flags=_flags(integer_starts, reduced) | (0x20000000 if _no_fused else 0) | int(_extra_flags),
                     t_start=t_start, t_end=t_end, total_rounds=max(rounds, 1))
    cap = (max(rounds, 1) // max(1, sync_every) + 3) if record_history else 0
    hw, he, hm = np.zeros(cap, np.float64), np.zeros(cap, np.int64), np.zeros(cap, np.float32)
    hl = C.c_int(0)
    ctl = _lib.SearchControl(rounds=max(rounds, 1), resample_every=int(resample_every), sync_every=max(1, int(sync_every)),
                             patience=int(patience or 0), heuristic_seeds=1 if heuristic_seeds else 0,
                             target_makespan=float(target_makespan or 0.0), time_budget_s=float(time_budget_s or 0.0),
                             history_cap=cap, history_len=C.pointer(hl),
                             history_wall_s=hw.ctypes.data_as(C.POINTER(C.c_double)),
                             history_evaluated=he.ctypes.data_as(C.POINTER(C.c_int64)),
                             history_makespan=hm.ctypes.data_as(C.POINTER(C.c_float)))
    wo = wp = None
    keep = None
    if warm is not None:
        keep = (np.ascontiguousarray(warm[0], dtype=np.uint8), np.ascontiguousarray(warm[1], dtype=pdt))
        wo, wp = C.c_void_p(keep[0].ctypes.data), C.c_void_p(keep[1].ctypes.data)
    opt, prio = np.empty(J, dtype=np.uint8), np.empty(J, dtype=pdt)
    res = _lib.SearchResultC()
    if len(handles) == 1:
        check(lib.sb_search_run(handles[0], C.byref(p), C.byref(ctl), wo, wp, C.c_void_p(opt.ctypes.data),
                                C.c_void_p(prio.ctypes.data), C.byref(res)))
    else:
        arr = (C.c_void_p * len(handles))(*[h.value for h in handles])
        check(lib.sb_search_run_multi(arr, len(handles), C.byref(p), C.byref(ctl), wo, wp,
                                      C.c_void_p(opt.ctypes.data), C.c_void_p(prio.ctypes.data), C.byref(res)))
    del keep
    n = int(hl.value)
    return {"opt": opt, "prio": prio, "makespan": float(res.makespan), "key": int(res.key),
            "evaluated": int(res.evaluated), "rounds": int(res.rounds), "stop_reason": int(res.stop_reason),
            "wall_s": float(res.wall_s), "history": [(float(hw[i]), int(he[i]), float(hm[i])) for i in range(n)]}


class MultiEngine:
    """N handles on N devices of THIS process behind the interface `saturn.solver.solve` uses: the table is
    replicated, the search population is sharded by global chain id (sb_search_run_multi: one MIN of a uint64
    per group of rounds over NVLink peer memory, no torchrun, no NCCL), the winner is decoded on the first
    device.  The reference calls its solver from one process (saturn/orchestrator.py:21-23,55,69); this is how
    that call site reaches every GPU of the node."""

    def __init__(self, devices):
        if isinstance(devices, int):
            devices = list(range(devices))
        devices = [int(d) for d in devices]
        if len(devices) < 1:
            raise ValueError("need at least one device")
        if len(set(devices)) != len(devices):
            raise ValueError("devices must be distinct")
        self.engines = [Engine(d, stream=torch.cuda.current_stream(torch.device("cuda", d))) for d in devices]
        self._lib = self.engines[0]._lib
        self.device = self.engines[0].device
        self.devices = devices
        self._pool = None

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        for e in self.engines:
            e.close()

    def set_table(self, T, gcount=None, sentinel=None, nodes: int = 1):
        # one host thread per device: sb_set_table synchronises its stream (ctypes releases the GIL in the call)
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=len(self.engines))
        list(self._pool.map(lambda e: e.set_table(T, gcount, sentinel=sentinel, nodes=nodes), self.engines))
        return self

    J = property(lambda self: self.engines[0].J)
    S = property(lambda self: self.engines[0].S)
    G = property(lambda self: self.engines[0].G)
    gcount = property(lambda self: self.engines[0].gcount)
    nodes = property(lambda self: self.engines[0].nodes)
    prio_dtype = property(lambda self: self.engines[0].prio_dtype)

    def reduced_table(self):
        return self.engines[0].reduced_table()

    def decode(self, *a, **kw):
        return self.engines[0].decode(*a, **kw)

    def search_wave(self, reduced: bool = False) -> int:
        """chains PER DEVICE that fill one device exactly once."""
        return self.engines[0].search_wave(reduced)

    def search_run(self, chains: int, rounds: int, seed: int = 0, chain_base: int = 0, integer_starts: bool = True,
                   reduced: bool = False, t_start: float = 5e-4, t_end: float = 1e-6, warm=None,
                   resample_every: int = -1, sync_every: int = 16, patience: int = 0, time_budget_s: float = 0.0,
                   target_makespan: float = 0.0, heuristic_seeds: bool = True, record_history: bool = False,
                   _no_fused: bool = False, _extra_flags: int = 0):
        """`chains` is per device; the result's `evaluated` counts every device."""
        return _search_run(self._lib, [e._h for e in self.engines], self.J, chains, rounds, seed, chain_base,
                           integer_starts, reduced, t_start, t_end, warm, resample_every, sync_every, patience,
                           time_budget_s, target_makespan, heuristic_seeds, record_history, _no_fused, _extra_flags)


class _CudaArrayView:
    """Minimal __cuda_array_interface__ carrier so torch can alias library-owned device memory."""

    def __init__(self, ptr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}
        self._nbytes = nbytes


def _alias_int64(ptr: int, device: torch.device) -> torch.Tensor:
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArrayView(ptr, 8, "<i8", (1,)), device=device)


# ---------------------------------------------------------------------- candidate helpers
def opt_by_position(opt: torch.Tensor, prio: torch.Tensor) -> torch.Tensor:
    """Re-encode job-indexed opt rows in schedule order (opt'[b][i] = opt[b][prio[b][i]]), keeping the row
    stride of `opt` — the encoding SB_FLAG_OPT_BY_POSITION evaluates."""
    B, J = opt.shape
    out = padded_rows(B, J, torch.uint8, opt.device)
    out.copy_(torch.gather(opt, 1, prio.to(torch.int32).to(torch.int64)))
    return out


def padded_rows(B: int, J: int, dtype: torch.dtype, device, pinned: bool = False) -> torch.Tensor:
    """A [B][J] view into storage whose rows are a multiple of 32 ELEMENTS apart, so that the byte
    stride of both opt (u8) and prio (u8/u16) rows is 32-byte aligned (the kernel's fast path)."""
    stride = (J + 31) // 32 * 32
    if pinned:
        buf = torch.zeros((B, stride), dtype=dtype, pin_memory=True)
    else:
        buf = torch.zeros((B, stride), dtype=dtype, device=device)
    return buf[:, :J]


def random_candidates(engine: Engine, B: int, valid: np.ndarray, seed: int = 0, gcount=None, device=None,
                      pinned: bool = False, nodes: int = 1):
    """opt ~ U{valid cells of each job}, prio = random permutations (torch RNG on `device`).
    nodes > 1: `valid` must describe the reduced table (S = 1); a uniform node index is OR-ed into
    bits 3.. of every opt byte."""
    J = engine.J
    device = engine.device if device is None else torch.device(device)
    gen_dev = device if device.type == "cuda" else torch.device("cpu")
    g = torch.Generator(device=gen_dev)
    g.manual_seed(seed)
    _, S, G = valid.shape
    gcount = engine.gcount if gcount is None else list(gcount)
    opt = padded_rows(B, J, torch.uint8, device, pinned)
    prio = padded_rows(B, J, engine.prio_dtype, device, pinned)
    nmax = int(valid.reshape(J, -1).sum(axis=1).max())
    cells = np.zeros((J, nmax), dtype=np.uint8)
    ncell = np.zeros(J, dtype=np.int64)
    for j in range(J):
        c = [(s << 3) | (gcount[gi] - 1) for s in range(S) for gi in range(G) if valid[j, s, gi]]
        cells[j, :len(c)] = c
        ncell[j] = len(c)
    cells_t = torch.from_numpy(cells).to(gen_dev)
    ncell_t = torch.from_numpy(ncell).to(gen_dev)
    chunk = max(1, min(B, (1 << 24) // max(J, 1)))
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        u = torch.rand((nb, J), generator=g, device=gen_dev)
        pick = torch.minimum((u * ncell_t[None, :]).long(), ncell_t[None, :] - 1)
        o = torch.gather(cells_t[None, :, :].expand(nb, -1, -1), 2, pick[:, :, None])[:, :, 0]
        keys = torch.rand((nb, J), generator=g, device=gen_dev)
        p = torch.argsort(keys, dim=1)
        if nodes > 1:
            nd = torch.randint(0, nodes, (nb, J), generator=g, device=gen_dev, dtype=torch.uint8)
            o = o | (nd << 3)
        opt[b0:b0 + nb].copy_(o)
        if engine.prio_dtype == torch.uint8:
            prio[b0:b0 + nb].copy_(p.to(torch.uint8))
        else:
            prio[b0:b0 + nb].copy_(p.to(torch.int32).to(torch.uint16))
    return opt, prio
