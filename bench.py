#!/usr/bin/env python
"""bench.py — candidate schedules evaluated / second on the SPASE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic candidates per GPU:
  sb_eval (k_eval_tiles) over B candidates of the BASELINE C4 workload (J=256 jobs, S=8
  strategies, G=1..8 GPUs; configs[3] of BASELINE.json, which fits one GPU), folding the 64-bit
  arg-min key, then — for N > 1 — ONE all_reduce(MIN) of that key over NCCL (the only exchange the
  path has; candidates shard by id, no data-path collective).  Weak scaling: B per GPU is fixed.

`value`  = candidates scored by all ranks / device time of the K steps (inputs resident in HBM).
`e2e`    = same metric through the public host-buffer call (Engine.eval_host -> sb_eval_host):
           candidate encodings start in pinned HOST memory, H2D + kernel + D2H of the makespans
           inside the timed region.
`roofline` = algorithmic bytes (J*(1+w)+4 per candidate) of one k_eval_tiles launch /
           its CUDA-event duration, against the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference` = the oracle's C restatement (oracle/ref_eval.c, kind "port")
           on the host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "candidate schedules/sec"
UNIT = "candidates/s"
J, S, G = 256, 8, 8
WORKLOAD = "C4: J=256 jobs x S=8 strategies x G=1..8 GPUs, synthetic T (seed 0), integer starts"
WAVE = 148 * 8 * 32           # candidates in one full wave of 32-candidate tiles (148 SMs x 8 warps)
B_PER_GPU = WAVE * 27         # 1,022,976 candidates = 528 MB of encodings per step (> 126 MB L2)
B_E2E = WAVE * 8              # host-buffer batch per step
FALLBACK_HBM_GBS = 6650.0


def bytes_per_candidate(j):
    w = 1 if j <= 256 else 2
    return j * (1 + w) + 4


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per k_eval_tiles launch from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("algorithmic_bytes_per_launch")
    except Exception:
        return None, None


def cpu_eval_rate(threads, seconds_target=12.0, seed=0):
    """The oracle's C restatement timed on the host: candidates/s on a bounded sample of C4."""
    from oracle import c_oracle, ref_eval as R
    T, valid = R.synth_table(J, S, G, seed=0)
    tab = R.canon_table(T, range(1, G + 1))
    opt, prio = R.synth_candidates(J, 20000, valid, seed=seed)
    t0 = time.perf_counter()
    c_oracle.evaluate(tab, opt, prio, True, np.float32, threads=threads)
    dt = time.perf_counter() - t0
    n = int(min(max(20000, 20000 * seconds_target / max(dt, 1e-3)), 16_000_000))
    reps = max(1, n // 20000)
    t0 = time.perf_counter()
    for _ in range(reps):
        c_oracle.evaluate(tab, opt, prio, True, np.float32, threads=threads)
    dt = time.perf_counter() - t0
    return reps * 20000 / dt, reps * 20000, dt


def run_reference(args):
    """--impl reference: the CPU restatement of the path (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)      # torchrun pins it to 1; the CPU arm may use every core
    from oracle import c_oracle, ref_eval as R
    threads = c_oracle.max_threads()
    T, valid = R.synth_table(J, S, G, seed=0)
    tab = R.canon_table(T, range(1, G + 1))
    per_step = 200000
    opt, prio = R.synth_candidates(J, per_step, valid, seed=1)
    for _ in range(args.warmup):
        c_oracle.evaluate(tab, opt, prio, True, np.float32, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.evaluate(tab, opt, prio, True, np.float32, threads=threads)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = "%d candidates/step of C4 (J=256,S=8,G=8), oracle/ref_eval.c fp32 integer starts, OpenMP" % per_step
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "candidates_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--real", action="store_true", help="real-valued starts instead of integer starts")
    ap.add_argument("--config", default="C4", choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE config shape (C4 is the headline; the others are diagnostic runs)")
    ap.add_argument("--reduced", action="store_true", help="evaluate on the min-over-strategies table")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from saturn_b200.synth import synth_table
    from saturn_b200.engine import Engine, padded_rows, random_candidates

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    ints = not args.real

    eng = Engine(local)
    global J, S, G, WORKLOAD
    if args.config != "C4":
        from saturn_b200.synth import CONFIGS
        J, S, G, _seed = CONFIGS[args.config]
        WORKLOAD = "%s: J=%d jobs x S=%d strategies x G=1..%d GPUs, synthetic T (diagnostic, not the headline)" % (
            args.config, J, S, G)
        if args.batch == B_PER_GPU:
            args.batch = max(WAVE, (B_PER_GPU * 256 // J) // WAVE * WAVE)
    T, valid = synth_table(J, S, G, seed=0)
    if args.reduced:
        import numpy as _np
        T = _np.where(valid, T, _np.inf).min(axis=1, keepdims=True).astype(_np.float32)
        valid = _np.isfinite(T)
        T = _np.where(valid, T, 1e8).astype(_np.float32)
    eng.set_table(T)
    B = args.batch
    opt, prio = random_candidates(eng, B, valid, seed=1 + rank)
    out = torch.empty(B, dtype=torch.float32, device=dev)
    key = torch.full((1,), 2 ** 63 - 1, dtype=torch.int64, device=dev)
    id_base = (rank * B) & 0xffffffff

    # the per-step exchange: NVLink peer-memory MIN (post fused into the evaluation kernel's tail, then a
    # one-warp fold), falling back to one NCCL all-reduce of the key if the IPC mappings cannot be opened
    use_xchg = False
    if world > 1 and os.environ.get("SATURN_B200_EXCHANGE", "peer") == "peer":
        try:
            use_xchg = eng.xchg_init(dist)
        except Exception:
            use_xchg = False
    gmin = torch.zeros(1, dtype=torch.int64, device=dev)

    pipelined = use_xchg and os.environ.get("SATURN_B200_EXCHANGE_PIPELINE", "1") != "0"

    def exchange_after_eval():
        if use_xchg:
            if not pipelined:
                eng.xchg_reduce(gmin, fold=key)             # the running best becomes the global one
        else:
            dist.all_reduce(key, op=dist.ReduceOp.MIN)

    def step():
        # pipelined: ONE kernel per step evaluates the batch, folds the keys every rank published in the
        # previous step (prologue, NVLink loads) and publishes this step's key (tail)
        eng.eval(opt, prio, integer_starts=ints, out=out, best_key=key, id_base=id_base,
                 post_key=use_xchg, fold_prev=pipelined)
        if world > 1:
            exchange_after_eval()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    if args.config == "C4":
        assert eng.last_eval_path() == 3, "bench must run the TMA + streaming tile kernel"
    kernel_path = eng.last_eval_path()

    # ---- timed region: K steps, device time, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        k_ev[i][0].record()
        eng.eval(opt, prio, integer_starts=ints, out=out, best_key=key, id_base=id_base, post_key=use_xchg,
                 fold_prev=pipelined)
        k_ev[i][1].record()
        if world > 1:
            exchange_after_eval()
    if pipelined:
        eng.xchg_reduce(gmin, fold=key)                     # fold the last step's keys inside the timed region
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if use_xchg:
        eng.xchg_check()
    ms_total = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the public call
    e2e = None
    if not args.no_e2e:
        Be = min(B_E2E, B)
        oh, ph = random_candidates(eng, Be, valid, seed=100 + rank, device="cpu", pinned=True)
        outh = torch.empty(Be, dtype=torch.float32, pin_memory=True)
        for _ in range(2):
            eng.eval_host(oh, ph, integer_starts=ints, out=outh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.eval_host(oh, ph, integer_starts=ints, out=outh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        stride = oh.stride(0)
        e2e = {"value": world * Be * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(Be * stride * 2), "d2h_bytes_per_step": int(Be * 4),
               "candidates_per_gpu_per_step": Be, "api": "saturn_b200.engine.Engine.eval_host -> sb_eval_host"}

    # ---- the reference-facing call itself: saturn.solver.solve(task_list) on host Task objects
    solve_leg = None
    if not args.no_e2e and args.config == "C4":
        from saturn_b200 import Strategy, solve
        from saturn_b200 import solver as sb_solver
        import numpy as _np

        class _Task:
            def __init__(self, name, strategies):
                self.name, self.strategies, self.selected_strategy = name, strategies, None

            def select_strategy(self, st_):
                self.selected_strategy = st_

        tmin_h = _np.where(valid, T, _np.inf).min(axis=1)
        tasks = [_Task("t%d" % j, {g + 1: Strategy("x", g + 1, {}, float(tmin_h[j, g])) for g in range(G)
                                    if _np.isfinite(tmin_h[j, g])}) for j in range(J)]
        solve(tasks, None, engine=eng, rounds=8)                  # warm-up
        barrier()
        t0 = time.perf_counter()
        plan = solve(tasks, None, engine=eng, rounds=200)
        dt = time.perf_counter() - t0
        stt = dict(sb_solver.last_stats)
        eng.set_table(T)                                                          # restore the bench table
        solve_leg = {"value": stt["candidates"] / dt, "unit": UNIT, "wall_s": dt, "candidates": stt["candidates"],
                     "makespan": plan[5], "h2d_bytes": int(J * 8 * 4), "d2h_bytes": int(J * (8 + 4 + 1 + 1 + 1)),
                     "api": "saturn.solver.solve(task_list) -> (sta, tga, bss, bna, boa, makespan); per rank"}

    # ---- one fused search round at a large population (diagnostic: the kernel the solver actually runs)
    search_leg = None
    if not args.no_e2e and args.config == "C4":
        wave = eng.search_wave(reduced=True)
        chains = wave * max(1, round((1 << 20) / wave))      # ~1 M chains in whole waves of the round kernel
        eng.search_init(chains, seed=rank, chain_base=rank * chains, integer_starts=ints, reduced=True,
                        t_start=5e-4, t_end=1e-6, total_rounds=64)
        eng.search_round(8)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        eng.search_round(32)
        s1.record()
        torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / 32
        search_leg = {"candidates_per_s_per_gpu": chains / (ms * 1e-3), "ms_per_round": ms, "chains_per_gpu": chains,
                      "fused": eng.search_is_fused(), "rounds_per_launch": 8,
                      "what": "the round kernel solve() runs (min-over-strategies table): move + evaluate + Metropolis "
                              "accept of every chain, 8 rounds per launch with the rows resident in shared memory"}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = B * bytes_per_candidate(J)
        achieved = alg / (kern_ms * 1e-3) / 1e9
        dram, _alg_ncu = ncu_traffic()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": dram if args.config == "C4" else None, "eval_path": kernel_path, "kernel": "k_eval_tiles<1,%s,true>" % ("true" if ints else "false"),
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                "note": "instruction-issue bound, not HBM bound: one list-scheduling step is ~53 SASS "
                        "instructions per warp of 32 candidates for 64 bytes of input (89 % of issue slots used); see "
                        "DESIGN.md 5.1 and profiles/r01_summary.md"}
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            threads = c_oracle.max_threads()
            rate, n, dt = cpu_eval_rate(threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d candidates of C4 (J=256,S=8,G=8) in %.1f s, oracle/ref_eval.c fp32 integer "
                             "starts, OpenMP over candidates" % (n, dt)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "candidates_per_gpu_per_step": B,
                           "integer_starts": ints,
                           "l2": "inputs (%.0f MB of encodings per GPU per step) exceed the 126 MB L2"
                                 % (B * 2 * opt.stride(0) / 1e6),
                           "exchange": ("none (N=1)" if world == 1 else
                                        ("one MIN of a uint64 per step over NVLink peer memory, fused into the "
                                         "evaluation kernel: publish in the tail, fold of the previous step's keys in "
                                         "the prologue (last step folded by a one-warp kernel inside the timed region)"
                                         if pipelined else
                                         "one MIN of a uint64 per step over NVLink peer memory (publish fused into the "
                                         "evaluation kernel, one-warp fold kernel)") if use_xchg else
                                        "one NCCL all_reduce(MIN) of a uint64 per step")},
                "clocks": clocks, "e2e": e2e, "solve_api": solve_leg, "search_round": search_leg, "gpu_launches": args.steps, "roofline": roof,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
