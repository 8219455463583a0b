"""Plan a synthetic multi-model workload with the B200 solver through Saturn's own API.

    python examples/plan_synthetic.py [--jobs 32] [--nodes 1]

Stands in for the reference's examples/wikitext103/WikiText103.py after the trial runner has
filled `task.strategies` (saturn/trial_runner/PerformanceEvaluator.py:96-115): here the runtimes are
synthetic.  Everything below the `--- Saturn API ---` line is what a Saturn user already writes.
"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from saturn.core.representations import HParams, Strategy, Task  # noqa: E402
from saturn.solver import convert_into_comprehensible, solve      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=32)
    ap.add_argument("--nodes", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    save_dir = tempfile.mkdtemp(prefix="saturn_b200_")
    tasks = []
    for j in range(args.jobs):
        t = Task(get_model=lambda: None, get_dataloader=lambda: range(100), loss_function=None,
                 hparams=HParams(lr=1e-4, epochs=1), name="model-%02d" % j, save_dir=save_dir)
        base = float(np.exp(rng.uniform(np.log(600), np.log(36000))))
        for g in range(1, 9):                       # what trial_runner.search() leaves behind
            if g == 1:
                t.strategies[g] = Strategy("spilled", g, {}, base * 1.3)
            elif rng.uniform() < 0.1:
                t.strategies[g] = Strategy(None, g, None, 1e8)       # every executor failed at this size
            else:
                t.strategies[g] = Strategy(rng.choice(["fsdp", "pipeline"]), g, {}, base / g ** rng.uniform(0.6, 0.95))
        tasks.append(t)

    # --- Saturn API -------------------------------------------------------------------------
    sta, tga, bss, bna, boa, makespan = solve(tasks, nodes=args.nodes)
    node_per_task, deps, starts = convert_into_comprehensible(tasks, bss, boa, tga, bna, sta)

    print("planned %d tasks on %d node(s): makespan %.0f s" % (len(tasks), args.nodes, makespan))
    for t, s in sorted(zip(tasks, starts), key=lambda x: x[1])[:12]:
        st = t.selected_strategy
        print("  %-9s node %d  start %8.0f  %d GPU(s)  %-8s  %8.0f s  after %s" % (
            t.name, node_per_task[t], s, st.gpu_apportionment, st.executor, st.runtime,
            [d.name for d in deps[t]][:3] if t in deps else []))


if __name__ == "__main__":
    main()
