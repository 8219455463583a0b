"""GPU, >= 2 devices: the NVLink peer-memory MIN exchange and the sharded search, one process per GPU."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import ref_eval as R
        from saturn_b200.engine import Engine, random_candidates
        from saturn_b200.search import key_makespan, run_search
        eng = Engine(rank)
        J, S, G = 64, 6, 8
        T, valid = R.synth_table(J, S, G, seed=0)
        eng.set_table(T)
        assert eng.xchg_init(dist), "peer mappings could not be opened"
        dev = eng.device
        ok = True
        for rnd in range(6):
            B = 20000 + 1000 * rank
            opt, prio = random_candidates(eng, B, valid, seed=10 * rnd + rank)
            key = torch.full((1,), 2 ** 63 - 1, dtype=torch.int64, device=dev)
            gmin = torch.zeros(1, dtype=torch.int64, device=dev)
            out = eng.eval(opt, prio, best_key=key, id_base=rank * 1_000_000, post_key=(rnd % 2 == 0))
            local = key.clone()
            if rnd % 2 == 1:
                eng.xchg_post(key)                       # unfused post
            eng.xchg_reduce(gmin, fold=key)
            torch.cuda.synchronize()
            ref = local.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.MIN)   # NCCL as the checker
            ok = ok and int(gmin.item()) == int(ref.item()) == int(key.item())
            ok = ok and key_makespan(int(local.item())) == float(out.min())
        eng.xchg_check()
        # pipelined form: every launch publishes its round and folds the peers' previous round
        key = torch.full((1,), 2 ** 63 - 1, dtype=torch.int64, device=dev)
        locals_ = []
        for rnd in range(5):
            opt, prio = random_candidates(eng, 30000, valid, seed=100 + 10 * rnd + rank)
            one = torch.full((1,), 2 ** 63 - 1, dtype=torch.int64, device=dev)
            eng.eval(opt, prio, best_key=one, id_base=rank * 1_000_000)          # this round alone (checker)
            locals_.append(one)
            eng.eval(opt, prio, best_key=key, id_base=rank * 1_000_000, post_key=True, fold_prev=True)
        gmin = torch.zeros(1, dtype=torch.int64, device=dev)
        eng.xchg_reduce(gmin, fold=key)
        torch.cuda.synchronize()
        eng.xchg_check()
        ref = torch.stack(locals_).min().reshape(1)
        dist.all_reduce(ref, op=dist.ReduceOp.MIN)
        ok = ok and int(key.item()) == int(ref.item())
        res = run_search(eng, chains=4096, rounds=24, seed=5, use_dist=True)
        tab = R.canon_table(T, range(1, 9))
        mk = float(R.list_schedule(tab, res.opt, res.prio, True, np.float32)[0])
        q.put((rank, ok, res.makespan, mk, res.opt.tolist(), res.evaluated))
    except Exception as e:
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_peer_memory_exchange_and_sharded_search():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(len(o) == 6 for o in out), out
    (r0, ok0, mk0, ev0, o0, n0), (r1, ok1, mk1, ev1, o1, n1) = out
    assert ok0 and ok1
    assert mk0 == mk1 == ev0 == ev1 and o0 == o1 and n0 == n1
