"""CPU: the oracle against the reference's own outputs (golden fixtures) and against itself."""
import numpy as np
import pytest

from oracle import c_oracle, ref_eval as R


def probe_tuples(J, options, seed):
    import random
    random.seed(seed)
    out = []
    for _ in range(J):
        base = random.uniform(500, 4000)
        out.append([(g, base / g ** 0.8) for g in options])
    return out


def test_known_answers_from_survey():
    # values obtained in the survey session from the reference MILP / brute force
    tab, om = R.table_from_tuples(probe_tuples(3, [8], 1))
    assert R.brute_force(tab, om, True)[0] == pytest.approx(1442.211, abs=1e-3)
    assert R.brute_force(tab, om, False)[0] == pytest.approx(1441.731, abs=1e-3)
    mk, start, mask, _ = R.list_schedule(tab, [7, 7, 7], [0, 1, 2], True)
    assert [float(s) for s in start] == [0.0, 184.0, 841.0]      # integer starts: ceil(183.83), 184+657
    assert mask == [255, 255, 255]


def test_k2_bruteforce():
    tab, om = R.table_from_tuples(probe_tuples(5, [1, 2, 4, 8], 0))
    assert R.brute_force(tab, om, True)[0] == pytest.approx(1940.617, abs=1e-3)


@pytest.mark.parametrize("variant", ["tight_m"])
def test_oracle_matches_reference_milp(golden, variant):
    """For every instance the reference MILP (run unmodified, sound big-M) solved to proven
    optimality, the oracle's exhaustive list-scheduling minimum equals the MILP optimum; where the
    MILP only has an incumbent, the oracle's minimum is not worse."""
    n_opt = 0
    for rec in golden["cases"]:
        if rec["variant"] != variant or not rec["incumbent"]:
            continue
        assert rec["feasible"], rec["name"]
        tab, om = R.table_from_tuples([[tuple(x) for x in t] for t in rec["gpu_time_tuples"]])
        bf = R.brute_force(tab, om, True)[0]
        assert bf == pytest.approx(rec["bruteforce_int"]["makespan"], rel=1e-12)
        if rec["proven_optimal"]:
            assert bf == pytest.approx(rec["makespan"], rel=1e-9), rec["name"]
            n_opt += 1
        else:
            assert bf <= rec["makespan"] * (1 + 1e-9), rec["name"]
    assert n_opt >= 4


def test_reference_plan_reevaluates_to_same_makespan(golden):
    """Feed the MILP's own plan through the oracle: decode (bss -> option, start order -> priority),
    list-schedule it, and check the makespan is not worse than the MILP's (equal when optimal);
    the MILP's arrays satisfy the restated constraint set."""
    for rec in golden["cases"]:
        if rec["variant"] != "tight_m" or not rec["incumbent"]:
            continue
        tuples = [[tuple(x) for x in t] for t in rec["gpu_time_tuples"]]
        assert R.milp_constraints_hold(tuples, rec["sta"], rec["tga"], rec["bss"], rec["bna"], rec["boa"],
                                       rec["makespan"]) == [], rec["name"]
        plan = R.plan_from_arrays(tuples, rec["sta"], rec["tga"], rec["bss"], rec["bna"])
        tab, om = R.table_from_tuples(tuples)
        opt = [om[t][p[4]] for t, p in enumerate(plan)]
        prio = sorted(range(len(plan)), key=lambda t: (plan[t][0], t))
        mk, start, mask, _ = R.list_schedule(tab, opt, prio, True)
        assert mk <= rec["makespan"] * (1 + 1e-9), rec["name"]
        if rec["proven_optimal"]:
            assert mk == pytest.approx(rec["makespan"], rel=1e-9)


def test_as_shipped_big_m_leak_is_detected(golden):
    """The reference's M = 1e10 leaks under HiGHS: some as-shipped plans overlap and
    the independent checker must say so."""
    leaks = [r for r in golden["cases"] if r["variant"] == "as_shipped" and r["incumbent"] and r["overlaps"] > 0]
    assert len(leaks) >= 2
    for rec in leaks:
        assert not rec["feasible"]
        assert rec["makespan"] < rec["bruteforce_int"]["makespan"]


@pytest.mark.parametrize("J,S,G", [(4, 2, 2), (8, 3, 8), (64, 6, 8), (33, 1, 5)])
@pytest.mark.parametrize("ints", [True, False])
def test_c_port_equals_python(J, S, G, ints):
    T, valid = R.synth_table(J, S, G, seed=J)
    tab = R.canon_table(T, range(1, G + 1))
    opt, prio = R.synth_candidates(J, 300, valid, seed=1)
    for dt in (np.float32, np.float64):
        a = R.list_schedule_batch(tab, opt, prio, ints, dt, want_plan=True)
        b = c_oracle.evaluate(tab, opt, prio, ints, dt, want_plan=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        m1 = [R.list_schedule(tab, opt[b_], prio[b_], ints, dt)[0] for b_ in range(20)]
        assert np.array_equal(np.asarray(m1, dtype=dt), a[0][:20])


def test_fp32_vs_fp64_tolerance():
    """Integer-start mode: the fp32 arithmetic of the CUDA path is exact up to one rounding of the
    final completion time (<= 2^-24 relative).  Real-valued mode: fp32 accumulation stays within
    2e-6 relative of float64 on the headline shape."""
    T, valid = R.synth_table(256, 8, 8, seed=0)
    tab = R.canon_table(T, range(1, 9))
    opt, prio = R.synth_candidates(256, 5000, valid, seed=2)
    a = c_oracle.evaluate(tab, opt, prio, True, np.float32)
    b = c_oracle.evaluate(tab, opt, prio, True, np.float64)
    assert np.max(np.abs(a - b) / b) <= 2.0 ** -23
    a = c_oracle.evaluate(tab, opt, prio, False, np.float32)
    b = c_oracle.evaluate(tab, opt, prio, False, np.float64)
    assert np.max(np.abs(a - b) / b) <= 2e-6


def test_plan_checker_catches_overlap():
    ok, ov, mk = R.check_plan([0, 5], [0b11, 0b10], [10.0, 3.0], [2, 1])
    assert not ok and ov == 1
    ok, ov, mk = R.check_plan([0, 10], [0b11, 0b10], [10.0, 3.0], [2, 1])
    assert ok and mk == 13.0


def test_ragged_and_edge_tables():
    # absent options are +inf and make a candidate infeasible (inf makespan)
    tab, om = R.table_from_tuples([[(2, 10.0)], [(1, 5.0), (8, 1.0)]])
    assert R.list_schedule(tab, [0, 0], [0, 1], True)[0] == float("inf")   # job 0 has no 1-GPU option
    mk, start, mask, _ = R.list_schedule(tab, [om[0][0], om[1][1]], [1, 0], True)
    assert mk == 11.0 and mask == [0b11, 0xff] and [float(s) for s in start] == [1.0, 0.0]
    # ties go to the lowest slot index
    mk, start, mask, _ = R.list_schedule(tab, [om[0][0], om[1][0]], [1, 0], True)
    assert mask[1] == 0b1 and mask[0] == 0b110


def test_milp_port_matches_reference_runs(golden):
    """oracle/ref_milp.py (the MILP restated for scipy/HiGHS, used on the GPU box where the reference
    tree is absent) reproduces the optimum the UNMODIFIED reference reached on the fast instances."""
    from oracle import ref_milp
    done = 0
    for rec in golden["cases"]:
        if rec["variant"] != "tight_m" or rec["name"] not in ("K1_J3_g8_seed1", "C1_J4_g12_seed0"):
            continue
        tup = [[tuple(x) for x in t] for t in rec["gpu_time_tuples"]]
        r = ref_milp.solve(tup, time_limit=60)
        assert r["proven_optimal"] and r["makespan"] == pytest.approx(rec["makespan"], rel=1e-9)
        assert r["n_vars"] == rec["highs"]["n_vars"]
        rts = [tup[t][o][1] for t, o in enumerate(r["opt_idx"])]
        ks = [tup[t][o][0] for t, o in enumerate(r["opt_idx"])]
        assert R.check_plan(r["start"], r["mask"], rts, ks)[0]
        done += 1
    assert done == 2
    # SURVEY §8a A2/A3 model-size formulas: C1 (measured 89/616), J=5 S=4 (126/1850), C4
    assert ref_milp.model_size(4, 2) == (89, 616) and ref_milp.model_size(5, 4) == (126, 1850)
    assert ref_milp.model_size(256, 8) == (71681, 8413696)


def test_multi_node_oracle_matches_reference_milp(golden_n2):
    """Two nodes of 8 GPUs (the reference with ray.nodes() reporting 2 nodes, milp.py:58): the
    exhaustive multi-node list-scheduling optimum equals the reference MILP's proven optimum, and the
    reference's arrays satisfy the restated constraint set."""
    assert golden_n2["nodes"] == 2
    for rec in golden_n2["cases"]:
        assert rec["incumbent"] and rec["proven_optimal"] and rec["feasible"], rec["name"]
        tuples = [[tuple(x) for x in t] for t in rec["gpu_time_tuples"]]
        tab, om = R.table_from_tuples(tuples)
        bf = R.brute_force(tab, om, True, nodes=2)
        assert bf[0] == pytest.approx(rec["makespan"], rel=1e-9), rec["name"]
        assert bf[0] == pytest.approx(rec["bruteforce_int"]["makespan"], rel=1e-12)
        assert R.milp_constraints_hold(tuples, rec["sta"], rec["tga"], rec["bss"], rec["bna"], rec["boa"],
                                       rec["makespan"]) == []
        # one node would be strictly worse on these instances: the second node is really used
        assert R.brute_force(tab, om, True, nodes=1)[0] > bf[0]


def test_multi_node_c_port_equals_python():
    T, valid = R.synth_table(40, 1, 8, seed=3, masked=False)
    tab = R.canon_table(T, range(1, 9))
    rng = np.random.default_rng(0)
    B = 200
    opt = (rng.integers(0, 8, size=(B, 40)) | (rng.integers(0, 3, size=(B, 40)) << 3)).astype(np.uint8)
    prio = np.argsort(rng.random((B, 40)), axis=1).astype(np.uint8)
    for dt in (np.float32, np.float64):
        for ints in (True, False):
            mk, st, ms = c_oracle.evaluate(tab, opt, prio, ints, dt, want_plan=True, nodes=3)
            for b in range(0, B, 9):
                m1, s1, k1, _ = R.list_schedule(tab, opt[b], prio[b], ints, dt, nodes=3)
                assert m1 == mk[b] and list(ms[b]) == k1 and all(float(x) == float(y) for x, y in zip(s1, st[b]))
    bad = opt.copy()
    bad[0, 0] |= 3 << 3            # node 3 of 3 does not exist
    assert np.isinf(c_oracle.evaluate(tab, bad, prio, True, np.float32, nodes=3)[0])
