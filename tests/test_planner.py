"""CPU: the planner hook (kolibrie_b200/planner.py) — Streamertail's plan search mirrored for the hot path, with device operators and
device cost constants: the optimizer must CHOOSE the device plans for the BASELINE configs and keep the CPU operators for point
queries, must reproduce the reference's plan shapes when no device is offered, and the chosen plans must run (oracle-backed context)."""
import numpy as np
import pytest

from kolibrie_b200 import datagen
from kolibrie_b200 import engine as E
from kolibrie_b200 import planner as P
from tests import oracle_api as O
from tests.oracle_ctx import OracleCtx

V, C = E.Variable, E.Constant


def employee_stats(d, n_employees):
    return P.DatabaseStats.from_counts(6 * n_employees, {pid: n_employees for pid in (1, 2, 4, 6, 8, 10)})


def cfg_patterns(d):
    ids = d.ids
    title = (V("?e"), C(ids["foaf:title"]), V("?t"))
    salary = (V("?e"), C(ids["ds:annual_salary"]), V("?s"))
    name = (V("?e"), C(ids["foaf:name"]), V("?n"))
    work = (V("?e"), C(ids["foaf:workplaceHomepage"]), V("?c"))
    gt = E.Condition(E.Comparison("?s", ">", "100000"))
    return {"cfg1": ([work, name], None, None), "cfg2": ([title, salary, name], gt, ["?e", "?t", "?s", "?n"]), "cfg3": ([title, salary, name, work], None, None)}


@pytest.mark.parametrize("employees", [10_000, 1_666_667, 16_666_667])
def test_optimizer_chooses_device_plans_for_the_baseline_configs(employees):
    d = datagen.employee_dataset(1000)
    q = cfg_patterns(d)
    opt = P.Streamertail(employee_stats(d, employees), gpu=True)
    # cfg2: Projection(Filter(StarJoin)) in the reference (optimizer.rs:200-212) — the star becomes the device operator
    pats, cond, vars_ = q["cfg2"]
    plan = opt.find_best_plan(P.build_logical_plan(pats, cond, vars_))
    assert isinstance(plan, E.Projection) and isinstance(plan.input, E.Filter) and isinstance(plan.input.input, P.GpuStarJoin)
    assert plan.input.input.join_var == "?e" and len(plan.input.input.patterns) == 3
    # cfg3: a bare 4-pattern star
    pats, _, _ = q["cfg3"]
    plan = opt.find_best_plan(P.build_logical_plan(pats))
    assert isinstance(plan, P.GpuStarJoin) and len(plan.patterns) == 4
    # cfg1: two patterns are not a star (optimizer.rs:88 needs three): a join — the device bind join is the cheapest candidate
    pats, _, _ = q["cfg1"]
    plan = opt.find_best_plan(P.build_logical_plan(pats))
    assert isinstance(plan, (P.GpuBindJoin, P.GpuHashJoin)) and P.uses_device(plan)
    # and the device plan is cheaper than what the reference would have picked
    ref = P.Streamertail(employee_stats(d, employees), gpu=False).find_best_plan(P.build_logical_plan(pats))
    assert opt.cost.estimate_cost(plan) < opt.cost.estimate_cost(ref)


def test_without_a_device_the_reference_shapes_come_out():
    d = datagen.employee_dataset(1000)
    q = cfg_patterns(d)
    opt = P.Streamertail(employee_stats(d, 100_000), gpu=False)
    pats, cond, vars_ = q["cfg2"]
    plan = opt.find_best_plan(P.build_logical_plan(pats, cond, vars_))
    assert isinstance(plan, E.Projection) and isinstance(plan.input, E.Filter) and isinstance(plan.input.input, E.StarJoin)
    assert not P.uses_device(plan)
    plan = opt.find_best_plan(P.build_logical_plan(q["cfg1"][0]))
    assert isinstance(plan, (E.OptimizedHashJoin, E.HashJoin, E.ParallelJoin))
    # scans: two bound positions -> IndexScan; one bound with a large estimate -> TableScan (optimizer.rs:482-501)
    assert isinstance(opt.choose_best_scan((C(5), C(1), V("?o"))), E.IndexScan)
    assert isinstance(opt.choose_best_scan((V("?s"), C(1), V("?o"))), E.TableScan)


def test_point_queries_stay_on_the_cpu_operators():
    """a device operator costs a launch and a synchronisation (~25 us = 250 cost units): a lookup that touches a handful of rows is
    cheaper on the host index, and the cost model says so"""
    d = datagen.employee_dataset(4)
    st = P.DatabaseStats(zip(d.s.tolist(), d.p.tolist(), d.o.tolist()))
    opt = P.Streamertail(st, gpu=True)
    q = cfg_patterns(d)
    plan = opt.find_best_plan(P.build_logical_plan(q["cfg3"][0]))
    assert isinstance(plan, E.StarJoin) and not P.uses_device(plan)
    scan = opt.choose_best_scan((C(int(d.s[0])), C(d.ids["foaf:title"]), V("?t")))
    assert isinstance(scan, E.IndexScan)
    # estimator arithmetic against the reference's formulas (estimator.rs:43-62, 118-131)
    ce = opt.cost
    assert ce.estimate_cost(E.TableScan((V("?s"), C(d.ids["foaf:title"]), V("?o")))) == 4 * 100
    assert ce.estimate_cost(E.IndexScan((C(int(d.s[0])), C(d.ids["foaf:title"]), V("?o")))) == (min(6, 4) * 1) // 100
    assert ce.estimate_cost(E.StarJoin("?e", q["cfg3"][0])) == 4 + (4 + 4 + 4) // 10


def test_chosen_plans_run_and_agree_with_the_oracle():
    d = datagen.employee_dataset(3000)
    q = cfg_patterns(d)
    db = E.SparqlDatabase(ctx=OracleCtx())
    for i in range(d.n_ids):  # a dictionary whose ids are the generator's
        db.dictionary.encode(str(int(d.num_or0[i])) if d.is_num[i] else f"t{i}")
    for t in zip(d.s.tolist(), d.p.tolist(), d.o.tolist()):
        db.add_triple(t)
    odb = O.Db(d.s, d.p, d.o, d.num_or0, d.is_num)
    js, opats, ofilt = datagen.employee_queries(d)["cfg2"]
    want = odb.bgp(opats, ofilt).n_rows
    for gpu in (True, False):
        opt = P.Streamertail(P.DatabaseStats(db.triples), gpu=gpu)
        pats, cond, vars_ = q["cfg2"]
        plan = opt.find_best_plan(P.build_logical_plan(pats, cond, vars_))
        rows = E.ExecutionEngine.execute_with_ids(P.lower_to_engine(plan), db)
        assert len(rows) == want and set(rows[0]) == {"e", "t", "s", "n"}
        plan1 = opt.find_best_plan(P.build_logical_plan(q["cfg1"][0]))
        rows1 = E.ExecutionEngine.execute_with_ids(P.lower_to_engine(plan1), db)
        assert len(rows1) == 3000
