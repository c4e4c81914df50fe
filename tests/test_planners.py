"""Product planners / plan vocabulary (easydist_b200.planners, .metair, .device_mesh) against the
fixtures generated from the reference (tests/golden/planners.json.gz). CPU only."""
import gzip
import json
import os

import numpy as np

from easydist_b200 import metair as M
from easydist_b200 import planners
from easydist_b200.device_mesh import DeviceMesh

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mk(t):
    return M.R() if t[0] == "R" else M.S(t[1]) if t[0] == "S" else M.P(t[1])


def _enc(steps):
    return [[i, list(a.key()), list(b.key())] for i, a, b in steps]


def test_planners_match_reference_fixture():
    with gzip.open(os.path.join(GOLDEN, "planners.json.gz"), "rt") as f:
        cases = json.load(f)
    for c in cases:
        src = M.VarSPMDStrategy(*[_mk(t) for t in c["src"]])
        dst = M.VarSPMDStrategy(*[_mk(t) for t in c["dst"]])
        assert _enc(planners.plan_greedy(src, dst)) == c["greedy"], (c["src"], c["dst"])
        assert _enc(planners.plan_replicate(src, dst)) == c["replicate"], (c["src"], c["dst"])
        steps, left = planners.plan_immediate(src, dst)
        assert _enc(steps) == c["immediate"], (c["src"], c["dst"])
        assert [list(s.key()) for s in left] == c["immediate_left"]


def test_step_kind_table():
    # sharding.py:739-793
    assert planners.step_kind(M.R(), M.S(0)) == "scatter"
    assert planners.step_kind(M.S(0), M.S(1)) == "all_to_all"
    assert planners.step_kind(M.S(1), M.S(1)) is None
    assert planners.step_kind(M.P("sum"), M.S(0)) == "reduce_scatter"
    assert planners.step_kind(M.S(2), M.R()) == "all_gather"
    assert planners.step_kind(M.P("avg"), M.R()) == "all_reduce"
    assert planners.step_kind(M.R(), M.P("sum")) is None


def test_spmd_equality_and_json_round_trip():
    assert M.S(1) == M.S(1) and M.S(1) != M.S(0) and M.P("sum") != M.P("max") and M.R() == M.R()
    plan = {"mm": {"node": "mm", "strategy": M.NodeSPMDStrategy(
        M.VarSPMDStrategyGroup(M.VarSPMDStrategy(M.S(0), M.R()), M.VarSPMDStrategy(M.R(), M.S(1))),
        M.VarSPMDStrategyGroup(M.VarSPMDStrategy(M.S(0), M.S(1)), None))}}
    back = M.plan_from_json(M.plan_to_json(plan))
    assert back["mm"]["strategy"].in_strtg_group == plan["mm"]["strategy"].in_strtg_group
    assert back["mm"]["strategy"].out_strtg_group[0] == M.VarSPMDStrategy(M.S(0), M.S(1))
    assert back["mm"]["strategy"].out_strtg_group[1] is None


def test_device_mesh_groups():
    mesh = np.arange(8).reshape(2, 2, 2)
    m = DeviceMesh(mesh, ["pp", "spmd0", "spmd1"], rank=5)
    assert m.get_coordinate() == [1, 0, 1]
    assert m.ranks_along(0) == [1, 5] and m.ranks_along(1) == [5, 7] and m.ranks_along(2) == [4, 5]
    assert m.spmd_dims() == [1, 2]
    sub = m.submesh([1, 2])
    assert sub.shape == (2, 2) and sub.get_coordinate() == [0, 1]
    one = DeviceMesh([0], ["spmd0"], rank=0)   # world size 1 is allowed here
    assert one.size(0) == 1 and one.ranks_along(0) == [0]


def test_product_partition_matches_reference_fixture():
    """easydist_b200.planners.Partition / recv_boxes == Partition.from_tensor_spec / gen_recv_meta
    of the reference (fixture generated from it)."""
    with gzip.open(os.path.join(GOLDEN, "partition.json.gz"), "rt") as f:
        cases = json.load(f)
    enc = lambda p: [list(p.start), list(p.end), p.rank, list(p.partial)]
    for c in cases[::7]:
        mesh = DeviceMesh(np.arange(int(np.prod(c["mesh"]))).reshape(c["mesh"]),
                          [f"spmd{i}" for i in range(len(c["mesh"]))], rank=0)
        src = M.VarSPMDStrategy(*[_mk(t) for t in c["src"]])
        dst = M.VarSPMDStrategy(*[_mk(t) for t in c["dst"]])
        sp = planners.partitions_from_spec(src, c["gshape"], mesh)
        tp = planners.partitions_from_spec(dst, c["gshape"], mesh)
        assert [enc(p) for p in sp] == c["src_parts"]
        assert [enc(p) for p in tp] == c["dst_parts"]
        for t in tp:
            got = [enc(p) for p in planners.recv_boxes(sp, t)]
            assert got == c["recv"].get(str(t.rank), []), (c["mesh"], c["src"], c["dst"], t.rank)
