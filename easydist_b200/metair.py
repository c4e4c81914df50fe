"""Plan vocabulary of the path — a mirror of the types the reference's solver emits
(easydist/metashard/metair.py:29-156 SPMD / VarSPMDStrategy / VarSPMDStrategyGroup /
NodeSPMDStrategy, ReduceOp from metashard/combination.py:33-37), so plans can be stored, loaded
and lowered where the reference itself is not importable (the GPU box).

`from_reference(obj)` converts the reference's own objects by duck typing (`.state`, `.args`,
`.var_spmd_strategy`, `.in_strtg_group` ...), so a plan produced by the unmodified AutoFlow solver
drops straight in.
"""
import json
from enum import Enum


class ReduceOp(Enum):
    SUM = 1
    MAX = 2
    MIN = 3
    AVG = 4


REDUCE_NAME = {ReduceOp.SUM: "sum", ReduceOp.MAX: "max", ReduceOp.MIN: "min", ReduceOp.AVG: "avg"}
_NAME_REDUCE = {v: k for k, v in REDUCE_NAME.items()}


def reduce_name(op):
    """'sum'/'max'/'min'/'avg' for ours or the reference's ReduceOp (reduce_map, sharding.py:68-73)."""
    if isinstance(op, str):
        return op
    return REDUCE_NAME[ReduceOp(op.value)]


class SPMD:
    REPLICATE = "REPLICATE"
    SHARD = "SHARD"
    PARTIAL = "PARTIAL"

    __slots__ = ("state", "args")

    def __init__(self, state, args=None):
        self.state = state
        self.args = args

    def is_shard(self):
        return self.state == SPMD.SHARD

    def is_replicate(self):
        return self.state == SPMD.REPLICATE

    def is_partial(self):
        return self.state == SPMD.PARTIAL

    @property
    def dim(self):
        return self.args["dim"]

    @property
    def op(self):
        return reduce_name(self.args["ops"])

    def __eq__(self, other):
        if other is None or self.state != other.state:
            return False
        if self.is_shard():
            return self.args["dim"] == other.args["dim"]
        if self.is_partial():
            return reduce_name(self.args["ops"]) == reduce_name(other.args["ops"])
        return True

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.key())

    def key(self):
        if self.is_shard():
            return ("S", self.args["dim"])
        if self.is_partial():
            return ("P", self.op)
        return ("R",)

    def __repr__(self):
        k = self.key()
        return "R" if k == ("R",) else f"{k[0]}({k[1]})"


def R():
    return SPMD(SPMD.REPLICATE)


def S(dim):
    return SPMD(SPMD.SHARD, {"dim": int(dim)})


def P(op="sum"):
    return SPMD(SPMD.PARTIAL, {"ops": _NAME_REDUCE[op] if isinstance(op, str) else op})


class VarSPMDStrategy:
    """One SPMD per mesh dim."""

    def __init__(self, *var_spmd_strategy):
        self.var_spmd_strategy = list(var_spmd_strategy)

    def __getitem__(self, idx):
        return self.var_spmd_strategy[idx]

    def __add__(self, other):
        return VarSPMDStrategy(*self.var_spmd_strategy, *other.var_spmd_strategy)

    def __eq__(self, other):
        if other is None or len(self) != len(other):
            return False
        return all(a == b for a, b in zip(self, other))

    def __ne__(self, other):
        return not self.__eq__(other)

    def __len__(self):
        return len(self.var_spmd_strategy)

    def __iter__(self):
        return iter(self.var_spmd_strategy)

    def __repr__(self):
        return f"VarSPMDStrategy({self.var_spmd_strategy})"

    def is_replicated(self):
        return all(s.is_replicate() for s in self)


class VarSPMDStrategyGroup:
    def __init__(self, *group):
        self.var_spmd_strategy_group = list(group)

    def append(self, s):
        self.var_spmd_strategy_group.append(s)

    def get_var_strtg(self, idx):
        return self.var_spmd_strategy_group[idx]

    def __getitem__(self, idx):
        return self.var_spmd_strategy_group[idx]

    def __setitem__(self, idx, v):
        self.var_spmd_strategy_group[idx] = v

    def __len__(self):
        return len(self.var_spmd_strategy_group)

    def __iter__(self):
        return iter(self.var_spmd_strategy_group)

    def __eq__(self, other):
        return len(self) == len(other) and all(a == b for a, b in zip(self, other))

    def __repr__(self):
        return f"VarSPMDStrategyGroup({self.var_spmd_strategy_group})"


class NodeSPMDStrategy:
    def __init__(self, in_strtg_group, out_strtg_group):
        self.in_strtg_group = in_strtg_group
        self.out_strtg_group = out_strtg_group

    def get_invar_strtg(self, idx):
        return self.in_strtg_group.get_var_strtg(idx)

    def get_outvar_strtg(self, idx):
        return self.out_strtg_group.get_var_strtg(idx)

    def __repr__(self):
        return f"NodeSPMDStrategy(in: {self.in_strtg_group}, out: {self.out_strtg_group})"


def replicate_strategy(mesh_ndim):
    return VarSPMDStrategy(*[R() for _ in range(mesh_ndim)])


# ---- conversion from the reference's objects (duck typed) and (de)serialisation ---------------------


def spmd_from_reference(s):
    if s is None:
        return None
    if s.state == SPMD.SHARD:
        return S(s.args["dim"])
    if s.state == SPMD.PARTIAL:
        return P(reduce_name(s.args["ops"]))
    return R()


def var_from_reference(v):
    if v is None:
        return None
    return VarSPMDStrategy(*[spmd_from_reference(s) for s in v])


def group_from_reference(g):
    return VarSPMDStrategyGroup(*[var_from_reference(v) for v in g])


def node_strategy_from_reference(ns):
    return NodeSPMDStrategy(group_from_reference(ns.in_strtg_group),
                            group_from_reference(ns.out_strtg_group))


def plan_from_reference(opt_strategy):
    """{node_name: {'node':…, 'strategy': NodeSPMDStrategy}} (solver.py:732-745) -> ours."""
    return {name: {"node": name, "strategy": node_strategy_from_reference(v["strategy"])}
            for name, v in opt_strategy.items()}


def _enc_var(v):
    return None if v is None else [list(s.key()) for s in v]


def _dec_var(v):
    if v is None:
        return None
    out = []
    for k in v:
        out.append(R() if k[0] == "R" else S(k[1]) if k[0] == "S" else P(k[1]))
    return VarSPMDStrategy(*out)


def plan_to_json(plan):
    return json.dumps({name: {"in": [_enc_var(v) for v in e["strategy"].in_strtg_group],
                              "out": [_enc_var(v) for v in e["strategy"].out_strtg_group]}
                       for name, e in plan.items()}, separators=(",", ":"))


def plan_from_json(text):
    raw = json.loads(text)
    return {name: {"node": name,
                   "strategy": NodeSPMDStrategy(
                       VarSPMDStrategyGroup(*[_dec_var(v) for v in e["in"]]),
                       VarSPMDStrategyGroup(*[_dec_var(v) for v in e["out"]]))}
            for name, e in raw.items()}
