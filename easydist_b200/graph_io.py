"""Serialise a traced FX graph + sharding plan to JSON and back.

The reference caches solved strategies per input signature (`ENABLE_COMPILE_CACHE`,
easydist/torch/compile_auto.py:97-106, 181-184: a pickled plan).  Here graph AND plan are stored
as plain JSON so that a plan solved where the reference's solver runs (CPU box) can be lowered
and executed where only this backend exists (GPU box): `parallel_mode="auto"` with
`plan=load_bundle(path)`.

FX GraphModules produced by make_fx are not picklable (pybind objects in metas), so nodes are
written out explicitly: (name, op, target, args, kwargs) with a small tagged encoding for Node
references, dtypes, devices, memory formats and layouts.  Devices are re-targeted on load.
"""
import json
import operator

import torch
import torch.utils._pytree as pytree
from torch.fx.graph import _PyTreeCodeGen, _PyTreeInfo
from torch.fx.node import Node

from . import metair


def _enc(a):
    if isinstance(a, Node):
        return {"$n": a.name}
    if isinstance(a, (list, tuple)):
        return {"$l" if isinstance(a, list) else "$t": [_enc(x) for x in a]}
    if isinstance(a, dict):
        return {"$d": {k: _enc(v) for k, v in a.items()}}
    if isinstance(a, torch.dtype):
        return {"$dtype": str(a).split(".")[-1]}
    if isinstance(a, torch.device):
        return {"$device": a.type}
    if isinstance(a, torch.memory_format):
        return {"$mf": str(a).split(".")[-1]}
    if isinstance(a, torch.layout):
        return {"$layout": str(a).split(".")[-1]}
    if a is None or isinstance(a, (bool, int, float, str)):
        return a
    if isinstance(a, torch.Size):
        return {"$l": [int(x) for x in a]}
    raise TypeError(f"graph_io: cannot serialise argument {a!r} of type {type(a)}")


def _dec(a, env, device):
    if isinstance(a, dict):
        if "$n" in a:
            return env[a["$n"]]
        if "$l" in a:
            return [_dec(x, env, device) for x in a["$l"]]
        if "$t" in a:
            return tuple(_dec(x, env, device) for x in a["$t"])
        if "$d" in a:
            return {k: _dec(v, env, device) for k, v in a["$d"].items()}
        if "$dtype" in a:
            return getattr(torch, a["$dtype"])
        if "$device" in a:
            return torch.device(device)
        if "$mf" in a:
            return getattr(torch, a["$mf"])
        if "$layout" in a:
            return getattr(torch, a["$layout"])
        raise ValueError(f"graph_io: unknown tag in {a}")
    return a


def _target_name(t):
    if t is operator.getitem:
        return "getitem"
    if isinstance(t, torch._ops.OpOverload):
        return "op:" + t._schema.name.replace("::", ".") + "." + t._overloadname
    if getattr(t, "__name__", "") == "md_embedding":
        # the reference's discovery-time stand-in for aten.embedding (passes/fix_embedding.py:19-23:
        # the same op behind an index range check; its own lowering maps it back, :33-34)
        return "op:aten.embedding.default"
    raise TypeError(f"graph_io: cannot serialise call target {t!r}")


def _resolve_target(name):
    if name == "getitem":
        return operator.getitem
    assert name.startswith("op:"), name
    ns, op, overload = name[3:].split(".")
    return getattr(getattr(getattr(torch.ops, ns), op), overload)


def dump_graph(gm):
    nodes = []
    for n in gm.graph.nodes:
        entry = {"name": n.name, "op": n.op}
        if n.op == "call_function":
            entry["target"] = _target_name(n.target)
            entry["args"] = _enc(tuple(n.args))
            entry["kwargs"] = _enc(dict(n.kwargs))
        elif n.op == "output":
            entry["args"] = _enc(tuple(n.args))
        elif n.op == "placeholder":
            entry["target"] = n.target
        else:
            raise TypeError(f"graph_io: unsupported node op {n.op}")
        nodes.append(entry)
    return {"nodes": nodes,
            "in_spec": pytree.treespec_dumps(gm._in_spec),
            "out_spec": pytree.treespec_dumps(gm._out_spec)}


def load_graph(d, device="cuda"):
    graph = torch.fx.Graph()
    env = {}
    for e in d["nodes"]:
        if e["op"] == "placeholder":
            env[e["name"]] = graph.placeholder(e["target"])
        elif e["op"] == "call_function":
            env[e["name"]] = graph.call_function(_resolve_target(e["target"]),
                                                 _dec(e["args"], env, device),
                                                 _dec(e["kwargs"], env, device))
        else:
            graph.output(*_dec(e["args"], env, device))
            continue
        env[e["name"]].name = e["name"]
    in_spec = pytree.treespec_loads(d["in_spec"])
    out_spec = pytree.treespec_loads(d["out_spec"])
    graph._codegen = _PyTreeCodeGen(_PyTreeInfo([f"arg{i}" for i in range(in_spec.num_children)],
                                                 in_spec, out_spec))
    gm = torch.fx.GraphModule(torch.nn.Module(), graph)
    # node names must survive exactly: plans are keyed by them
    for n, e in zip(gm.graph.nodes, d["nodes"]):
        assert n.name == e["name"], (n.name, e["name"])
    return gm


def dump_bundle(gm, plan, state_io_names, extra=None):
    """JSON text holding the traced graph, the plan (easydist_b200.metair vocabulary or the
    reference's objects) and the state in->out name map."""
    if plan and not isinstance(next(iter(plan.values()))["strategy"], metair.NodeSPMDStrategy):
        plan = metair.plan_from_reference(plan)
    return json.dumps({"graph": dump_graph(gm), "plan": json.loads(metair.plan_to_json(plan)),
                       "state_io": state_io_names, "extra": extra or {}}, separators=(",", ":"))


def load_bundle(text, device="cuda"):
    raw = json.loads(text)
    gm = load_graph(raw["graph"], device)
    plan = metair.plan_from_json(json.dumps(raw["plan"]))
    return gm, plan, raw["state_io"], raw.get("extra", {})
