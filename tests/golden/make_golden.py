"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/refcompat) in the CPU container.

    python tests/golden/make_golden.py            # parts A (single process) + B (gloo, 2 & 4 ranks)

The fixtures pin oracle/reshard_oracle.py (tests/test_oracle_golden.py) and, through the oracle,
the CUDA kernels.  /root/reference does not exist on the GPU box; only the committed fixtures
travel.  Fixture contents:
  planners.json.gz     _gen_transform_infos{,_greedy} / _gen_immediate_transform_infos outputs
  partition.json.gz    Partition.from_tensor_spec / gen_recv_meta boxes
  combination.npz     CombinationFunc.{gather,reduce,identity} + halo_padding outputs
  reshard_w{2,4}.npz  inputs/outputs of the reference's ten reshard callables under gloo
"""
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def _spmd_to_tuple(s):
    if s.is_replicate():
        return ["R"]
    if s.is_shard():
        return ["S", int(s.args["dim"])]
    return ["P", {1: "sum", 2: "max", 3: "min", 4: "avg"}[s.args["ops"].value]]


def _mk_spmd(t):
    from easydist.metashard.combination import ReduceOp
    from easydist.metashard.metair import SPMD
    if t[0] == "R":
        return SPMD(SPMD.REPLICATE)
    if t[0] == "S":
        return SPMD(SPMD.SHARD, {"dim": t[1]})
    return SPMD(SPMD.PARTIAL, {"ops": {"sum": ReduceOp.SUM, "max": ReduceOp.MAX,
                                       "min": ReduceOp.MIN, "avg": ReduceOp.AVG}[t[1]]})


PLACEMENT_ALPHABET = [["R"], ["S", 0], ["S", 1], ["P", "sum"]]


def part_a():
    from oracle import refcompat
    refcompat.install()
    import torch
    import easydist.torch.passes.sharding as sh
    from easydist.metashard.combination import CombinationFunc, ReduceOp
    from easydist.metashard.halo import HaloInfo, halo_padding
    from easydist.metashard.metair import VarSPMDStrategy
    import easydist.platform as platform
    platform.init_backend("torch")

    # ---- planners -------------------------------------------------------------------------
    cases = []
    for ndim in (1, 2, 3):
        alphabet = PLACEMENT_ALPHABET if ndim < 3 else PLACEMENT_ALPHABET[:3]
        combos = list(itertools.product(alphabet, repeat=ndim))
        for src in combos:
            for dst in combos:
                if any(d[0] == "P" for d in dst):
                    continue  # targets are never partial (sharding.py:739-793)
                vs = VarSPMDStrategy(*[_mk_spmd(t) for t in src])
                vd = VarSPMDStrategy(*[_mk_spmd(t) for t in dst])
                greedy = sh._gen_transform_infos_greedy(vs, vd)
                repl = sh._gen_transform_infos(list(vs), list(vd))
                imm, left = sh._gen_immediate_transform_infos(vs, vd)
                enc = lambda infos: [[int(i), _spmd_to_tuple(a), _spmd_to_tuple(b)]
                                     for i, a, b in infos]
                cases.append({"src": list(src), "dst": list(dst), "greedy": enc(greedy),
                              "replicate": enc(repl), "immediate": enc(imm),
                              "immediate_left": [_spmd_to_tuple(s) for s in left]})
    import gzip
    with gzip.open(os.path.join(HERE, "planners.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("planners.json.gz", len(cases))

    # ---- Partition -------------------------------------------------------------------------
    class _Mesh:
        def __init__(self, mesh):
            self.mesh = mesh
    pcases = []
    for mesh_shape in [(2,), (4,), (2, 2), (2, 4), (2, 2, 2)]:
        mesh = torch.arange(int(np.prod(mesh_shape))).reshape(mesh_shape)
        orig = sh.get_device_mesh
        sh.get_device_mesh = lambda *a, _m=mesh: _Mesh(_m)
        try:
            nd = len(mesh_shape)
            alphabet = [["R"], ["S", 0], ["S", 1]]
            combos = list(itertools.product(alphabet, repeat=nd))
            for gshape in [(8, 12), (7, 10), (16, 3)]:
                for src in combos:
                    for dst in combos:
                        vs = VarSPMDStrategy(*[_mk_spmd(t) for t in src])
                        vd = VarSPMDStrategy(*[_mk_spmd(t) for t in dst])
                        sp = sh.Partition.from_tensor_spec(vs, torch.Size(gshape))
                        tp = sh.Partition.from_tensor_spec(vd, torch.Size(gshape))
                        recv = sh.Partition.gen_recv_meta(sp, tp)
                        encp = lambda p: [list(p.start_coord), list(p.end_coord), int(p.rank),
                                          list(p.partial_coord)]
                        pcases.append({
                            "mesh": list(mesh_shape), "gshape": list(gshape), "src": list(src),
                            "dst": list(dst), "src_parts": [encp(p) for p in sp],
                            "dst_parts": [encp(p) for p in tp],
                            "recv": {str(k): [encp(p) for p in v] for k, v in recv.items()}})
        finally:
            sh.get_device_mesh = orig
    import gzip
    with gzip.open(os.path.join(HERE, "partition.json.gz"), "wt") as f:
        json.dump(pcases, f, separators=(",", ":"))
    print("partition.json.gz", len(pcases))

    # ---- combination / halo -----------------------------------------------------------------
    rng = np.random.RandomState(7)
    out = {}
    k = 0
    for n in (2, 3, 4):
        for shape in [(6,), (4, 6), (2, 6, 4)]:
            for dim in range(len(shape)):
                shards = [rng.randint(-9, 9, size=shape).astype(np.float32) for _ in range(n)]
                ts = [torch.from_numpy(s) for s in shards]
                for hw in (0, 1, 2, -1):
                    if hw != 0 and abs(hw) * 2 >= shape[dim]:
                        continue
                    for ch in ((1, 2, 3) if hw == 0 else (1,)):
                        if shape[dim] % ch:
                            continue
                        g = CombinationFunc.gather(ts, dim=dim, halowidth=hw, chunk=ch)
                        out[f"gather_{k}_in"] = np.stack(shards)
                        out[f"gather_{k}_meta"] = np.array([dim, hw, ch])
                        out[f"gather_{k}_out"] = g.numpy()
                        k += 1
                for halo in (1, 2):
                    if halo > shape[dim]:
                        continue
                    padded = halo_padding(ts, HaloInfo(halo, dim))
                    out[f"halo_{k}_in"] = np.stack(shards)
                    out[f"halo_{k}_meta"] = np.array([dim, halo])
                    for i, p in enumerate(padded):
                        out[f"halo_{k}_out{i}"] = p.numpy()
                    k += 1
                for opname, op in (("sum", ReduceOp.SUM), ("max", ReduceOp.MAX),
                                   ("min", ReduceOp.MIN), ("avg", ReduceOp.AVG)):
                    r = CombinationFunc.reduce(ts, ops=op)
                    out[f"reduce_{k}_in"] = np.stack(shards)
                    out[f"reduce_{k}_op"] = np.array(opname)
                    out[f"reduce_{k}_out"] = np.asarray(r.numpy() if hasattr(r, "numpy") else r)
                    k += 1
    np.savez_compressed(os.path.join(HERE, "combination.npz"), **out)
    print("combination.npz", k)


# cases for part B: (name, op, kwargs, shape of the per-rank input, dtype)
def reshard_cases(world):
    cases = []
    f32, i64, f16 = "float32", "int64", "float16"
    for dt in (f32, i64):
        cases += [("ag_d0", "all_gather", {"dim": 0}, (3, 4), dt),
                  ("ag_d1", "all_gather", {"dim": 1}, (3, 4), dt),
                  ("ag_d2", "all_gather", {"dim": 2}, (2, 3, 5), dt),
                  ("ag_vec", "all_gather", {"dim": 0}, (7,), dt),
                  ("a2a_01", "all_to_all", {"g": 0, "s": 1}, (2, 4 * world), dt),
                  ("a2a_10", "all_to_all", {"g": 1, "s": 0}, (2 * world, 3), dt),
                  ("a2a_02", "all_to_all", {"g": 0, "s": 2}, (2, 3, 2 * world), dt),
                  ("a2a_21", "all_to_all", {"g": 2, "s": 1}, (2, world, 5), dt),
                  ("sc_d0", "scatter", {"dim": 0}, (world * 2, 3), dt),
                  ("sc_d1", "scatter", {"dim": 1}, (3, world * 2), dt),
                  ("sc_ragged", "scatter", {"dim": 0}, (2 * world - 1, 2), dt)]
    for dt in (f32, i64, f16):
        for op in ("sum", "max", "min") + (("avg",) if dt != i64 else ()):
            if dt == i64 and op in ("max", "min"):
                pass
            cases += [(f"ar_{op}", "all_reduce", {"op": op}, (5, 3), dt),
                      (f"ar_{op}_scalar", "all_reduce", {"op": op}, (), dt),
                      (f"rs_{op}_d0", "reduce_scatter", {"op": op, "dim": 0}, (2 * world, 3), dt),
                      (f"rs_{op}_d1", "reduce_scatter", {"op": op, "dim": 1}, (3, 2 * world), dt),
                      (f"rs_{op}_d2", "reduce_scatter", {"op": op, "dim": 2}, (2, 3, world), dt)]
    return cases


def case_input(name, shape, dtype, rank, world):
    import zlib
    rng = np.random.RandomState(zlib.crc32(f"{name}|{dtype}".encode()) % (2 ** 31 - 100000) + 1000 * rank)
    if dtype == "int64":
        return rng.randint(-8, 9, size=shape).astype(np.int64)
    # integer-valued floats: sums are exact in any order => bit-exact fixtures
    return rng.randint(-8, 9, size=shape).astype(dtype)


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import refcompat
    refcompat.install()
    import easydist.torch.passes.sharding as sh
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    group = list(range(world))
    results = {}
    for name, op, kw, shape, dtype in reshard_cases(world):
        key = f"{name}_{dtype}"
        x_np = case_input(key, shape, dtype, rank, world)
        x = torch.from_numpy(np.ascontiguousarray(x_np))
        if op == "all_gather":
            y = sh.all_gather_end(sh.all_gather_start(x, kw["dim"], group), kw["dim"], group)
        elif op == "all_to_all":
            a = (kw["g"], kw["s"], world, rank, group)
            y = sh.all_to_all_end(sh.all_to_all_start(x, *a), *a)
        elif op == "scatter":
            y = sh.scatter_wrapper(x, world, kw["dim"], rank)
        elif op == "all_reduce":
            if kw["op"] == "avg":
                # gloo has no AVG; NCCL's avg is sum * (1/n): emulate on the summed result
                y = sh.all_reduce_end(sh.all_reduce_start(x, "sum", group), "sum", group)
                y = (y.float() * (1.0 / world)).to(x.dtype)
            else:
                y = sh.all_reduce_end(sh.all_reduce_start(x, kw["op"], group), kw["op"], group)
        elif op == "reduce_scatter":
            o = "sum" if kw["op"] == "avg" else kw["op"]
            y = sh.reduce_scatter_end(sh.reduce_scatter_start(x, o, kw["dim"], group), o,
                                      kw["dim"], group)
            if kw["op"] == "avg":
                y = (y.float() * (1.0 / world)).to(x.dtype)
        results[key] = (x_np, y.numpy().copy())
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        out = {}
        for r, res in enumerate(gathered):
            for key, (xi, yo) in res.items():
                out[f"{key}__in{r}"] = xi
                out[f"{key}__out{r}"] = yo
        np.savez_compressed(out_path, **out)
        print(os.path.basename(out_path), len(out))
    dist.barrier()
    dist.destroy_process_group()


def part_b():
    import torch.multiprocessing as mp
    for world, port in ((2, 29611), (4, 29612)):
        out_path = os.path.join(HERE, f"reshard_w{world}.npz")
        mp.spawn(_worker, args=(world, port, out_path), nprocs=world, join=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["a", "b"]
    if "a" in which:
        part_a()
    if "b" in which:
        part_b()
