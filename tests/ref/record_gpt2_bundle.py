"""Record an auto-SPMD plan bundle for the token-level GPT-2 workload (CPU container only; needs
/root/reference): the UNMODIFIED reference traces, annotates and solves (AutoFlow ILP through
oracle/refcompat), `compile_auto.sharding_transform` is rebound to this repo's lowering (the
drop-in hook of INTEGRATION.md §3), the result is checked against vanilla PyTorch for two steps
and graph + plan are written as a bundle the GPU box can lower without the reference
(`api.compile_from_bundle`).

STATUS: compiles, then fails at run time with "Target ... is out of bounds" — with the reference's
own lowering too (EDB_NO_HOOK=1): its discovery materialises integer inputs with randint(high=8)
(init_helper.py:57-63), so sharding the class dimension of nll_loss_forward looks legal to it
(DESIGN.md section 4).  Kept as the reproducer; the embedding-input GPT of the reference's own
example goes through tests/ref/auto_worker.py instead.

  EDB_GPT2=gpt2-tiny EDB_BATCH=4 EDB_SEQ=32 EDB_RECORD=tests/golden/auto_gpt2_tiny_mesh2.json \\
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ref/record_gpt2_bundle.py
"""
import copy
import dataclasses
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    mesh_shape = tuple(int(v) for v in os.environ.get("EDB_TEST_MESH", str(world)).split("x"))
    record = os.environ.get("EDB_RECORD", "")
    name = os.environ.get("EDB_GPT2", "gpt2-tiny")
    batch, seq = int(os.environ.get("EDB_BATCH", "4")), int(os.environ.get("EDB_SEQ", "32"))
    run_steps = int(os.environ.get("EDB_STEPS", "2"))
    torch.set_num_threads(int(os.environ.get("EDB_THREADS", "1")))
    dist.init_process_group("gloo")
    from oracle import refcompat
    refcompat.install()
    from easydist import easydist_setup
    from easydist.torch.api import easydist_compile
    from easydist.torch.device_mesh import set_device_mesh
    import easydist.torch.compile_auto as ref_auto
    from torch.distributed.device_mesh import DeviceMesh
    easydist_setup(backend="torch", device="cpu", allow_tf32=False)
    names = [f"spmd{i}" for i in range(len(mesh_shape))]
    set_device_mesh(DeviceMesh("cpu", torch.arange(world).reshape(mesh_shape), mesh_dim_names=names))

    from easydist_b200 import graph_io, lowering, workloads
    from easydist_b200.device_mesh import set_device_mesh as edb_set_mesh
    from tests import gloo_ops
    my_mesh = edb_set_mesh(torch.arange(world).reshape(mesh_shape).numpy(), names, rank=rank)
    gloo_ops.init_groups(my_mesh.mesh)

    cfg = dataclasses.replace(workloads.GPT2_CONFIGS[name], attn="unfused", pos_as_buffer=True,
                              block_size=seq)  # == seq: no slice of the position buffer in the graph
    torch.manual_seed(42)
    model0 = workloads.GPT2(cfg)
    data = [workloads.synthetic_tokens(cfg, batch, seq, seed=100 + i) for i in range(run_steps)]
    saved = {}
    orig = ref_auto.sharding_transform

    def mine(fx_module, opt_strategy, state_io_map):
        if record and rank == 0:
            saved["plan"] = graph_io.dump_bundle(
                fx_module, opt_strategy, [(a.name, b.name) for a, b in state_io_map.items()],
                extra={"mesh": list(mesh_shape), "model": name, "attn": "unfused", "seed": 42,
                       "batch": [batch, seq], "optimizer": "SGD(lr=0.1, momentum=0.9, foreach)"})
        return lowering.sharding_transform(fx_module, opt_strategy, state_io_map, ops=gloo_ops,
                                           mesh=my_mesh, planner="GREEDY")

    if os.environ.get("EDB_NO_HOOK") != "1":  # EDB_NO_HOOK=1: the pure reference, for comparison
        ref_auto.sharding_transform = mine
    model = copy.deepcopy(model0)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, foreach=True)
    t0 = time.time()
    try:
        step = easydist_compile(workloads.gpt2_train_step, "auto", "fake", cuda_graph=False)
        outs = [step(tok, tgt, model, opt).detach().clone() for tok, tgt in data]
    finally:
        ref_auto.sharding_transform = orig
    t1 = time.time()
    vmodel = copy.deepcopy(model0)
    vopt = torch.optim.SGD(vmodel.parameters(), lr=0.1, momentum=0.9, foreach=True)
    vouts = [workloads.gpt2_train_step(tok, tgt, vmodel, vopt).detach().clone() for tok, tgt in data]
    ok = all(torch.allclose(a, b, rtol=1e-4, atol=1e-5) for a, b in zip(outs, vouts))
    hist = {}
    for n in step.compiled_func.graph.graph.nodes:
        if n.op == "call_function":
            nm = getattr(n.target, "__name__", str(n.target))
            if any(k in nm for k in ("all_", "reduce_scatter", "scatter_wrapper", "copy_wrapper")):
                hist[nm] = hist.get(nm, 0) + 1
    if rank == 0:
        print(f"GPT2_BUNDLE ok={ok} compile+run {t1 - t0:.1f}s losses {[float(o) for o in outs]} "
              f"vanilla {[float(o) for o in vouts]} hist {hist}", flush=True)
        if record and "plan" in saved and ok:
            with open(record, "w") as f:
                f.write(saved["plan"])
            print(f"recorded {record}: {len(saved['plan']) / 1e6:.2f} MB", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
