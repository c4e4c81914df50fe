"""lowering.bucket_small_comm on a hand-built graph, executed over gloo (world 2): small
reduce-scatters along different dims, all-reduces and dim-0 all-gathers whose results are read late
collapse into one collective per kind; values are unchanged bit for bit (integer-valued inputs)."""
import os

import torch
import torch.distributed as dist

from tests._procs import run_world


def _worker(rank, world, port, q):
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    import numpy as np
    from torch.fx import Graph, GraphModule
    from easydist_b200 import lowering
    from tests import gloo_ops as ops
    ops.init_groups(np.arange(world).reshape((world,)))
    group = list(range(world))
    g = Graph()
    shapes = [(4 * world, 3), (1, 8 * world), (2, 3 * world, 5), (6,), (3, 4), (5,), (2, 7)]
    phs = [g.placeholder(f"x{i}") for i in range(len(shapes))]
    outs = []
    for ph, dim in zip(phs[:3], (0, 1, -2)):              # reduce-scatters, three scatter dims
        s_ = g.call_function(ops.reduce_scatter_start, (ph, "sum", dim, group))
        outs.append(g.call_function(ops.reduce_scatter_end, (s_, "sum", dim, group)))
    for ph in phs[3:5]:                                    # all-reduces
        s_ = g.call_function(ops.all_reduce_start, (ph, "sum", group))
        outs.append(g.call_function(ops.all_reduce_end, (s_, "sum", group)))
    for ph in phs[5:]:                                     # dim-0 all-gathers
        s_ = g.call_function(ops.all_gather_start, (ph, 0, group))
        outs.append(g.call_function(ops.all_gather_end, (s_, 0, group)))
    total = None
    for o in outs:                                         # every result is read late
        sm = g.call_function(torch.ops.aten.sum.default, (o,))
        total = sm if total is None else g.call_function(torch.ops.aten.add.Tensor, (total, sm))
    g.output((tuple(outs), total))
    gm = GraphModule(torch.nn.Module(), g)
    gen = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randint(-8, 9, s, generator=gen).float() for s in shapes]
    lowering.propagate_local_meta(gm, xs)
    want, want_total = gm(*xs)
    before = lowering.count_nodes(gm, ops)
    done = lowering.bucket_small_comm(gm, ops)
    lowering.propagate_local_meta(gm, xs)
    got, got_total = gm(*xs)
    after = lowering.count_nodes(gm, ops)
    ok = all(a.shape == b.shape and a.is_contiguous() and torch.equal(a, b) for a, b in zip(got, want)) \
        and torch.equal(got_total, want_total)
    if rank == 0:
        q.put((ok, done, before, after))
    dist.barrier()
    dist.destroy_process_group()


def test_small_collectives_bucket_without_changing_values():
    ok, done, before, after = run_world(_worker, 2, lambda r, port, q: (r, 2, port, q), timeout=240)
    assert ok
    assert done == {"all_reduce": 1, "all_gather": 1, "reduce_scatter": 1}, done
    assert (before["reduce_scatter_start"], before["all_reduce_start"], before["all_gather_start"]) == (3, 2, 2)
    assert (after["reduce_scatter_start"], after["all_reduce_start"], after["all_gather_start"]) == (1, 1, 1)
