"""Compatibility layer that lets the UNMODIFIED reference (/root/reference) import and run its
auto-SPMD path on CPU/gloo in this container.  TEST INFRASTRUCTURE ONLY: used by
tests/golden/make_golden.py to generate golden fixtures and by the oracle-pinning tests when
/root/reference exists.  Nothing in easydist_b200/ imports this, and it never travels to the GPU
box in a way the product could use (the reference itself is absent there).

Recipe: SURVEY.md Appendix A (verified there): inert stubs for optional deps, a `mip` provider on
scipy/HiGHS, a `bitarray` stand-in, torch-2.11 shims and three post-import patches.
"""
import functools
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EASYDIST_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_installed = False


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "easydist"))


def _torch_shims():
    import torch
    import torch.distributed.tensor._utils as tu
    from torch.distributed.tensor import placement_types as pt_new
    import torch.distributed._tensor.placement_types as pt_old

    # (1) compute_local_shape was folded into compute_local_shape_and_global_offset
    if not hasattr(tu, "compute_local_shape"):
        def compute_local_shape(shape, mesh, placements):
            return tu.compute_local_shape_and_global_offset(shape, mesh, placements)[0]
        tu.compute_local_shape = compute_local_shape

    # (2) _Partial(reduce_op) -> Partial(str)
    try:
        from torch.distributed.distributed_c10d import ReduceOp
        names = {ReduceOp.SUM: "sum", ReduceOp.MAX: "max", ReduceOp.MIN: "min", ReduceOp.AVG: "avg"}
    except Exception:  # pragma: no cover
        names = {}

    def _Partial(reduce_op="sum"):
        op = names.get(reduce_op, reduce_op)
        return pt_new.Partial(op if isinstance(op, str) else "sum")

    for mod in (pt_new, pt_old):
        mod._Partial = _Partial  # 2.11 keeps a `_Partial` alias that no longer takes reduce_op

    # (5) CPU plumbing: compile_auto.py:126 queries the CUDA device unconditionally
    if not torch.cuda.is_available():
        torch.cuda.get_device_properties = lambda *a, **k: types.SimpleNamespace(
            total_memory=64 << 30, name="cpu-stub", multi_processor_count=1)
        torch.cuda.current_device = lambda: 0


def _post_import_patches():
    import torch
    import easydist.torch.device_mesh as dm
    from easydist.torch import compile_auto
    import easydist.torch.utils as ed_utils
    from torch.distributed.device_mesh import DeviceMesh

    # (3) NDDeviceMesh.__getitem__ pokes removed DeviceMesh internals
    map_binding = getattr(dm.NDDeviceMesh, "_NDDeviceMesh__map_binding")

    def _getitem(self, names):
        key = names if isinstance(names, str) else tuple(names)
        cache = self.__dict__.setdefault("_edb_sub_cache", {})
        if key in cache:
            return cache[key]
        resolved = tuple(map_binding(self, names))
        if resolved == tuple(self._device_mesh.mesh_dim_names):
            sub = dm.NDDeviceMesh(self._device_mesh)
        else:
            sub = dm.NDDeviceMesh(self._device_mesh[resolved])
        cache[key] = sub
        return sub

    dm.NDDeviceMesh.__getitem__ = _getitem

    # (4) make_fx records profiler nodes from Optimizer.step; strip them before the passes
    orig_pre = compile_auto.preprocess_traced_graph

    def preprocess(fx_module):
        for node in reversed(list(fx_module.graph.nodes)):
            if node.op == "call_function" and "profiler._record_function" in str(node.target):
                if len(node.users) == 0:
                    fx_module.graph.erase_node(node)
        fx_module.recompile()
        return orig_pre(fx_module)

    compile_auto.preprocess_traced_graph = preprocess

    # (6) stride-dependent aten.view becomes illegal after reshards -> retry as reshape
    orig_meta = ed_utils.create_meta_from_node

    def create_meta_from_node(node):
        try:
            return orig_meta(node)
        except (RuntimeError, ValueError) as e:  # utils.py:57-65 re-raises as ValueError(msg)
            aten = torch.ops.aten
            if node.target in (aten.view.default, aten._unsafe_view.default) and \
                    "view" in str(e).lower():
                node.target = aten.reshape.default
                return orig_meta(node)
            raise

    ed_utils.create_meta_from_node = create_meta_from_node

    # (7) fix_embedding.md_embedding range-checks the indices with `.item()` (fix_embedding.py:19-21);
    # torch 2.11 FakeTensors turn that into an unbacked symbol and the flop-counting dry run of
    # reachability.py:52-55 dies on it.  Keep the check for real tensors only.
    import importlib
    # (`easydist.torch.passes.fix_embedding` the attribute is the function of the same name)
    fe = importlib.import_module("easydist.torch.passes.fix_embedding")
    fe = sys.modules[fe.__module__] if not hasattr(fe, "md_embedding") else fe
    from torch._subclasses.fake_tensor import FakeTensor

    def md_embedding(weight, indices, padding_idx=-1, scale_grad_by_freq=False, sparse=False):
        if not isinstance(indices, FakeTensor) and indices.numel() and \
                int(torch.max(indices).item()) >= weight.shape[0]:
            raise RuntimeError("embedding indice overflow")
        return torch.ops.aten.embedding.default(weight, indices, padding_idx, scale_grad_by_freq,
                                                sparse)

    md_embedding.__module__ = fe.__name__
    fe.md_embedding = md_embedding
    import easydist.torch.passes.sharding as sh
    import easydist.torch.passes.edinfo_utils as eu
    sh.create_meta_from_node = create_meta_from_node
    eu.create_meta_from_node = create_meta_from_node


def install():
    """Make `import easydist` work. Idempotent. Raises if the reference is absent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    _torch_shims()
    import easydist  # noqa: F401
    import easydist.torch  # noqa: F401
    _post_import_patches()
    _installed = True
