"""Pin the oracle (oracle/reshard_oracle.py) against
  (1) the reference's own known-answer vectors (tests/test_combination/*.py in /root/reference,
      restated here with their file:line), and
  (2) fixtures produced by running the unmodified reference in the CPU container
      (tests/golden/make_golden.py).
CPU only."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import reshard_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- (1) the reference's golden vectors ---------------------------------------------------------

def test_gather_known_answers():
    # tests/test_combination/test_gather.py:26-37
    shards = [np.ones((3, 4))] * 4
    assert np.array_equal(O.comb_gather(shards, 0), np.ones((12, 4)))
    assert np.array_equal(O.comb_gather(shards, 1), np.ones((3, 16)))


def test_gather_halo_known_answers():
    # tests/test_combination/test_gather.py:42-55
    shards = [np.array([1, 1, 1])] * 3
    assert O.comb_gather(shards, 0, halowidth=1).tolist() == [1, 1, 2, 1, 2, 1, 1]
    assert O.comb_gather(shards, 0, halowidth=-1).tolist() == [1, 1, 1, 1, 1]


def test_gather_chunk_known_answer():
    # tests/test_combination/test_gather.py:58-66
    shards = [np.array([1, 2, 3])] * 3
    assert O.comb_gather(shards, 0, chunk_=3).tolist() == [1, 1, 1, 2, 2, 2, 3, 3, 3]


def test_reduce_known_answers():
    # tests/test_combination/test_reduce.py:25-41
    shards = [np.array([i, i, i]) for i in range(4)]
    assert O.comb_reduce(shards, "max").tolist() == [3, 3, 3]
    assert O.comb_reduce(shards, "min").tolist() == [0, 0, 0]
    assert O.comb_reduce(shards, "sum").tolist() == [6, 6, 6]


def test_help_func_known_answers():
    # tests/test_combination/test_help_func.py:22-52
    assert O.aligned_prefix(np.array([1, 2, 3, 4]), np.array([1, 2, 3, 4]), 0) == 4
    assert O.aligned_prefix(np.array([1, 2, 3, 4]), np.array([2, 2, 3, 4]), 0) == 0
    t1 = np.array([[1, 2, 3, 4], [1, 2, 3, 4]])
    t2 = np.array([[1, 2, 3, 4], [1, 2, 3, 5]])
    assert O.aligned_prefix(t1, t2, 0) == 1
    assert O.aligned_prefix(t1, t2, 1) == 3
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 13, 12), 1) is True
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 13, 12), 2) is False
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 13, 13), 1) is False
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 13, 13), 2) is False
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 11, 12, 13), 2) is False
    assert O.shape_aligned_otherdim((10, 11, 12), (10, 11, 12, 13), 3) is False


def test_try_combination_round_trip():
    # tests/test_combination/test_try_combination_single.py:28-87: recover the generating func
    rng = np.random.RandomState(0)
    shards4 = [rng.uniform(size=(3, 4)) for _ in range(4)]
    for op in ("max", "min", "sum"):
        assert O.try_combination_single(shards4, O.comb_reduce(shards4, op)) == ("reduce", op)
    for dim in (0, 1):
        assert O.try_combination_single(shards4, O.comb_gather(shards4, dim)) == \
            ("gather", dim, 0, 1)
    shards3 = [rng.uniform(size=(3, 4)) for _ in range(3)]
    for dim, halo in zip([0, 1], [1, 2]):
        g = O.comb_gather(shards3, dim, halowidth=halo)
        assert O.try_combination_single(shards3, g) == ("gather", dim, halo, 1)
    for dim, ch in zip([0, 1], [3, 2]):
        g = O.comb_gather(shards3, dim, chunk_=ch)
        assert O.try_combination_single(shards3, g) == ("gather", dim, 0, ch)
    same = [shards3[0]] * 3
    assert O.try_combination_single(same, shards3[0]) == ("identity",)


# ---- (2) fixtures generated from the reference ---------------------------------------------------

def _dec(t):
    return tuple(t)


def test_planners_match_reference():
    with gzip.open(os.path.join(GOLDEN, "planners.json.gz"), "rt") as f:
        cases = json.load(f)
    assert len(cases) > 500
    dec = lambda infos: [(i, _dec(a), _dec(b)) for i, a, b in infos]
    for c in cases:
        src = [_dec(t) for t in c["src"]]
        dst = [_dec(t) for t in c["dst"]]
        assert O.gen_transform_infos_greedy(src, dst) == dec(c["greedy"]), (src, dst)
        assert O.gen_transform_infos(src, dst) == dec(c["replicate"]), (src, dst)
        imm, left = O.gen_immediate_transform_infos(src, dst)
        assert imm == dec(c["immediate"]), (src, dst)
        assert left == [_dec(t) for t in c["immediate_left"]], (src, dst)


def test_partition_matches_reference():
    with gzip.open(os.path.join(GOLDEN, "partition.json.gz"), "rt") as f:
        cases = json.load(f)
    assert len(cases) > 1000
    enc = lambda p: [list(p.start), list(p.end), p.rank, list(p.partial)]
    for c in cases:
        mesh = np.arange(int(np.prod(c["mesh"]))).reshape(c["mesh"])
        src = [_dec(t) for t in c["src"]]
        dst = [_dec(t) for t in c["dst"]]
        sp = O.partitions_from_spec(src, c["gshape"], mesh)
        tp = O.partitions_from_spec(dst, c["gshape"], mesh)
        assert [enc(p) for p in sp] == c["src_parts"]
        assert [enc(p) for p in tp] == c["dst_parts"]
        recv = O.gen_recv_meta(sp, tp)
        got = {str(k): [enc(p) for p in v] for k, v in recv.items()}
        assert got == c["recv"], (c["mesh"], c["gshape"], src, dst)


def test_combination_fixtures():
    z = np.load(os.path.join(GOLDEN, "combination.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files})
    n_checked = 0
    for base in keys:
        kind = base.split("_")[0]
        shards = list(z[base + "_in"])
        if kind == "gather":
            dim, hw, ch = (int(v) for v in z[base + "_meta"])
            got = O.comb_gather(shards, dim, halowidth=hw, chunk_=ch)
            assert np.array_equal(got, z[base + "_out"]), base
        elif kind == "halo":
            dim, halo = (int(v) for v in z[base + "_meta"])
            got = O.halo_padding(shards, halo, dim)
            for i, g in enumerate(got):
                assert np.array_equal(g, z[f"{base}_out{i}"]), base
        elif kind == "reduce":
            op = str(z[base + "_op"])
            got = O.comb_reduce(shards, op)
            assert np.allclose(got, z[base + "_out"], rtol=0, atol=1e-6), base
        n_checked += 1
    assert n_checked > 100


@pytest.mark.parametrize("world", [2, 4])
def test_reshard_ops_match_reference(world):
    """The ten callables of sharding.py:94-163 run by the reference under gloo == the oracle,
    bit for bit (inputs are integer-valued so sums are exact)."""
    z = np.load(os.path.join(GOLDEN, f"reshard_w{world}.npz"))
    from tests.golden.make_golden import reshard_cases
    n = 0
    for name, op, kw, shape, dtype in reshard_cases(world):
        key = f"{name}_{dtype}"
        xs = [z[f"{key}__in{r}"] for r in range(world)]
        want = [z[f"{key}__out{r}"] for r in range(world)]
        if op == "all_gather":
            got = O.all_gather(xs, kw["dim"])
        elif op == "all_to_all":
            got = O.all_to_all(xs, kw["g"], kw["s"])
        elif op == "scatter":
            got = [O.scatter(xs[r], world, kw["dim"], r) for r in range(world)]
        elif op == "all_reduce":
            got = O.all_reduce(xs, kw["op"])
        elif op == "reduce_scatter":
            got = O.reduce_scatter(xs, kw["op"], kw["dim"])
        for r in range(world):
            if shape == ():
                # legacy c10d_functional.all_reduce returns a 0-dim input as shape (1,)
                assert want[r].size == 1
                want[r] = want[r].reshape(())
            assert got[r].shape == want[r].shape, (key, r, got[r].shape, want[r].shape)
            assert got[r].dtype == want[r].dtype, (key, r)
            assert np.array_equal(got[r], want[r]), (key, r)
        n += 1
    assert n > 50


def test_scatter_index_error_like_reference():
    # sharding.py:122-123: chunk() may return fewer pieces; indexing past them raises IndexError
    x = np.zeros((5, 2))
    with pytest.raises(IndexError):
        O.scatter(x, 4, 0, 3)


def test_redistribute_simulation_matches_global_semantics():
    """Property: for every planner, applying its transform list to locals of `src` yields the
    locals of `dst` (N-D meshes) — the invariant sharding_transform relies on."""
    g = np.arange(16 * 16, dtype=np.float32).reshape(16, 16)
    alphabet = [O.R, O.S(0), O.S(1), O.P("sum")]
    for mesh_shape in [(2,), (4,), (2, 2), (2, 4)]:
        mesh = np.arange(int(np.prod(mesh_shape))).reshape(mesh_shape)
        nd = len(mesh_shape)
        import itertools
        for src in itertools.product(alphabet, repeat=nd):
            for dst in itertools.product(alphabet[:3], repeat=nd):
                # skip layouts where the same tensor dim is sharded on two mesh dims (nested
                # sharding order matters there and is covered by the planners fixture)
                loc = O.make_locals(g, mesh, src)
                for planner in (O.gen_transform_infos_greedy, O.gen_transform_infos):
                    out = O.apply_transform(loc, mesh, planner(list(src), list(dst)))
                    want = O.make_locals(g, mesh, dst)
                    for r in want:
                        assert np.array_equal(out[r], want[r]), (mesh_shape, src, dst, planner)
