"""Edge planners: (src placements, dst placements) -> ordered single-collective steps
`(mesh_dim, cur, tgt)`.  Behavioural mirror of the three planners the reference selects with
EXPERIMENTAL_SHARDING_TRANSFORM (easydist/config.py:115-117):
  GREEDY (default)  easydist/torch/passes/sharding.py:615-650 _gen_transform_infos_greedy
  REPLICATE         sharding.py:672-702 _gen_transform_infos (+ _replicate_then_shard :653-670)
  P2P               sharding.py:477-592 _gen_immediate_transform_infos
Checked against the reference on 885 placement pairs through the oracle fixtures
(tests/test_planners.py).
"""
from itertools import permutations

from .metair import R, VarSPMDStrategy


def plan_greedy(src, dst):
    """Walk mesh dims from the innermost outwards; a target shard whose outer nesting differs
    from the current one is first replicated; a second pass fixes what is left."""
    cur = list(src)
    steps = []
    for i in range(len(cur) - 1, -1, -1):
        want = dst[i]
        if want.is_shard():
            nest_cur = [j for j in range(i) if cur[j].is_shard() and cur[j].dim == want.dim]
            nest_dst = [j for j in range(i) if dst[j].is_shard() and dst[j].dim == want.dim]
            if nest_cur != nest_dst:
                want = R()
        if cur[i] != want:
            steps.append((i, cur[i], want))
            cur[i] = want
    for i, want in enumerate(dst):
        if cur[i] != want:
            steps.append((i, cur[i], want))
            cur[i] = want
    return steps


def _replicate_first_key(step):
    mesh_dim, s, d = step
    if (d.is_replicate() or d.is_partial()) and s.is_shard():
        return -mesh_dim
    if (s.is_replicate() or s.is_partial()) and d.is_shard():
        return mesh_dim
    return 0


def plan_replicate(src, dst):
    """Decompose S(i)->S(j) into S(i)->R->S(j) (always on N-D meshes), then order all gathers
    (inner mesh dims first) before all scatters (outer mesh dims first)."""
    steps = []
    n_src, n_dst = {}, {}
    ndim = len(src)
    for i, (s, d) in enumerate(zip(src, dst)):
        if s.is_shard():
            n_src[s.dim] = n_src.get(s.dim, 0) + 1
        if d.is_shard():
            n_dst[d.dim] = n_dst.get(d.dim, 0) + 1
        if s.is_shard() and d.is_shard() and (ndim > 1 or n_src[s.dim] != n_dst[d.dim]):
            steps.append((i, s, R()))
            steps.append((i, R(), d))
        else:
            steps.append((i, s, d))
    steps.sort(key=_replicate_first_key)
    return steps


def plan_immediate(src, dst):
    """Largest set of single-collective steps over all mesh-dim orders; what cannot be done in
    one step is left for the Partition/P2P box exchange.  Returns (steps, placements reached)."""

    def count(pl, sh):
        return sum(1 for p in pl if p.is_shard() and p.dim == sh.dim)

    best_n, best_steps, best_cur = 0, [], list(src)
    for order in permutations(reversed(range(len(src)))):
        cur, steps = list(src), []

        def can_shard(k, t):
            return count(cur[k + 1:], t) == 0

        def can_unshard(k, c):
            return count(cur[k + 1:], c) == 0

        for k in order:
            c, t = cur[k], dst[k]
            if c == t:
                continue
            if c.is_replicate():
                ok = (t.is_shard() and can_shard(k, t)) or t.is_partial()
            elif c.is_shard():
                if t.is_shard():
                    ok = c.dim != t.dim and can_unshard(k, c) and can_shard(k, t)
                else:
                    ok = can_unshard(k, c)
            else:
                ok = t.is_replicate() or (t.is_shard() and can_shard(k, t))
            if ok:
                steps.append((k, c, t))
                cur[k] = t
        if len(steps) > best_n:
            best_n, best_steps, best_cur = len(steps), steps, cur
    for k, (c, t) in enumerate(zip(best_cur, dst)):
        if c.is_partial() and t.is_shard():  # P2P cannot reduce: P -> R first
            best_steps.append((k, c, R()))
            best_cur[k] = R()
    return best_steps, VarSPMDStrategy(*best_cur)


PLANNERS = {"GREEDY": plan_greedy, "REPLICATE": plan_replicate}


def step_kind(cur, tgt):
    """Which reshard op one step lowers to (sharding.py:739-793); None = nothing to do.
    R->P, S->P and P->P are never emitted by the reference either."""
    if cur == tgt:
        return None
    if tgt.is_shard():
        if cur.is_replicate():
            return "scatter"
        if cur.is_shard():
            return "all_to_all" if cur.dim != tgt.dim else None
        return "reduce_scatter"
    if tgt.is_replicate():
        if cur.is_shard():
            return "all_gather"
        if cur.is_partial():
            return "all_reduce"
    return None


# ---- Partition / P2P box planner (sharding.py:336-474) ------------------------------------------------


class Partition:
    """An N-D box [start, end) of the global tensor held by `rank`, plus the coordinates of the
    partial (unreduced) mesh dims it belongs to (sharding.py:336-374)."""

    __slots__ = ("start", "end", "rank", "partial")

    def __init__(self, start, end, rank, partial=()):
        self.start, self.end, self.rank, self.partial = tuple(start), tuple(end), int(rank), tuple(partial)

    def shard(self, tensor_dim, shard_num, shard_idx):
        s, e = self.start[tensor_dim], self.end[tensor_dim]
        block = (e - s + shard_num - 1) // shard_num
        ns, ne = min(s + block * shard_idx, e), min(s + block * (shard_idx + 1), e)
        return Partition(self.start[:tensor_dim] + (ns,) + self.start[tensor_dim + 1:],
                         self.end[:tensor_dim] + (ne,) + self.end[tensor_dim + 1:], self.rank,
                         self.partial)

    def intersect(self, src):
        """Part of `self` that `src` can provide (None if empty or partial coords differ)."""
        s = tuple(max(a, b) for a, b in zip(self.start, src.start))
        e = tuple(min(a, b) for a, b in zip(self.end, src.end))
        if any(a >= b for a, b in zip(s, e)) or src.partial != self.partial:
            return None
        return Partition(s, e, src.rank, src.partial)

    def shape(self):
        return tuple(e - s for s, e in zip(self.start, self.end))

    def __repr__(self):
        return f"{{{self.start}->{self.end} partial {self.partial} on rank{self.rank}}}"


def partitions_from_spec(placements, global_shape, mesh):
    """Partition of every rank of `mesh` (easydist_b200.device_mesh.DeviceMesh) under `placements`
    (sharding.py:376-398 Partition.from_tensor_spec); returned in mesh order (row-major)."""
    import numpy as np
    parts = []
    for idx in np.ndindex(*mesh.mesh.shape):
        p = Partition((0,) * len(global_shape), tuple(int(v) for v in global_shape),
                      int(mesh.mesh[idx]))
        for mdim, pl in enumerate(placements):
            if pl.is_shard():
                p = p.shard(pl.dim, mesh.mesh.shape[mdim], idx[mdim])
            elif pl.is_partial():
                p = Partition(p.start, p.end, p.rank, p.partial + (idx[mdim],))
        parts.append(p)
    return parts


def recv_boxes(src_parts, tgt_part):
    """Boxes `tgt_part`'s rank must fetch: intersections with the sources ordered by rank distance
    (stable), first provider of each distinct box wins (sharding.py:410-425 gen_recv_meta)."""
    out, seen = [], set()
    for src in sorted(src_parts, key=lambda x: abs(x.rank - tgt_part.rank)):
        inter = tgt_part.intersect(src)
        if inter is None:
            continue
        key = (inter.start, inter.end)
        if key not in seen:
            seen.add(key)
            out.append(inter)
    return out
