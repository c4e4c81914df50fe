"""The end-to-end comparator (tools/parity.py) has teeth: planted defects of the kind a broken
reshard path produces — a shard that was never reduced, a shard swapped with its neighbour, a stale
parameter — are flagged, while vanilla bf16 rounding noise is not (that is what calibrates the
tolerance)."""
import torch

from tools import parity as P


def _setup():
    from easydist_b200.workloads import GPT2, GPT2_CONFIGS, synthetic_tokens
    cfg = GPT2_CONFIGS["gpt2-tiny"]
    torch.manual_seed(0)
    m = GPT2(cfg).bfloat16()
    state0 = {k: v.clone() for k, v in m.state_dict().items()}
    steps = [[synthetic_tokens(cfg, 4, 64, seed=10 * b + r) for r in range(2)] for b in range(3)]
    mk = lambda ps: torch.optim.SGD(ps, lr=1e-2, momentum=0.9, foreach=True)
    ref = P.vanilla_run(lambda: GPT2(cfg), state0, steps, mk, torch.float32, "cpu")
    van = P.vanilla_run(lambda: GPT2(cfg), state0, steps, mk, torch.bfloat16, "cpu")
    return ref, van


def _as_bf16(params, states):
    return ({k: v.bfloat16() for k, v in params.items()},
            {k: {kk: vv.bfloat16() for kk, vv in st.items()} for k, st in states.items()})


def test_comparator_accepts_bf16_noise_and_flags_planted_defects():
    (ref_l, ref_p, ref_s), (van_l, van_p, van_s) = _setup()
    got_p, got_s = _as_bf16(van_p, van_s)
    base = P.compare(got_p, got_s, ref_p, ref_s, low_precision=True)
    tol_state, tol_ulp = max(2e-2, 2.0 * base["state_rel_l2"]), max(2.0, 2.0 * base["param_max_ulp"])
    # the same run again is within its own calibrated tolerance
    assert base["state_rel_l2"] <= tol_state and base["param_max_ulp"] <= tol_ulp
    assert base["checks"] == len(ref_p) + sum(len(s) for s in ref_s.values())
    name = "h.0.c_fc.weight"
    # (1) one rank's half of a gradient was never added (momentum of the first half halves)
    bad_s = {k: dict(v) for k, v in got_s.items()}
    m = bad_s[name]["momentum_buffer"].clone()
    m[: m.shape[0] // 2] *= 0.5
    bad_s[name]["momentum_buffer"] = m
    r = P.compare(got_p, bad_s, ref_p, ref_s, low_precision=True)
    assert r["state_rel_l2"] > tol_state and r["worst"] == f"{name}.momentum_buffer"
    # (2) two shards of a parameter swapped
    bad_p = dict(got_p)
    w = bad_p[name].clone()
    h = w.shape[0] // 2
    bad_p[name] = torch.cat([w[h:], w[:h]])
    r = P.compare(bad_p, got_s, ref_p, ref_s, low_precision=True)
    assert r["param_max_ulp"] > tol_ulp
    # (3) a parameter that missed its updates (stale: still the initial value) — the updates of a
    # bias are far above bf16 resolution at this learning rate
    bname = "h.0.c_fc.bias"
    bad_p = dict(got_p)
    bad_p[bname] = torch.zeros_like(bad_p[bname])
    r = P.compare(bad_p, got_s, ref_p, ref_s, low_precision=True)
    assert r["param_max_ulp"] > tol_ulp
    # fp32 mode = the reference's assert_close(rtol 1e-4, atol 1e-5)
    r = P.compare(ref_p, ref_s, ref_p, ref_s, low_precision=False)
    assert r["assert_close_violation"] == 0.0
    off = {k: v + 3e-4 * v.abs().max() for k, v in ref_p.items()}
    r = P.compare(off, ref_s, ref_p, ref_s, low_precision=False)
    assert r["assert_close_violation"] > 1.0
