"""bench.py's host-side helpers (no GPU): the in-graph timeline is rebuilt from the stamps and the names the library hands out
(`Net::next_trace`: role[layer][.chunk][@MxNxK])."""
import importlib.util
import os

import numpy as np


def _bench():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_step_timeline_from_names_and_stamps():
    b = _bench()
    names = ["fwd0@8192x1024x2000", "fwd_out2@8192x256x512", "dW0.1@976x1024x8192", "xchg_B1", "opt"]
    st = np.zeros((5, 16), np.uint64)
    t0 = 1_000_000
    spans = [(0, 27_000), (28_000, 38_000), (40_000, 59_000), (60_000, 81_000), (82_000, 90_000)]
    for i, (a, e) in enumerate(spans):
        st[i, 0] = t0 + a - 500          # entry
        st[i, 2] = t0 + a                # dependencies resolved
        st[i, 10] = t0 + e               # last CTA exit
    st[0, 3], st[0, 6], st[0, 7], st[0, 8] = t0 + 900, t0 + 13_000, t0 + 15_000, t0 + 26_000
    st[3, 3], st[3, 4] = t0 + 60_000 + 9_500, t0 + 60_000 + 18_000      # LL exchange: pushed / owned runs updated
    tl = b.step_timeline(names, st, 8192, 2000, [1024, 512, 256])
    k = {r["kernel"]: r for r in tl["kernels"]}
    assert set(k) == {"fwd0", "fwd2+out", "dW0.1", "xchg_B1", "opt"}
    assert k["fwd0"]["flops"] == 2 * 8192 * 1024 * 2000 and abs(k["fwd0"]["us"] - 27.0) < 1e-9
    assert abs(k["fwd0"]["tflops"] - k["fwd0"]["flops"] / 27e-6 / 1e12) < 1e-6
    assert k["fwd0"]["cta0"]["first_acc"] == 13.0 and k["fwd2+out"]["cta0"] is None
    assert k["dW0.1"]["flops"] == 2 * 976 * 1024 * 8192
    assert k["xchg_B1"]["flops"] == 0 and k["xchg_B1"]["cta0"]["peers_arrived"] == 9.5 and k["xchg_B1"]["cta0"]["runs_done"] == 18.0
    assert k["xchg_B1"]["cta0"]["fenced"] is None
    assert tl["span_us"] == 90.0 and abs(tl["idle_us"] - (1 + 2 + 1 + 1)) < 1e-9


def test_flops_per_row_matches_the_layer_dims():
    b = _bench()
    cfg = b.CONFIGS["cfg2"]
    dims = [cfg["F"]] + list(cfg["hidden"]) + [1]
    per_layer = [2 * a * c for a, c in zip(dims[:-1], dims[1:])]
    # forward + dW for every layer, dA for all but the first
    want = 3 * sum(per_layer) - per_layer[0]
    f_train, f_hidden, f_score = b.flops_per_row(cfg["F"], cfg["hidden"])
    assert f_train == want and f_hidden == want - 6 * cfg["hidden"][-1] and f_score == sum(per_layer)
