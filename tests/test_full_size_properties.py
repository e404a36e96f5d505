"""Size-independent properties of the training step and the scorer at BASELINE.json's FULL configurations
(cfg1: 1000 cols x 4096 rows, [512,256,128]; cfg2: 2000 cols x 8192 rows, [1024,512,256]), through the C-ABI.

The element-wise oracle comparison lives in test_trainer_parity.py; these tests check what must hold at any size
for the reference's loss (MSE on sigmoid, SUM_BY_NONZERO_WEIGHTS, ssgd_monitor.py:124-126) and its gradient mean
(SyncReplicasOptimizer, ssgd_monitor.py:136-141):
  * linearity in the sample weights          L(2w) = 2 L(w), g(2w) = 2 g(w)
  * zero-weight rows are absent rows         (they count neither in the sum nor in the divisor)
  * row order does not matter
  * the mean of the gradients of R equal shards is the gradient of the whole batch (the data-parallel identity the
    all-reduce relies on), and accumulate-R-then-apply equals one step on the whole batch
  * scores do not depend on how the rows are chunked, and the trainer's predict equals the exported scorer
Float tolerance: the only difference between the two sides of each identity is fp32 summation order (atomic-add
order, split-K plan): 5e-6 relative on the loss (a sum of thousands of fp32 terms), per test relative to max|g| on gradients."""
import numpy as np
import pytest

from oracle import shifu_oracle as so
from util import make_pair

FULL = {
    "cfg1": dict(F=1000, hidden=[512, 256, 128], rows=4096, optimizer=so.OPT_ADAM),
    "cfg2": dict(F=2000, hidden=[1024, 512, 256], rows=8192, optimizer=so.OPT_MOMENTUM),
}


def _trainer(sb, name, precision, optimizer=None, lr=0.05):
    c = FULL[name]
    acts = [so.ACT_RELU] * len(c["hidden"])
    net, params, cfg, desc = make_pair(sb, c["F"], c["hidden"], acts, optimizer=c["optimizer"] if optimizer is None else optimizer,
                                       lr=lr, max_batch=c["rows"], precision=precision)
    t = sb.Trainer(desc)
    t.set_params(so.flatten_params(params))
    return c, net, desc, t


def _grad(t, X, y, w):
    """loss and mini-batch gradient WITHOUT an update (the accumulate leg of the epoch-sync schedule)"""
    L = t.accumulate(X, y, w)
    return L, t.get_grads()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_full_size_weight_linearity_and_zero_weight_rows(sb, name):
    c, net, desc, t = _trainer(sb, name, sb.PREC_BF16)
    with t:
        X, y, w = so.synth_batch(c["rows"], c["F"], 11, weights="mixed")
        L1, g1 = _grad(t, X, y, w)
        L2, g2 = _grad(t, X, y, 2.0 * w)
        gmax = np.abs(g1).max()
        assert gmax > 0 and np.isfinite(g1).all()
        assert abs(L2 - 2 * L1) <= 5e-6 * abs(L1) + 1e-9
        assert np.abs(g2 - 2 * g1).max() <= 1e-5 * gmax          # power-of-two scaling commutes with bf16 rounding
        # rows with w == 0 removed from the batch: same loss, same gradient (divisor = number of non-zero weights)
        keep = (w[:, 0] != 0)
        L3, g3 = _grad(t, np.ascontiguousarray(X[keep]), y[keep], w[keep])
        assert abs(L3 - L1) <= 5e-6 * abs(L1) + 1e-9
        assert np.abs(g3 - g1).max() <= 2e-5 * gmax
        # all-zero weights: loss 0, gradient 0 (the _safe_div of SUM_BY_NONZERO_WEIGHTS)
        L0, g0 = _grad(t, X, y, np.zeros_like(w))
        assert L0 == 0.0 and not g0.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_full_size_row_permutation_invariance(sb, name):
    c, net, desc, t = _trainer(sb, name, sb.PREC_BF16)
    with t:
        X, y, w = so.synth_batch(c["rows"], c["F"], 12, weights="mixed")
        L1, g1 = _grad(t, X, y, w)
        p = np.random.RandomState(0).permutation(c["rows"])
        L2, g2 = _grad(t, np.ascontiguousarray(X[p]), y[p], w[p])
        assert abs(L2 - L1) <= 5e-6 * abs(L1)
        assert np.abs(g2 - g1).max() <= 2e-5 * np.abs(g1).max()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_full_size_shard_mean_is_whole_batch_gradient(sb, name, precision):
    """R = 4 equal shards with all weights non-zero: mean over shards of (loss, gradient) == whole batch.
    This is the identity the gradient all-reduce implements; here it is checked on one GPU at full size."""
    c, net, desc, t = _trainer(sb, name, precision)
    with t:
        X, y, w = so.synth_batch(c["rows"], c["F"], 13, weights="ones")
        w *= np.random.RandomState(1).choice(np.array([0.5, 1.0, 2.0], np.float32), size=w.shape)
        L, g = _grad(t, X, y, w)
        R = 4
        n = c["rows"] // R
        Ls, gs = [], np.zeros_like(g, dtype=np.float64)
        for r in range(R):
            s = slice(r * n, (r + 1) * n)
            Lr, gr = _grad(t, np.ascontiguousarray(X[s]), y[s], w[s])
            Ls.append(Lr); gs += gr
        assert abs(np.mean(Ls) - L) <= 5e-6 * abs(L)
        assert np.abs(gs / R - g).max() <= 2e-5 * np.abs(g).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_full_size_accumulate_then_apply_equals_one_step(sb, name):
    """accumulate R shards + apply_accumulated (the reference's one-update-per-epoch schedule) == one step on the
    whole batch, parameters compared after the update (optimizer of the config: Adam / Momentum)."""
    c, net, desc, ta = _trainer(sb, name, sb.PREC_BF16, lr=0.01)
    _, _, _, tb = _trainer(sb, name, sb.PREC_BF16, lr=0.01)
    with ta, tb:
        X, y, w = so.synth_batch(c["rows"], c["F"], 14, weights="ones")
        theta0 = ta.get_params()
        ta.step(X, y, w)
        R = 2
        n = c["rows"] // R
        for r in range(R):
            s = slice(r * n, (r + 1) * n)
            tb.accumulate(np.ascontiguousarray(X[s]), y[s], w[s])
        tb.apply_accumulated()
        pa, pb = ta.get_params(), tb.get_params()
        moved = np.abs(pa - theta0).max()
        assert moved > 0
        # Adam normalises the step, so a gradient that differs in the last fp32 bits can move a weight whose gradient
        # is ~0 by up to lr; compare where the gradient is not negligible, and everything else to lr.
        g = ta.get_grads()
        big = np.abs(g) > 1e-3 * np.abs(g).max()
        assert np.abs(pa - pb)[big].max() <= 1e-3 * moved
        assert np.abs(pa - pb).max() <= 2.1 * 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_full_size_gradient_is_deterministic_enough_and_finite(sb, name):
    """same batch twice without an update: loss and gradients equal up to atomic-add ordering"""
    c, net, desc, t = _trainer(sb, name, sb.PREC_BF16)
    with t:
        X, y, w = so.synth_batch(c["rows"], c["F"], 15, weights="mixed")
        L1, g1 = _grad(t, X, y, w)
        L2, g2 = _grad(t, X, y, w)
        assert abs(L1 - L2) <= 5e-6 * abs(L1)
        assert np.abs(g1 - g2).max() <= 1e-6 * np.abs(g1).max()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_full_size_scores_do_not_depend_on_chunking(sb, precision):
    """cfg4's net (2000 cols, [1024,512,256]): every row's score is a function of that row alone, so scoring the
    set whole, in ragged chunks or reversed gives bit-identical numbers; trainer.predict == exported scorer."""
    c = FULL["cfg2"]
    acts = [so.ACT_RELU] * 3
    net, params, cfg, desc = make_pair(sb, c["F"], c["hidden"], acts, max_batch=4096, precision=precision)
    flat = so.flatten_params(params)
    rows = 40000 + 17 if precision else 6000 + 17
    X = np.clip(np.random.RandomState(5).standard_normal((rows, c["F"])), -4, 4).astype(np.float32)
    with sb.Model.create(desc, flat) as m:
        whole = m.score(X)
        assert whole.shape[0] == rows and np.isfinite(whole).all() and (whole > 0).all() and (whole < 1).all()
        parts, i = [], 0
        for n in [1, 127, 128, 129, 4096, 1000, 1 << 30]:
            if i < rows:
                parts.append(m.score(np.ascontiguousarray(X[i:i + n]))); i = min(rows, i + n)
        np.testing.assert_array_equal(np.concatenate(parts).ravel(), whole.ravel())
        rev = m.score(np.ascontiguousarray(X[::-1]))
        np.testing.assert_array_equal(rev.ravel()[::-1], whole.ravel())
        with sb.Trainer(desc) as t:
            t.set_params(flat)
            pred = t.predict(X[:4096])
        tol = 1e-6 if precision == 0 else 1e-3      # bf16: the two paths may round one activation differently
        assert np.abs(pred.ravel() - whole[:4096].ravel()).max() <= tol
        ref = so.score_rows(net, params, X[:512].astype(np.float64))
        assert np.abs(np.asarray(ref).ravel() - whole[:512].ravel()).max() <= (1e-5 if precision == 0 else 2e-2)
