"""Parity of the CUDA training step against the CPU oracle, through the C-ABI (sb_trainer_*).

Tolerances (BASELINE.json north_star): per-step loss and gradients within 1e-4 absolute in fp32 parity mode.
The bf16 performance mode is checked against the same oracle with a bf16-sized tolerance, stated per test."""
import numpy as np
import pytest

from oracle import shifu_oracle as so
from util import make_pair

ACT_NAMES = {0: "sigmoid", 1: "tanh", 2: "relu", 3: "leakyrelu"}


def _one_step(sb, n_features, hidden, acts, rows, loss, optimizer, precision, weights="mixed", seed=3, lr=0.05):
    net, params, cfg, desc = make_pair(sb, n_features, hidden, acts, loss=loss, optimizer=optimizer, lr=lr,
                                       max_batch=rows, precision=precision)
    X, y, w = so.synth_batch(rows, n_features, seed, weights=weights)
    ref = so.CleanTrainer(net, params, cfg, loss=loss)
    ref_loss = ref.step([(X, y, w)])[0]
    with sb.Trainer(desc) as t:
        t.set_params(so.flatten_params(params))
        got_loss = t.step(X, y, w)
        return ref_loss, ref.last_grads, ref.theta, got_loss, t.get_grads(), t.get_params()


# the two parity modes: fp32 FFMA on the CUDA cores, and fp32-class accuracy on the tensor cores (three bf16 parts per value,
# six tcgen05 products per contraction) - both must meet the north star's fp32 tolerances
FP32_MODES = [0, 2]     # sb.PREC_FP32, sb.PREC_FP32_TC


@pytest.mark.gpu
@pytest.mark.parametrize("prec", FP32_MODES)
@pytest.mark.parametrize("act", [0, 1, 2, 3])
@pytest.mark.parametrize("loss", [so.LOSS_MSE, so.LOSS_SIGMOID_CE])
def test_fp32_step_cfg0_all_activations(sb, act, loss, prec):
    """cfg0: 200 cols, [100, 50], B=100 (the reference's hard-coded BATCH_SIZE, ssgd_monitor.py:33)."""
    rl, rg, rt, gl, gg, gt = _one_step(sb, 200, [100, 50], [act, act], 100, loss, so.OPT_SGD, prec)
    assert abs(gl - rl) <= 1e-4
    assert np.abs(gg - rg).max() <= 1e-4
    # in practice fp32 vs fp32 agrees far tighter than the contract; keep a regression guard too
    assert np.abs(gg - rg).max() <= 5e-6 + 1e-4 * np.abs(rg).max()
    assert np.abs(gt - rt).max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("prec", FP32_MODES)
@pytest.mark.parametrize("optimizer", [so.OPT_ADADELTA, so.OPT_ADAM, so.OPT_SGD, so.OPT_MOMENTUM])
def test_fp32_three_steps_each_optimizer(sb, optimizer, prec):
    net, params, cfg, desc = make_pair(sb, 64, [48, 24], [so.ACT_TANH, so.ACT_RELU], optimizer=optimizer, lr=0.05,
                                       max_batch=96, precision=prec)
    ref = so.CleanTrainer(net, params, cfg)
    with sb.Trainer(desc) as t:
        t.set_params(so.flatten_params(params))
        for s in range(3):
            X, y, w = so.synth_batch(96, 64, 100 + s, weights="mixed")
            rl = ref.step([(X, y, w)])[0]
            gl = t.step(X, y, w)
            assert abs(gl - rl) <= 1e-4, "step %d" % s
            # Adam divides by sqrt(v) ~ |g|: a coordinate whose gradient is ~1e-7 turns a 1e-9 summation-order difference into
            # 1e-4 of parameter movement (lr = 0.05 here); loss and gradients are what the contract bounds
            tol_p = 5e-4 if (optimizer == so.OPT_ADAM and prec == 2) else 2e-5
            assert np.abs(t.get_params() - ref.theta).max() <= tol_p, "step %d" % s
        assert t.global_step == 3


@pytest.mark.gpu
@pytest.mark.parametrize("prec", FP32_MODES)
def test_fp32_ragged_shapes_and_zero_weights(sb, prec):
    """odd widths (not multiples of 4/8/32), rows not a multiple of 32, and an all-zero weight batch
    (loss must be 0 and no update must happen: _safe_div in SUM_BY_NONZERO_WEIGHTS)."""
    net, params, cfg, desc = make_pair(sb, 37, [19, 7], [so.ACT_LEAKYRELU, so.ACT_SIGMOID], optimizer=so.OPT_SGD,
                                       max_batch=101, precision=prec)
    X, y, w = so.synth_batch(101, 37, 5, weights="mixed")
    ref = so.CleanTrainer(net, params, cfg)
    rl = ref.step([(X, y, w)])[0]
    with sb.Trainer(desc) as t:
        flat0 = so.flatten_params(params)
        t.set_params(flat0)
        assert abs(t.step(X, y, w) - rl) <= 1e-5
        assert np.abs(t.get_grads() - ref.last_grads).max() <= 1e-5
        before = t.get_params()
        assert t.step(X, y, np.zeros_like(w)) == 0.0
        np.testing.assert_array_equal(t.get_grads(), np.zeros_like(before))
        np.testing.assert_array_equal(t.get_params(), before)
        # w = None means all ones
        ref2 = so.CleanTrainer(net, so.unflatten_params(net, before.copy()), cfg)
        rl2 = ref2.step([(X, y, np.ones_like(w))])[0]
        assert abs(t.step(X, y, None) - rl2) <= 1e-5


def _bf16_case(sb, F, hidden, acts, rows, loss, weights, seed=3):
    net, params, cfg, desc = make_pair(sb, F, hidden, acts, loss=loss, optimizer=so.OPT_SGD, max_batch=rows,
                                       precision=sb.PREC_BF16)
    X, y, w = so.synth_batch(rows, F, seed, weights=weights)
    L32, g32, _ = so.loss_and_grads(net, params, X, y, w, loss)
    Lb, gb, _ = so.loss_and_grads_bf16(net, params, X, y, w, loss, fused_out=hidden[-1] <= 256)
    with sb.Trainer(desc) as t:
        t.set_params(so.flatten_params(params))
        return net, L32, so.flatten_params(g32), Lb, so.flatten_params(gb), t.step(X, y, w), t.get_grads()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name,F,hidden,rows", [("cfg0", 200, [100, 50], 100), ("cfg1", 1000, [512, 256, 128], 4096),
                                                    ("cfg2", 2000, [1024, 512, 256], 8192)])
def test_bf16_step_against_bf16_oracle(sb, cfg_name, F, hidden, rows):
    """tcgen05 path (bf16 operands, fp32 TMEM accumulation) against the oracle that rounds to bf16 at exactly the
    points the kernels do (oracle.loss_and_grads_bf16).  What is left is fp32-vs-fp64 accumulation order, so the
    bound is tight: 1e-5 absolute on the loss, 2e-3 of the gradient's max magnitude on gradients (an activation
    sitting on a rounding boundary may flip one bf16 ulp).  The distance to the pure fp32 oracle is the bf16
    quantisation itself and is only sanity-bounded here (it is reported in DESIGN.md)."""
    acts = [so.ACT_RELU] * len(hidden)
    net, L32, g32, Lb, gb, gl, gg = _bf16_case(sb, F, hidden, acts, rows, so.LOSS_MSE, "ones")
    assert abs(gl - Lb) <= 2e-5
    gmax = np.abs(gb).max()
    assert np.abs(gg - gb).max() <= 2e-3 * gmax, (np.abs(gg - gb).max(), gmax)
    assert abs(gl - L32) <= 2e-3 * abs(L32)
    assert np.abs(gg - g32).max() <= 0.15 * np.abs(g32).max()
    for a, b in zip(so.unflatten_params(net, gg), so.unflatten_params(net, g32)):   # direction per block
        a = a.ravel().astype(np.float64); b = b.ravel().astype(np.float64)
        if np.linalg.norm(b) > 0:
            assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300) > 0.995


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_bf16_small_all_activations(sb, act):
    """ragged everything (72 cols, widths 40/24, 130 rows), every activation, CE loss, mixed weights"""
    net, L32, g32, Lb, gb, gl, gg = _bf16_case(sb, 72, [40, 24], [act, act], 130, so.LOSS_SIGMOID_CE, "mixed")
    assert abs(gl - Lb) <= 1e-5
    assert np.abs(gg - gb).max() <= 2e-3 * np.abs(gb).max()
    assert abs(gl - L32) <= 5e-3 * max(1e-3, abs(L32))


@pytest.mark.gpu
def test_resident_dataset_equals_host_steps(sb):
    """sb_trainer_step_resident over an HBM-resident set must equal sb_trainer_step fed the same rows."""
    net, params, cfg, desc = make_pair(sb, 50, [32, 16], [so.ACT_RELU, so.ACT_TANH], optimizer=so.OPT_MOMENTUM,
                                       max_batch=64, precision=sb.PREC_FP32)
    X, y, w = so.synth_batch(64 * 4 + 13, 50, 11, weights="mixed")
    flat = so.flatten_params(params)
    with sb.Trainer(desc) as a, sb.Trainer(desc) as b:
        a.set_params(flat); b.set_params(flat)
        b.load_dataset(X, y, w)
        offs = [(0, 64), (64, 64), (128, 64), (192, 64), (256, 13)]
        for off, n in offs:
            la = a.step(X[off:off + n], y[off:off + n], w[off:off + n])
            lb = b.step_resident(off, n)
            assert abs(la - lb) <= 1e-6
        assert np.abs(a.get_params() - b.get_params()).max() <= 1e-6
        with pytest.raises(sb.ShifuB200Error):
            b.step_resident(260, 64)      # runs past the end of the resident set
        with pytest.raises(sb.ShifuB200Error):
            a.step(X[:65], y[:65], w[:65])  # rows > max_batch


@pytest.mark.gpu
def test_epoch_sync_schedule_matches_syncreplicas_oracle(sb):
    """Reference schedule (D3): mean of R mini-batch gradients, one Adadelta update per 'epoch'
    (ssgd_monitor.py:136-141).  Oracle = the clean mean-of-batch-means; batches from np.array_split."""
    n_rows, F = 1030, 30
    net, params, cfg, desc = make_pair(sb, F, [20, 10], [so.ACT_TANH, so.ACT_TANH], optimizer=so.OPT_ADADELTA, lr=1.0,
                                       max_batch=128, precision=sb.PREC_FP32)
    X, y, w = so.synth_batch(n_rows, F, 21, weights="mixed")
    batches = so.split_batches(n_rows, 100)
    assert len(batches) == 10 and {len(b) for b in batches} == {103}
    theta = so.flatten_params(params).astype(np.float32)
    opt = so.Optimizer(cfg, theta.size)
    with sb.Trainer(desc) as t:
        t.set_params(theta)
        for epoch in range(2):
            P = so.unflatten_params(so.NetDesc(F, [20, 10], [1, 1]), theta)
            gsum = np.zeros_like(theta)
            for idx in batches:
                L, g, _ = so.loss_and_grads(net, P, X[idx], y[idx], w[idx])
                gsum += so.flatten_params(g)
                gl = t.accumulate(X[idx], y[idx], w[idx])
                assert abs(gl - L) <= 1e-5
            theta = opt.apply(theta, gsum / np.float32(len(batches)))
            t.apply_accumulated()
            assert np.abs(t.get_grads() - gsum / len(batches)).max() <= 1e-5
            assert np.abs(t.get_params() - theta).max() <= 1e-4


@pytest.mark.gpu
def test_bf16_resident_set_is_read_in_place(sb):
    """bf16 mode keeps the resident set as bf16 in HBM and the layer-0 GEMMs TMA-load the batch from it at a row offset
    (no load kernel).  Same rows fed from the host must give the same step; offsets / sizes deliberately ragged, the
    last batch ends exactly at the end of the set, weights include zeros (n_nz comes from the prefix counts)."""
    net, params, cfg, desc = make_pair(sb, 72, [40, 24], [so.ACT_RELU, so.ACT_TANH], optimizer=so.OPT_ADAM, max_batch=130,
                                       precision=sb.PREC_BF16)
    X, y, w = so.synth_batch(130 * 3 + 37, 72, 5, weights="mixed")
    flat = so.flatten_params(params)
    with sb.Trainer(desc) as a, sb.Trainer(desc) as b:
        a.set_params(flat); b.set_params(flat)
        b.load_dataset(X, y, w)
        for off, n in [(0, 130), (130, 130), (263, 101), (390, 37), (3, 130)]:
            la = a.step(X[off:off + n], y[off:off + n], w[off:off + n])
            lb = b.step_resident(off, n)
            assert abs(la - lb) <= 1e-6, (off, n, la, lb)
            assert np.abs(a.get_grads() - b.get_grads()).max() <= 1e-6 * max(1.0, np.abs(a.get_grads()).max())
        assert np.abs(a.get_params() - b.get_params()).max() <= 1e-6
        # epoch-sync schedule over the resident set
        la = a.accumulate(X[:130], y[:130], w[:130]); lb = b.accumulate_resident(0, 130)
        assert abs(la - lb) <= 1e-6
        a.apply_accumulated(); b.apply_accumulated()
        assert np.abs(a.get_params() - b.get_params()).max() <= 1e-6


@pytest.mark.gpu
def test_step_async_pipeline_equals_synchronous_steps(sb):
    """sb_trainer_step_async (double-buffered H2D on a copy stream) must produce the synchronous trajectory (up to the
    order of the fp32 atomic adds inside a step, ~1e-7 on a parameter)"""
    net, params, cfg, desc = make_pair(sb, 64, [48, 24], [so.ACT_RELU, so.ACT_TANH], optimizer=so.OPT_ADAM, max_batch=96,
                                       precision=sb.PREC_FP32)
    batches = [so.synth_batch(96 if s % 3 else 80, 64, 50 + s, weights="mixed") for s in range(7)]
    flat = so.flatten_params(params)
    with sb.Trainer(desc) as a, sb.Trainer(desc) as b:
        a.set_params(flat); b.set_params(flat)
        for X, y, w in batches:
            la = a.step(X, y, w)
            b.step_async(X, y, w)
        assert abs(b.last_loss() - la) <= 1e-7
        np.testing.assert_allclose(a.get_params(), b.get_params(), rtol=0, atol=2e-6)
        assert b.global_step == 7


@pytest.mark.gpu
@pytest.mark.parametrize("n_steps", [1, 4, 7, 9])
def test_run_resident_equals_step_resident(sb, n_steps):
    """sb_trainer_run_resident (four steps per captured graph, descriptors of the next chunk written ahead) must follow
    the trajectory of n single sb_trainer_step_resident calls, and single steps must continue correctly after it.
    Momentum keeps the comparison linear in the gradient (atomic-add order noise ~1e-7)."""
    rows, nb, F = 64, 5, 72
    net, params, cfg, desc = make_pair(sb, F, [48, 24], [so.ACT_RELU, so.ACT_TANH], optimizer=so.OPT_MOMENTUM, lr=0.05,
                                       max_batch=rows, precision=sb.PREC_BF16)
    X, y, w = so.synth_batch(rows * nb, F, 21, weights="mixed")
    offs = [((i * 3) % nb) * rows for i in range(n_steps)]
    flat = so.flatten_params(params)
    with sb.Trainer(desc) as a, sb.Trainer(desc) as b:
        for t in (a, b):
            t.set_params(flat); t.load_dataset(X, y, w)
        la = None
        for o in offs:
            la = a.step_resident(o, rows)
        b.run_resident(offs, rows)
        assert abs(b.last_loss() - la) <= 1e-6
        assert b.global_step == n_steps == a.global_step
        np.testing.assert_allclose(a.get_params(), b.get_params(), rtol=0, atol=2e-6)
        assert np.abs(a.get_params() - flat).max() > 1e-4          # something was learned
        # single steps after a run (the descriptor prefetch re-joins the main stream), then another run
        assert abs(a.step_resident(rows, rows) - b.step_resident(rows, rows)) <= 1e-6
        b.run_resident(offs[::-1] + offs, rows)
        for o in offs[::-1] + offs:
            la = a.step_resident(o, rows)
        assert abs(b.last_loss() - la) <= 1e-6
        np.testing.assert_allclose(a.get_params(), b.get_params(), rtol=0, atol=5e-6)
