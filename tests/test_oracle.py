"""CPU tests of the oracle itself (no GPU): analytic gradients vs finite differences in fp64, optimizer forms,
SyncReplicas state machine, load_data restatement."""
import gzip
import os

import numpy as np
import pytest

from oracle import shifu_oracle as so


@pytest.mark.parametrize("loss", [so.LOSS_MSE, so.LOSS_SIGMOID_CE])
@pytest.mark.parametrize("act", [so.ACT_SIGMOID, so.ACT_TANH, so.ACT_LEAKYRELU])
def test_backward_matches_finite_differences_fp64(loss, act):
    net = so.NetDesc(9, [7, 5], [act, act])
    params = so.xavier_init(net, 3, dtype=np.float64)
    X, y, w = so.synth_batch(23, 9, 4, weights="mixed")
    X, y, w = X.astype(np.float64), y.astype(np.float64), w.astype(np.float64)
    L, grads, _ = so.loss_and_grads(net, params, X, y, w, loss)
    flat, g = so.flatten_params(params), so.flatten_params(grads)
    rng = np.random.RandomState(0)
    for i in rng.choice(flat.size, 25, replace=False):
        e = np.zeros_like(flat); e[i] = 1e-6
        Lp = so.loss_and_grads(net, so.unflatten_params(net, flat + e), X, y, w, loss)[0]
        Lm = so.loss_and_grads(net, so.unflatten_params(net, flat - e), X, y, w, loss)[0]
        assert abs((Lp - Lm) / 2e-6 - g[i]) <= 1e-6 + 1e-5 * abs(g[i])


def test_mse_sum_by_nonzero_weights():
    z = np.array([[0.0], [2.0], [-1.0]], np.float32)
    yhat = so._sigmoid(z)
    y = np.array([[1.0], [0.0], [1.0]], np.float32)
    w = np.array([[2.0], [0.0], [0.5]], np.float32)
    L, n = so.loss_value(z, yhat, y, w, so.LOSS_MSE)
    assert n == 2
    expect = (2.0 * (0.5 - 1) ** 2 + 0.5 * (yhat[2, 0] - 1) ** 2) / 2
    assert abs(L - expect) < 1e-7
    assert so.loss_value(z, yhat, y, np.zeros_like(w), so.LOSS_MSE) == (0, 0)


def test_optimizer_forms_one_step():
    th = np.array([1.0, -2.0], np.float32); g = np.array([0.5, -0.25], np.float32)
    sgd = so.Optimizer(so.OptConfig(kind=so.OPT_SGD, lr=0.1), 2)
    np.testing.assert_allclose(sgd.apply(th, g), th - 0.1 * g, rtol=1e-7)
    mom = so.Optimizer(so.OptConfig(kind=so.OPT_MOMENTUM, lr=0.1, momentum=0.9), 2)
    t1 = mom.apply(th, g); t2 = mom.apply(t1, g)
    np.testing.assert_allclose(t2, th - 0.1 * g - 0.1 * (0.9 * g + g), rtol=1e-6)
    adam = so.Optimizer(so.OptConfig(kind=so.OPT_ADAM, lr=0.01), 2)
    # first Adam step: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1 -> step = lr * g/|g| (up to eps)
    np.testing.assert_allclose(adam.apply(th, g), th - 0.01 * np.sign(g), rtol=1e-4)
    ada = so.Optimizer(so.OptConfig(kind=so.OPT_ADADELTA, lr=1.0), 2)
    acc = 0.05 * g * g
    upd = np.sqrt(1e-8) / np.sqrt(acc + 1e-8) * g
    np.testing.assert_allclose(ada.apply(th, g), th - upd, rtol=1e-5)


def test_syncreplicas_state_machine_single_worker():
    """single worker, R = 3: the R initial tokens carry step 0, so the run that follows the FIRST update still
    holds a stale token and its push is dropped (one extra run in the second cycle); from then on the queue
    holds exactly the R fresh tokens of the last update and every push counts."""
    net = so.NetDesc(4, [3], [so.ACT_TANH])
    params = so.xavier_init(net, 1)
    R = 3
    tr = so.SyncReplicasTrainer(net, params, so.OptConfig(kind=so.OPT_SGD, lr=0.1), R)
    steps = []
    for i in range(12):
        X, y, w = so.synth_batch(5, 4, i)
        steps.append(tr.run(X, y, w)[1])
    assert steps == [0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3]


def test_split_batches_and_R():
    b = so.split_batches(1030, 100)
    assert len(b) == 10 and all(len(x) == 103 for x in b)
    b = so.split_batches(1099, 100)
    assert sorted({len(x) for x in b}) == [109, 110]
    assert so.replicas_to_aggregate(10000, 0.1, 100) == 90


def test_load_data_restatement(tmp_path):
    rows = ["1|0.5|-1.25|2.0|3", "0|1.5|0.25|-1.0|-2", "1|x|0.0|1.0|0.5"]
    p = tmp_path / "part-0.gz"
    with gzip.open(p, "wb") as f:
        f.write(("\n".join(rows) + "\n").encode())

    class Fixed:
        def __init__(self, seq): self.seq = list(seq)
        def random(self): return self.seq.pop(0)

    out = so.load_data([str(p)], [1, 2, 3], 0, 4, 0.2, rng=Fixed([0.9, 0.1, 0.5]))
    assert out["train_target"] == [[1.0], [1.0]] and out["valid_target"] == [[0.0]]
    assert out["train_data"][0] == [0.5, -1.25, 2.0]
    assert out["train_data"][1] == [0.0, 1.0]            # unparsable cell silently skipped (ssgd_monitor.py:409-411)
    assert out["train_data_sample_weight"] == [[3.0], [0.5]]
    assert out["valid_data_sample_weight"] == [[1.0]]     # negative weight -> 1.0 (:414-416)
    assert out["feature_count"] == 3


@pytest.mark.parametrize("optimizer", [so.OPT_ADADELTA, so.OPT_ADAM, so.OPT_SGD, so.OPT_MOMENTUM])
def test_torch_cpu_worker_equals_numpy_oracle(optimizer):
    """the CPU-baseline arm (bench.py cpu_baseline / --impl reference) computes the same thing as the oracle"""
    torch = pytest.importorskip("torch")
    from oracle.torch_cpu_worker import TorchCpuWorker
    net = so.NetDesc(24, [16, 8], [so.ACT_RELU, so.ACT_TANH])
    params = so.xavier_init(net, 5)
    cfg = so.OptConfig(kind=optimizer, lr=0.05)
    ref = so.CleanTrainer(net, params, cfg)
    wk = TorchCpuWorker(net, params, cfg, threads=2)
    for s in range(3):
        X, y, w = so.synth_batch(40, 24, s, weights="mixed")
        rl = ref.step([(X, y, w)])[0]
        gl = wk.step(torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(w))
        assert abs(gl - rl) < 1e-6
    assert np.abs(wk.flat_params() - ref.theta).max() < 2e-6


def test_oracle_matches_tf_golden():
    """When somebody has run oracle/tf_golden.py on a machine WITH TensorFlow and committed tests/golden/tf_golden.npz, the
    oracle is pinned to real TF output here (forward, both losses, all gradients, three steps of all four optimizers,
    including the ApplyAdadelta evaluation order).  Until then this test is skipped and parity stays 'unpinned'."""
    import os
    import pytest
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tf_golden.npz absent: no TensorFlow was available to write it (oracle/tf_golden.py)")
    G = np.load(path)
    net = so.NetDesc(int(G["F"]), [int(h) for h in G["hidden"]], [int(a) for a in G["acts"]])
    params = [G["param%d" % i] for i in range(2 * len(net.acts) + 2)]
    X, y, w = G["X"], G["y"], G["w"]
    kinds = {"adadelta": so.OptConfig(kind=so.OPT_ADADELTA, lr=0.5), "adam": so.OptConfig(kind=so.OPT_ADAM, lr=0.01),
             "sgd": so.OptConfig(kind=so.OPT_SGD, lr=0.1), "momentum": so.OptConfig(kind=so.OPT_MOMENTUM, lr=0.1, momentum=0.9)}
    for loss_name, loss_id in (("mse", so.LOSS_MSE), ("ce", so.LOSS_SIGMOID_CE)):
        for opt_name, cfg in kinds.items():
            key = "%s_%s_" % (loss_name, opt_name)
            L, g, yhat = so.loss_and_grads(net, params, X, y, w, loss_id)
            assert abs(float(L) - float(G[key + "loss"])) <= 1e-6
            assert np.abs(yhat - G[key + "yhat"]).max() <= 1e-6
            for i, gi in enumerate(g):
                assert np.abs(gi - G[key + "grad%d" % i].reshape(gi.shape)).max() <= 1e-6
            tr = so.CleanTrainer(net, params, cfg, loss=loss_id)
            for step in range(3):
                tr.step([(X, y, w)])
                for i, p in enumerate(tr.params()):
                    assert np.abs(p - G[key + "step%d_param%d" % (step + 1, i)].reshape(p.shape)).max() <= 2e-6, (key, step, i)
