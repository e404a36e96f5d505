"""Independent pins of the oracle: until a real TF can write golden vectors (oracle/tf_golden.py), the
restatement is checked against machinery that shares NO code with it:

  * gradients: torch.autograd through a forward written with torch ops (loss = sum w (yhat-y)^2 / #{w != 0},
    tf.losses.mean_squared_error's SUM_BY_NONZERO_WEIGHTS, ssgd_monitor.py:129; and the sigmoid-CE variant);
  * optimizer forms: torch.optim.{Adadelta, SGD(momentum), Adam} where torch's documented update coincides with the TF-1.x
    kernel (Adadelta, Momentum: identical algebra; Adam: identical once epsilon is negligible), plus closed-form one-step
    values for the place where TF's Adam differs from torch's (epsilon OUTSIDE the bias-corrected sqrt)."""
import numpy as np
import pytest
import torch

from oracle import shifu_oracle as so

ACTS = {so.ACT_SIGMOID: torch.sigmoid, so.ACT_TANH: torch.tanh, so.ACT_RELU: torch.relu,
        so.ACT_LEAKYRELU: lambda z: torch.nn.functional.leaky_relu(z, 0.2)}


def _torch_loss(net, params, X, y, w, loss):
    a = torch.from_numpy(X).double()
    for l, act in enumerate(net.acts):
        a = ACTS[act](a @ params[2 * l] + params[2 * l + 1])
    z = a @ params[-2] + params[-1]
    yt, wt = torch.from_numpy(y).double(), torch.from_numpy(w).double()
    nnz = int((w != 0).sum())
    if loss == so.LOSS_MSE:
        per = (torch.sigmoid(z) - yt) ** 2
    else:
        per = torch.nn.functional.binary_cross_entropy_with_logits(z, yt, reduction="none")
    return (per * wt).sum() / nnz


@pytest.mark.parametrize("loss", [so.LOSS_MSE, so.LOSS_SIGMOID_CE])
@pytest.mark.parametrize("acts", [[so.ACT_RELU, so.ACT_TANH], [so.ACT_SIGMOID, so.ACT_LEAKYRELU], [so.ACT_LEAKYRELU]])
def test_backward_equals_torch_autograd(loss, acts):
    hidden = [13, 7][:len(acts)]
    net = so.NetDesc(11, hidden, acts)
    params = [p.astype(np.float64) for p in so.xavier_init(net, 5)]
    X, y, w = so.synth_batch(37, 11, 9, weights="mixed")
    L, g, _ = so.loss_and_grads(net, params, X.astype(np.float64), y.astype(np.float64), w.astype(np.float64), loss)
    tp = [torch.from_numpy(p.copy()).requires_grad_(True) for p in params]
    Lt = _torch_loss(net, tp, X, y, w, loss)
    Lt.backward()
    assert abs(float(Lt) - float(L)) <= 1e-12
    for a, b in zip(g, tp):
        assert np.abs(a - b.grad.numpy().reshape(a.shape)).max() <= 1e-12


def _run_oracle(kind, steps, grads, theta0, **kw):
    opt = so.Optimizer(so.OptConfig(kind=kind, **kw), theta0.size, np.float64)
    th = theta0.copy()
    for g in grads[:steps]:
        th = opt.apply(th, g)
    return th


def _run_torch(make, steps, grads, theta0):
    p = torch.from_numpy(theta0.copy()).requires_grad_(True)
    opt = make([p])
    for g in grads[:steps]:
        p.grad = torch.from_numpy(g.copy())
        opt.step()
    return p.detach().numpy()


def test_optimizer_forms_equal_torch_where_the_algebra_coincides():
    rng = np.random.RandomState(0)
    theta0 = rng.randn(50)
    grads = [rng.randn(50) * 0.1 for _ in range(7)]
    # ApplyAdadelta == torch.optim.Adadelta (same accumulators, same epsilon placement)
    a = _run_oracle(so.OPT_ADADELTA, 7, grads, theta0, lr=0.7, rho=0.9, eps=1e-6)
    b = _run_torch(lambda ps: torch.optim.Adadelta(ps, lr=0.7, rho=0.9, eps=1e-6), 7, grads, theta0)
    assert np.abs(a - b).max() <= 1e-12
    # ApplyMomentum (use_nesterov=False) == torch SGD with momentum, dampening 0
    a = _run_oracle(so.OPT_MOMENTUM, 7, grads, theta0, lr=0.05, momentum=0.8)
    b = _run_torch(lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.8), 7, grads, theta0)
    assert np.abs(a - b).max() <= 1e-12
    a = _run_oracle(so.OPT_SGD, 7, grads, theta0, lr=0.05)
    b = _run_torch(lambda ps: torch.optim.SGD(ps, lr=0.05), 7, grads, theta0)
    assert np.abs(a - b).max() <= 1e-12
    # ApplyAdam: with a negligible epsilon TF's and torch's forms are the same algebra (bias corrections, moments)
    a = _run_oracle(so.OPT_ADAM, 7, grads, theta0, lr=0.01, beta1=0.8, beta2=0.95, eps=1e-300)
    b = _run_torch(lambda ps: torch.optim.Adam(ps, lr=0.01, betas=(0.8, 0.95), eps=1e-300), 7, grads, theta0)
    assert np.abs(a - b).max() <= 1e-10


def test_tf_adam_epsilon_placement_closed_form():
    """TF: theta -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps); at t = 1 that is lr * g / (|g| + eps / sqrt(1-b2)),
    torch would give lr * g / (|g| + eps).  The oracle must follow TF."""
    g = np.array([1e-3, -2e-6, 5e-9, 0.0])
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    th = so.Optimizer(so.OptConfig(kind=so.OPT_ADAM, lr=lr, beta1=b1, beta2=b2, eps=eps), 4, np.float64).apply(np.zeros(4), g)
    want = -lr * g / (np.abs(g) + eps / np.sqrt(1 - b2))
    assert np.abs(th - want).max() <= 1e-15
    torch_form = -lr * g / (np.abs(g) + eps)
    assert np.abs(th - torch_form).max() > 1e-4     # the two forms really differ for small gradients


def test_adadelta_first_step_closed_form():
    """one ApplyAdadelta step from zero state: accum = (1-rho) g^2; update = sqrt(eps)/sqrt(accum+eps) g; var -= lr update"""
    g = np.array([0.3, -1e-4, 2.0])
    lr, rho, eps = 1.0, 0.95, 1e-8
    th = so.Optimizer(so.OptConfig(kind=so.OPT_ADADELTA, lr=lr, rho=rho, eps=eps), 3, np.float64).apply(np.zeros(3), g)
    want = -lr * np.sqrt(eps) / np.sqrt((1 - rho) * g * g + eps) * g
    assert np.abs(th - want).max() <= 1e-15
