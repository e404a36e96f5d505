"""The wide+deep spec (oracle/wide_deep.py): evaluating the first layer sparsely over one-hot indices is the dense
reference layer on the materialised one-hot columns - loss, every gradient and an SGD step agree."""
import numpy as np
import pytest

from oracle import shifu_oracle as so
from oracle import wide_deep as wd


@pytest.mark.parametrize("loss", [so.LOSS_MSE, so.LOSS_SIGMOID_CE])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_first_layer_equals_dense_on_onehot_columns(loss, dtype):
    n_dense, vocab = 7, [5, 1, 300, 3]
    n_onehot = sum(vocab)
    net = so.NetDesc(n_dense + n_onehot, [16, 8], [so.ACT_LEAKYRELU, so.ACT_TANH])
    params = [p.astype(dtype) for p in so.xavier_init(net, 3)]
    Xd, idx, y, w = wd.synth_wide_deep_batch(257, n_dense, vocab, 9, missing=0.2)
    Xd, y, w = Xd.astype(dtype), y.astype(dtype), w.astype(dtype)
    X = np.concatenate([Xd, wd.onehot_matrix(idx, n_onehot, dtype)], axis=1)
    assert set(np.unique(X[:, n_dense:])) <= {0.0, 1.0} and (idx == -1).any()
    Ld, gd, yd = so.loss_and_grads(net, params, X, y, w, loss)
    Ls, gs, ys = wd.loss_and_grads_sparse(net, params, Xd, idx, y, w, loss)
    tol = 1e-12 if dtype == np.float64 else 2e-6
    assert abs(Ld - Ls) <= tol
    np.testing.assert_allclose(ys, yd, rtol=0, atol=tol)
    for a, b in zip(gs, gd):
        np.testing.assert_allclose(a, b, rtol=0, atol=tol)
    # embedding rows nobody indexed get exactly zero gradient
    untouched = np.setdiff1d(np.arange(n_onehot), idx[idx >= 0])
    assert len(untouched) and not gs[0][n_dense + untouched].any()


def test_repeated_index_counts_twice_and_missing_counts_zero():
    idx = np.array([[2, 2, -1], [0, -1, -1]])
    M = wd.onehot_matrix(idx, 4)
    np.testing.assert_array_equal(M, [[0, 0, 2, 0], [1, 0, 0, 0]])
    net = so.NetDesc(1 + 4, [3], [so.ACT_RELU])
    params = [p.astype(np.float64) for p in so.xavier_init(net, 1)]
    Xd = np.array([[0.5], [-1.0]]); y = np.array([[1.0], [0.0]]); w = np.ones((2, 1))
    Ld, gd, _ = so.loss_and_grads(net, params, np.concatenate([Xd, M], 1), y, w)
    Ls, gs, _ = wd.loss_and_grads_sparse(net, params, Xd, idx, y, w)
    assert abs(Ld - Ls) < 1e-14
    for a, b in zip(gs, gd):
        np.testing.assert_allclose(a, b, atol=1e-14)
