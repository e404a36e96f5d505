"""TEST INFRASTRUCTURE ONLY - spec + CPU oracle for the wide+deep configuration (BASELINE.json configs[3], SURVEY.md 8f
rank 3).  Nothing under shifu-tensorflow_b200/ imports this file.  PARITY UNPINNED: the reference has no wide+deep
implementation at all; what IS pinned is the equivalence below.

Spec.  Shifu's one-hot normalisation turns C categorical columns with vocabularies V_1..V_C into sum(V_c) = n_onehot
0/1 columns of which exactly one per categorical column (or none, for a missing value) is 1 in every row.  Fed to the
reference worker those are ordinary numeric columns of the first dense layer (res/ssgd_monitor.py:57-71):

    Z_0 = [X_dense | X_onehot] W_0 + b_0          W_0 = [[W_d]   n_dense  rows
                                                         [W_e]]  n_onehot rows  ("embedding table", width h_0)

so "embedding -> MLP" is not a different model but the SAME first layer evaluated sparsely:

    forward     Z_0[r] = X_dense[r] W_d + sum_c W_e[idx[r, c]] + b_0              idx[r, c] in [0, n_onehot) or -1
    backward    dW_d   = X_dense^T dZ_0          dW_e[j] = sum over (r, c) with idx[r, c] == j of dZ_0[r]
                (no dA for layer 0);  every other layer, the loss and the optimizer are unchanged.

The oracle of the sparse path is therefore the dense oracle (shifu_oracle.loss_and_grads) on the materialised one-hot
matrix; this file provides the sparse restatement a CUDA path must match and the checks that the two agree."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import shifu_oracle as so


def onehot_matrix(idx: np.ndarray, n_onehot: int, dtype=np.float32) -> np.ndarray:
    """idx [rows, C] (global one-hot column index per categorical column, -1 = missing) -> 0/1 matrix [rows, n_onehot]"""
    idx = np.asarray(idx)
    rows, C = idx.shape
    M = np.zeros((rows, n_onehot), dtype)
    r = np.repeat(np.arange(rows), C)
    j = idx.reshape(-1)
    keep = j >= 0
    np.add.at(M, (r[keep], j[keep]), 1)      # add, not assign: a repeated index counts twice, as two 1-columns would
    return M


def forward_sparse(net: so.NetDesc, params: Sequence[np.ndarray], Xd: np.ndarray, idx: np.ndarray):
    """same return convention as shifu_oracle.forward; A[0] is the dense part only"""
    n_dense = Xd.shape[1]
    W0, b0 = params[0], params[1]
    Wd, We = W0[:n_dense], W0[n_dense:]
    z0 = Xd @ Wd + b0
    for c in range(idx.shape[1]):
        j = idx[:, c]
        hit = j >= 0
        z0[hit] += We[j[hit]]
    A = [Xd, so.act_forward(z0, net.acts[0])]
    for l in range(1, len(net.acts)):
        A.append(so.act_forward(A[-1] @ params[2 * l] + params[2 * l + 1], net.acts[l]))
    z = A[-1] @ params[-2] + params[-1]
    return A, z, so._sigmoid(z)


def loss_and_grads_sparse(net: so.NetDesc, params, Xd, idx, y, w, loss=so.LOSS_MSE) -> Tuple[float, List[np.ndarray], np.ndarray]:
    A, z, yhat = forward_sparse(net, params, Xd, idx)
    L, n_nz = so.loss_value(z, yhat, y, w, loss)
    grads = [np.zeros_like(p) for p in params]
    if n_nz == 0:
        return L, grads, yhat
    dt = z.dtype.type
    dz = dt(2) * w * (yhat - y) * yhat * (dt(1) - yhat) / dt(n_nz) if loss == so.LOSS_MSE else w * (yhat - y) / dt(n_nz)
    grads[-2] = A[-1].T @ dz
    grads[-1] = dz.sum(axis=0)
    dA = dz @ params[-2].T
    n_dense = Xd.shape[1]
    for l in range(len(net.acts) - 1, -1, -1):
        dZ = dA * so.act_grad_from_output(A[l + 1], net.acts[l])
        grads[2 * l + 1] = dZ.sum(axis=0)
        if l > 0:
            grads[2 * l] = A[l].T @ dZ
            dA = dZ @ params[2 * l].T
        else:
            g0 = np.zeros_like(params[0])
            g0[:n_dense] = Xd.T @ dZ
            for c in range(idx.shape[1]):                      # scatter-add of dZ_0 rows into the embedding rows
                j = idx[:, c]
                hit = j >= 0
                np.add.at(g0, (n_dense + j[hit],), dZ[hit])
            grads[0] = g0
    return L, grads, yhat


def synth_wide_deep_batch(rows: int, n_dense: int, vocab: Sequence[int], seed: int, missing: float = 0.05):
    """dense part as shifu_oracle.synth_batch; one index per categorical column (Zipf-ish), `missing` of them -1"""
    X, y, w = so.synth_batch(rows, n_dense, seed, weights="mixed")
    rng = np.random.RandomState(seed + 1)
    offs = np.concatenate([[0], np.cumsum(vocab)[:-1]])
    idx = np.empty((rows, len(vocab)), np.int64)
    for c, (V, o) in enumerate(zip(vocab, offs)):
        p = 1.0 / np.arange(1, V + 1); p /= p.sum()
        idx[:, c] = o + rng.choice(V, size=rows, p=p)
    idx[rng.uniform(size=idx.shape) < missing] = -1
    return X, idx, y, w
