"""Reference-equivalent CPU worker: the oracle's train step restated on torch-CPU fp32 tensors so that it uses all
host cores (MKL/oneDNN GEMM), i.e. the fastest honest stand-in for the reference's TF-CPU worker loop
(res/ssgd_monitor.py:268-277), which cannot run here (Python 2 + TF 1.x, neither installable).

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline leg and `--impl reference`, and checked against
oracle/shifu_oracle.py in tests/test_oracle.py.  Same math, same flat parameter order, same optimizer forms.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import shifu_oracle as so


def _act(z, act):
    if act == so.ACT_SIGMOID:
        return torch.sigmoid(z)
    if act == so.ACT_TANH:
        return torch.tanh(z)
    if act == so.ACT_RELU:
        return torch.relu(z)
    return torch.where(z > 0, z, z * so.LEAKY_ALPHA)


def _act_grad(a, act):
    if act == so.ACT_SIGMOID:
        return a * (1 - a)
    if act == so.ACT_TANH:
        return 1 - a * a
    if act == so.ACT_RELU:
        return (a > 0).to(a.dtype)
    return torch.where(a > 0, torch.ones_like(a), torch.full_like(a, so.LEAKY_ALPHA))


class TorchCpuWorker:
    def __init__(self, net: so.NetDesc, params, opt: so.OptConfig, loss=so.LOSS_MSE, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.net, self.cfg, self.loss = net, opt, loss
        self.P = [torch.from_numpy(np.array(p, dtype=np.float32, copy=True)) for p in params]
        self.s1 = [torch.zeros_like(p) for p in self.P]
        self.s2 = [torch.zeros_like(p) for p in self.P]
        self.t = 0

    def flat_params(self) -> np.ndarray:
        return np.concatenate([p.numpy().ravel() for p in self.P])

    def step(self, X, y, w):
        """one sess.run([train_step, loss]) in the clean schedule; X [B,F], y [B,1], w [B,1] torch fp32."""
        net, P = self.net, self.P
        A = [X]
        for l, act in enumerate(net.acts):
            A.append(_act(torch.addmm(P[2 * l + 1], A[-1], P[2 * l]), act))
        z = torch.addmm(P[-1], A[-1], P[-2])
        yhat = torch.sigmoid(z)
        n_nz = int(torch.count_nonzero(w))
        if n_nz == 0:
            return 0.0
        if self.loss == so.LOSS_MSE:
            d = yhat - y
            L = float((w * d * d).sum() / n_nz)
            dz = 2 * w * d * yhat * (1 - yhat) / n_nz
        else:
            L = float((w * (torch.clamp(z, min=0) - z * y + torch.log1p(torch.exp(-z.abs())))).sum() / n_nz)
            dz = w * (yhat - y) / n_nz
        G = [None] * len(P)
        G[-2] = A[-1].t() @ dz
        G[-1] = dz.sum(0)
        dA = dz @ P[-2].t()
        for l in range(len(net.acts) - 1, -1, -1):
            dZ = dA * _act_grad(A[l + 1], net.acts[l])
            G[2 * l] = A[l].t() @ dZ
            G[2 * l + 1] = dZ.sum(0)
            if l > 0:
                dA = dZ @ P[2 * l].t()
        self._apply(G)
        return L

    def _apply(self, G):
        c = self.cfg
        self.t += 1
        for i, (p, g) in enumerate(zip(self.P, G)):
            if c.kind == so.OPT_SGD:
                p.sub_(g, alpha=c.lr)
            elif c.kind == so.OPT_MOMENTUM:
                self.s1[i].mul_(c.momentum).add_(g)
                p.sub_(self.s1[i], alpha=c.lr)
            elif c.kind == so.OPT_ADAM:
                lr_t = c.lr * math.sqrt(1 - c.beta2 ** self.t) / (1 - c.beta1 ** self.t)
                self.s1[i].add_((g - self.s1[i]) * (1 - c.beta1))
                self.s2[i].add_((g * g - self.s2[i]) * (1 - c.beta2))
                p.sub_(lr_t * self.s1[i] / (self.s2[i].sqrt() + c.eps))
            else:
                self.s1[i].mul_(c.rho).add_(g * g, alpha=1 - c.rho)
                upd = (self.s2[i] + c.eps).sqrt() / (self.s1[i] + c.eps).sqrt() * g
                self.s2[i].mul_(c.rho).add_(upd * upd, alpha=1 - c.rho)
                p.sub_(upd, alpha=c.lr)

    def score(self, X):
        A = X
        for l, act in enumerate(self.net.acts):
            A = _act(torch.addmm(self.P[2 * l + 1], A, self.P[2 * l]), act)
        return torch.sigmoid(torch.addmm(self.P[-1], A, self.P[-2]))


def time_train(net, params, opt, batches, min_seconds=10.0, max_steps=50, threads=None, loss=so.LOSS_MSE):
    """Times the batch loop only, like the reference (res/ssgd_monitor.py:270-277).  batches: list of (X,y,w) numpy.
    -> dict(rows_per_sec, steps, seconds, cores)"""
    w = TorchCpuWorker(net, params, opt, loss, threads)
    tb = [(torch.from_numpy(X), torch.from_numpy(y.reshape(-1, 1)), torch.from_numpy(wt.reshape(-1, 1))) for X, y, wt in batches]
    w.step(*tb[0])  # warm-up (thread pool, allocator)
    rows, steps, t0 = 0, 0, time.perf_counter()
    while True:
        X, y, wt = tb[steps % len(tb)]
        w.step(X, y, wt)
        rows += X.shape[0]; steps += 1
        el = time.perf_counter() - t0
        if steps >= max_steps or el >= min_seconds:
            break
    return {"rows_per_sec": rows / el, "steps": steps, "seconds": el, "cores": torch.get_num_threads()}
