"""Shared helpers for the parity tests: build the same network on both sides (oracle / C-ABI)."""
import numpy as np

from oracle import shifu_oracle as so


def make_pair(sb, n_features, hidden, acts, *, loss=0, optimizer=0, lr=0.05, max_batch=128, precision=0, seed=7,
              dtype=np.float32, **hyper):
    """-> (oracle NetDesc, oracle params list, oracle OptConfig, C-ABI desc)"""
    net = so.NetDesc(n_features, list(hidden), list(acts))
    params = so.xavier_init(net, seed)
    cfg = so.OptConfig(kind=optimizer, lr=lr, **{k: v for k, v in hyper.items() if k in ("rho", "eps", "beta1", "beta2", "momentum")})
    desc = sb.make_desc(n_features, hidden, acts, loss=loss, optimizer=optimizer, learning_rate=lr,
                        rho=cfg.rho, epsilon=cfg.eps, beta1=cfg.beta1, beta2=cfg.beta2, momentum=cfg.momentum,
                        max_batch=max_batch, precision=precision)
    return net, params, cfg, desc


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())
