"""debug aid: accumulate + apply over the in-process exchange, prints what each rank holds at every stage"""
import os, sys, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ["SB_XCHG_BLOCKS"] = "8"; os.environ["SB_XCHG_TIMEOUT_S"] = "20"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import shifu_tensorflow_b200 as sb
from oracle import shifu_oracle as so
import importlib.util
spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests/test_data_parallel_one_gpu.py")); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
F, hidden, acts, B = 64, [48, 24], [so.ACT_TANH, so.ACT_RELU], 96
net = so.NetDesc(F, hidden, acts); params = so.xavier_init(net, 4)
desc = sb.make_desc(F, hidden, acts, optimizer=so.OPT_SGD, learning_rate=1.0, max_batch=B, precision=0)
ts = [sb.Trainer(desc, device=0, nccl_id=None, rank=r, world=2) for r in range(2)]
for t in ts:
    t.set_peer_pointers([x.exchange_base for x in ts]); t.set_params(so.flatten_params(params))
shards = m._shards(2, 3, B, F, 5)
for t, (X, y, w) in zip(ts, shards):
    t.load_dataset(X, y, w)
want = {}
for r, n_acc in ((0, 3), (1, 2)):
    X, y, w = shards[r]
    for k in range(n_acc):
        ts[r].accumulate_resident(k * B, B)
        g = so.flatten_params(so.loss_and_grads(net, params, X[k*B:(k+1)*B], y[k*B:(k+1)*B], w[k*B:(k+1)*B])[1])
        got = ts[r].get_grads()
        print("rank", r, "batch", k, "per-step grad err", np.abs(got - g).max(), "elems", got[[0, 1, -1]], g[[0, 1, -1]])
        want[(r, k)] = g
import time
t0 = time.time()
for t in ts:
    t.apply_accumulated(5)
for t in ts:
    t.sync()
print("apply took %.2f s" % (time.time() - t0))
tot = sum(want.values())
for r in range(2):
    g = ts[r].get_grads()
    print("rank", r, "applied mean err", np.abs(g - tot / 5).max(), g[[0, 1, 2000, 3000, -1]], (tot / 5)[[0, 1, 2000, 3000, -1]])
