"""N > 1 host logic on CPU: world_size-2 gloo process group (no GPU).  Covers the id broadcast bench.py / the Java host
rely on, the max-over-ranks timing reduction, row / file sharding, and - with the oracle standing in for the ranks' math -
that 'mean over ranks of per-rank gradients' is what the data-parallel step applies."""
import os
import socket

import numpy as np
import pytest

from oracle import shifu_oracle as so

torch = pytest.importorskip("torch")
import torch.distributed as dist            # noqa: E402
import torch.multiprocessing as mp          # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import shifu_tensorflow_b200  # noqa: F401
    from shifu_tensorflow_b200 import dist_util as du
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = du.broadcast_bytes(dist, lambda: bytes((7 * i + 3) % 256 for i in range(128)), 128, rank)
    mx = du.max_over_ranks(dist, 10.0 + rank, world)
    # data-parallel step with the oracle as each rank's math: own shard, own n_nz, gradient MEAN over ranks
    net = so.NetDesc(10, [6], [so.ACT_TANH])
    params = so.xavier_init(net, 4)
    X, y, w = so.synth_batch(64, 10, 9, weights="mixed")
    idx = du.shard_rows(64, rank, world)
    L, g, _ = so.loss_and_grads(net, params, X[idx], y[idx], w[idx])
    g = torch.from_numpy(so.flatten_params(g).copy())
    dist.all_reduce(g); g /= world
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), uid=np.frombuffer(uid, np.uint8), mx=mx, g=g.numpy(), loss=L, idx=idx)
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(str(tmp_path / ("r%d.npz" % i))) for i in range(world)]
    want_uid = np.array([(7 * i + 3) % 256 for i in range(128)], np.uint8)
    for x in r:
        np.testing.assert_array_equal(x["uid"], want_uid)       # every rank holds rank 0's id
        assert float(x["mx"]) == 11.0                            # max over ranks
    np.testing.assert_array_equal(r[0]["g"], r[1]["g"])          # identical averaged gradient everywhere
    assert sorted(np.concatenate([r[0]["idx"], r[1]["idx"]]).tolist()) == list(range(64))
    # and it equals the oracle's own data-parallel restatement
    net = so.NetDesc(10, [6], [so.ACT_TANH]); params = so.xavier_init(net, 4)
    X, y, w = so.synth_batch(64, 10, 9, weights="mixed")
    ref = so.CleanTrainer(net, params, so.OptConfig(kind=so.OPT_SGD, lr=0.0))
    ref.step([(X[i::2], y[i::2], w[i::2]) for i in range(2)])
    assert np.abs(ref.last_grads - r[0]["g"]).max() <= 1e-6


def test_shard_helpers():
    import shifu_tensorflow_b200  # noqa: F401
    from shifu_tensorflow_b200 import dist_util as du
    assert du.shard_rows(10, 1, 4).tolist() == [1, 5, 9]
    assert du.shard_files(["a", "b", "c", "d", "e"], 1, 2) == ["b", "d"]
    assert du.max_over_ranks(None, 3.5, 1) == 3.5
