"""2-GPU data-parallel parity (needs `gpurun --gpus 2`): two processes, NCCL inside libshifu_b200.so, each rank steps on its
own shard; parameters after the steps must equal the oracle's data-parallel trainer (mean of per-rank gradients)."""
import os
import socket

import numpy as np
import pytest

from oracle import shifu_oracle as so

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, out_dir, precision, exchange="nccl"):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import shifu_tensorflow_b200 as sb
    from shifu_tensorflow_b200 import dist_util as du
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = du.broadcast_bytes(dist, sb.capi.nccl_unique_id, sb.capi.SB_NCCL_ID_BYTES, rank)
    net = so.NetDesc(96, [64, 32], [so.ACT_RELU, so.ACT_TANH])
    params = so.xavier_init(net, 4)
    desc = sb.make_desc(96, [64, 32], net.acts, optimizer=so.OPT_MOMENTUM, learning_rate=0.1, max_batch=128, precision=precision)
    t = sb.Trainer(desc, device=rank, nccl_id=uid, rank=rank, world=world)
    if exchange == "p2p":
        du.enable_peer_exchange(dist, t, world)
    t.set_params(so.flatten_params(params))
    losses = []
    for s in range(3):
        X, y, w = so.synth_batch(256, 96, 20 + s, weights="mixed")
        idx = du.shard_rows(256, rank, world)
        losses.append(t.step(X[idx], y[idx], w[idx]))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), theta=t.get_params(), grads=t.get_grads(), losses=np.array(losses))
    t.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("precision", [0, 1])
def test_two_gpu_data_parallel_matches_oracle(sb, tmp_path, precision, exchange):
    if sb.capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path), precision, exchange), nprocs=world, join=True)
    r = [np.load(str(tmp_path / ("r%d.npz" % i))) for i in range(world)]
    np.testing.assert_array_equal(r[0]["theta"], r[1]["theta"])      # replicas stay bit-identical
    np.testing.assert_array_equal(r[0]["grads"], r[1]["grads"])
    net = so.NetDesc(96, [64, 32], [so.ACT_RELU, so.ACT_TANH])
    ref = so.CleanTrainer(net, so.xavier_init(net, 4), so.OptConfig(kind=so.OPT_MOMENTUM, lr=0.1))
    for s in range(3):
        X, y, w = so.synth_batch(256, 96, 20 + s, weights="mixed")
        want = ref.step([(X[i::2], y[i::2], w[i::2]) for i in range(2)])
        if precision == 0:
            assert abs(want[0] - r[0]["losses"][s]) <= 1e-5 and abs(want[1] - r[1]["losses"][s]) <= 1e-5
    tol = 1e-5 if precision == 0 else 5e-3
    assert np.abs(r[0]["theta"] - ref.theta).max() <= tol
