"""Micro-benchmark of the gradient exchange alone (run under torchrun, N ranks): times K back-to-back exchanges of the
cfg1 / cfg2 flat gradient through the profile hook is not possible, so this uses tiny nets with the same parameter count
... simpler: run real steps with tiny batches so that compute is negligible and the step time ~ exchange + optimizer."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import shifu_tensorflow_b200 as sb
from shifu_tensorflow_b200 import dist_util

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
for name, (F, hidden) in {"cfg1": (1000, [512, 256, 128]), "cfg2": (2000, [1024, 512, 256])}.items():
    for exch in ("nccl", "p2p"):
        B = 128
        uid = dist_util.broadcast_bytes(dist, sb.capi.nccl_unique_id, 128, rank, device="cuda")  # one id per communicator
        desc = sb.make_desc(F, hidden, [2, 2, 2], optimizer=sb.OPT_SGD, learning_rate=0.0, max_batch=B, precision=sb.PREC_BF16)
        t = sb.Trainer(desc, device=local, nccl_id=uid, rank=rank, world=world)
        if exch == "p2p":
            dist_util.enable_peer_exchange(dist, t, world, device="cuda")
        t.init_xavier(1)
        X = np.random.default_rng(rank).standard_normal((B * 4, F), dtype=np.float32)
        t.load_dataset(X, np.zeros(B * 4, np.float32), None)
        for i in range(10):
            t.step_resident_async(0, B)
        t.sync(); dist.barrier(); torch.cuda.synchronize()
        st = torch.cuda.ExternalStream(t.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(100):
            t.step_resident_async((i % 4) * B, B)
        e1.record(st); t.sync()
        res[name + "_" + exch] = e0.elapsed_time(e1) / 100 * 1e3
        t.close(); dist.barrier()
if rank == 0:
    print(json.dumps({"n_gpus": world, "us_per_tiny_step (128 rows: ~ fixed step latency + exchange + optimizer)": res}))
dist.destroy_process_group()
