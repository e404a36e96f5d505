#!/usr/bin/env python
"""bench.py - rows/sec of the tabular-DNN train step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1|cfg2|cfg0] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one mini-batch: load -> forward -> loss -> backward -> gradient mean
over ranks (peer-memory all-reduce kernel, NCCL as fallback) -> optimizer update.  `value` times steps whose
mini-batches are already resident in HBM (sb_trainer_run_resident = the per-epoch batch loop in one call, CUDA events on
the trainer's stream, max over ranks); `e2e` times the same step through the public C-ABI call with HOST (pinned)
buffers, H2D of every batch and the read-back of the loss scalars inside the timed region.  Weak scaling: every rank owns its own `batch` rows per step.  PyTorch is used for plumbing only
(rendezvous, barrier, max-reduce, events); all compute is libshifu_b200.so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260921
# BASELINE.json configs.  Activation: relu (ModelConfig ActivationFunc); loss: the reference's MSE-on-sigmoid
# (res/ssgd_monitor.py:129) - same cost as the sigmoid-CE variant.
CONFIGS = {
    "cfg0": dict(F=200, hidden=[100, 50], batch=100, optimizer="adadelta", lr=1.0, n_batches=90),
    "cfg1": dict(F=1000, hidden=[512, 256, 128], batch=4096, optimizer="adam", lr=0.001, n_batches=64),
    "cfg2": dict(F=2000, hidden=[1024, 512, 256], batch=8192, optimizer="momentum", lr=0.01, n_batches=32),
}
OPT_ID = {"adadelta": 0, "adam": 1, "sgd": 2, "momentum": 3}


def flops_per_row(F, hidden):
    """F_train = 6*sum(W) - 2*W_1; hidden-GEMM share drops the out=1 layer (6*h_L)."""
    dims = [F] + list(hidden) + [1]
    sw = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    f_train = 6 * sw - 2 * dims[0] * dims[1]
    return f_train, f_train - 6 * hidden[-1], 2 * sw


def synth_dataset(cfg, rank, n_batches=None):
    """X~N(0,1) clipped +-4 fp32 row-major, y~Bernoulli(0.2), w=1; per-rank seed."""
    nb = n_batches or cfg["n_batches"]
    rows = nb * cfg["batch"]
    rng = np.random.default_rng(SEED + 1000 * rank)
    X = rng.standard_normal((rows, cfg["F"]), dtype=np.float32)
    np.clip(X, -4, 4, out=X)
    y = (rng.random(rows, dtype=np.float32) < 0.2).astype(np.float32)
    w = np.ones(rows, np.float32)
    return X, y, w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """threads the CPU arm may really use: affinity mask and cgroup quota, not the box's core count"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p))))
    except Exception:
        pass
    return n


def run_reference(args, cfg, rank, world):
    """`--impl reference`: the reference-equivalent CPU worker (oracle port on torch-CPU, all host threads) on the same
    config / metric.  TF 1.x + Python 2 cannot be installed here, so the oracle port IS the CPU arm (kind "port").
    Under torchrun only rank 0 works."""
    if rank != 0:
        return
    import torch
    from oracle import shifu_oracle as so
    from oracle.torch_cpu_worker import TorchCpuWorker
    threads = host_threads()
    torch.set_num_threads(threads)
    net = so.NetDesc(cfg["F"], cfg["hidden"], [so.ACT_RELU] * len(cfg["hidden"]))
    params = so.xavier_init(net, SEED % 100000)
    opt = so.OptConfig(kind=OPT_ID[cfg["optimizer"]], lr=cfg["lr"])
    nb = min(4, cfg["n_batches"])
    X, y, w = synth_dataset(cfg, 0, nb)
    B = cfg["batch"]
    tb = [(torch.from_numpy(X[i * B:(i + 1) * B]), torch.from_numpy(y[i * B:(i + 1) * B].reshape(-1, 1)),
           torch.from_numpy(w[i * B:(i + 1) * B].reshape(-1, 1))) for i in range(nb)]
    wk = TorchCpuWorker(net, params, opt)
    for i in range(args.warmup):
        wk.step(*tb[i % nb])
    t0 = time.perf_counter()
    for i in range(args.steps):
        wk.step(*tb[i % nb])
    el = time.perf_counter() - t0
    val = args.steps * B / el
    out = {
        "impl": "reference", "metric": "rows/sec tabular-DNN train", "value": val, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.config, cfg, 1),
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of %s (batch %d) on torch-CPU fp32, batch loop only" % (args.steps, args.config, B)},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def config_block(name, cfg, world):
    return {"workload": "%s: %d cols x %d rows/GPU/step, MLP %s relu, %s, MSE-on-sigmoid loss" %
                        (name, cfg["F"], cfg["batch"], cfg["hidden"], cfg["optimizer"]),
            "global_batch": cfg["batch"] * world, "rows_per_gpu": cfg["batch"], "parallelism": "dp%d" % world,
            "resident_set": "%d batches (%.0f MB fp32 per GPU) cycled, larger than the 126 MB L2" %
                            (cfg["n_batches"], cfg["n_batches"] * cfg["batch"] * cfg["F"] * 4 / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="cfg1", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps, 50))")
    ap.add_argument("--also", default="cfg2", help="second config measured on the resident leg only and reported under 'also' ('' = none)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    import shifu_tensorflow_b200 as sb

    torch.cuda.set_device(local_rank)
    from shifu_tensorflow_b200 import dist_util
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tns = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    prec = sb.PREC_BF16 if args.precision == "bf16" else sb.PREC_FP32
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if peaks else "fallback 1.59 PF (B200_PROFILING.md)"
    # DRAM bytes of the GEMM launches of one step, from the committed `ncu --set full` capture (dram__bytes_read+write
    # summed over the 8 gemm_tc launches); algorithmic bytes (bf16 operands once + fp32 gradient) beside it
    NCU_TRAFFIC = {"cfg1": 50.7e6, "cfg2": 206.5e6}

    def measure(name, full):
        c = CONFIGS[name]
        B, F, hidden = c["batch"], c["F"], c["hidden"]
        nb = c["n_batches"] if full else min(c["n_batches"], 16)
        uid = None
        if world > 1:
            uid = dist_util.broadcast_bytes(dist, sb.capi.nccl_unique_id, sb.capi.SB_NCCL_ID_BYTES, rank, device="cuda")
        desc = sb.make_desc(F, hidden, [sb.ACT_RELU] * len(hidden), loss=sb.LOSS_MSE, optimizer=OPT_ID[c["optimizer"]],
                            learning_rate=c["lr"], max_batch=B, precision=prec)
        t = sb.Trainer(desc, device=local_rank, nccl_id=uid, rank=rank, world=world)
        exchange = "none"
        if world > 1:
            exchange = "nccl"
            if os.environ.get("SB_EXCHANGE", "p2p") == "p2p":
                dist_util.enable_peer_exchange(dist, t, world, device="cuda")
                exchange = "p2p (two-shot all-reduce kernel over CUDA-IPC peer memory)"
        t.init_xavier(SEED)  # same seed on every rank -> identical replicas
        X, y, w = synth_dataset(c, rank, nb)
        t.load_dataset(X, y, w)
        stream = torch.cuda.ExternalStream(t.stream, device=torch.device("cuda", local_rank))

        # ---------------- device-resident leg (value) ----------------
        # the per-epoch batch loop as ONE call (sb_trainer_run_resident: four steps per captured graph);
        # SB_BENCH_PER_STEP=1: one call per step (sb_trainer_step_resident_async) instead
        per_step_calls = os.environ.get("SB_BENCH_PER_STEP") == "1"
        if per_step_calls:
            for i in range(args.warmup):
                t.step_resident_async((i % nb) * B, B)
        else:
            # at least two chunks of four steps, so that both captured multi-step graphs exist before the timed region
            t.run_resident([(i % nb) * B for i in range(max(args.warmup, 8))], B)
            for i in range(2):      # and the two single-step graphs a step count that is not a multiple of four ends with
                t.step_resident_async((i % nb) * B, B)
        t.sync()
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0 and full:
            clocks.start()
            time.sleep(0.25)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.perf_counter()
        ev0.record(stream)
        if per_step_calls:
            for i in range(args.steps):
                t.step_resident_async(((args.warmup + i) % nb) * B, B)
        else:
            t.run_resident([((args.warmup + i) % nb) * B for i in range(args.steps)], B)
        ev1.record(stream)
        t.sync()
        barrier()
        wall1 = time.perf_counter()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        clk = clocks.stop(wall0, wall1) if (rank == 0 and full) else None
        last_loss = t.last_loss()
        value = world * B * args.steps / (ms / 1e3)

        # ---------------- per-kernel times for the roofline (live CUDA events, un-graphed steps) ----------------
        prof = {}
        n_prof = 5
        for i in range(n_prof + 1):
            rec = t.profile_step((i % nb) * B, B)
            if i == 0:
                continue  # first un-graphed step pays lazy module loading
            for kname, v in rec:
                prof[kname] = prof.get(kname, 0.0) + v / n_prof
        f_train, f_gemm, _ = flops_per_row(F, hidden)
        gemm_ms = sum(v for k, v in prof.items() if k.startswith("gemm_"))
        ach_tf = (B * f_gemm / (gemm_ms / 1e3)) / 1e12 if gemm_ms > 0 else 0.0
        dims = [F] + list(hidden)
        alg_bytes = sum(2 * (B * dims[i] + dims[i] * dims[i + 1] + B * dims[i + 1]) * 3 for i in range(len(hidden)))
        roofline = {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                    "traffic": NCU_TRAFFIC.get(name), "traffic_source": "profiles/ncu_r01_%s_gemm_full.txt" % name,
                    "algorithmic_operand_bytes": alg_bytes,
                    "kernel": "gemm_tc_kernel (all hidden-layer fwd/dA/dW GEMMs of one step; tcgen05 + TMA)",
                    "flops_per_launch_set": B * f_gemm, "kernel_ms_per_step": gemm_ms, "peak_source": peak_src,
                    "step_fraction_of_peak": (value / world * f_train / 1e12) / peak_tf,
                    "kernels_ms": {k: round(v, 5) for k, v in prof.items()}}
        res = {"value": value, "ms_per_step": ms / args.steps, "roofline": roofline, "last_loss": last_loss, "clocks": clk,
               "gradient_exchange": exchange, "gpu_launches": t.kernels_per_step(B) * args.steps,
               "step_api": "sb_trainer_step_resident_async per step" if per_step_calls else
                           "sb_trainer_run_resident (one call for all steps, four steps per captured graph)",
               "config": config_block(name, dict(c, n_batches=nb), world)}
        if not full:
            t.close()
            return res

        # ---------------- end-to-end leg: host (pinned) buffers through sb_trainer_step ----------------
        e2e_steps = args.e2e_steps or min(args.steps, 50)
        n_pin = 4
        pin = []
        for i in range(n_pin):
            px = torch.empty((B, F), dtype=torch.float32).pin_memory()
            py = torch.empty(B, dtype=torch.float32).pin_memory()
            pw = torch.empty(B, dtype=torch.float32).pin_memory()
            px.numpy()[:] = X[i * B:(i + 1) * B]; py.numpy()[:] = y[i * B:(i + 1) * B]; pw.numpy()[:] = w[i * B:(i + 1) * B]
            pin.append((px.numpy(), py.numpy(), pw.numpy()))
        for i in range(3):
            t.step(*pin[i % n_pin])
        barrier()
        e0 = time.perf_counter()
        for i in range(e2e_steps):
            t.step(*pin[i % n_pin])  # synchronous: H2D batch, load/cast kernel, step, D2H loss, host sees the loss
        barrier()
        sync_s = max_over_ranks(time.perf_counter() - e0)
        # pipelined public call: every step still copies its own batch H2D and its loss scalars D2H inside the timed
        # region, but the copy of batch i+1 overlaps the compute of batch i; the host reads the loss after the last step
        for i in range(3):
            t.step_async(*pin[i % n_pin])
        t.sync()
        barrier()
        e0 = time.perf_counter()
        for i in range(e2e_steps):
            t.step_async(*pin[i % n_pin])
        loss_h = t.last_loss()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - e0)
        res["e2e"] = {"value": world * B * e2e_steps / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": B * (F + 2) * 4,
                      "d2h_bytes_per_step": 16, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                      "synchronous_value": world * B * e2e_steps / sync_s, "last_loss": loss_h,
                      "timer": "host wall clock around sb_trainer_step_async x steps + sb_trainer_last_loss (pinned host buffers, "
                               "H2D of every batch and D2H of every step's loss scalars inside), max over ranks; "
                               "synchronous_value = the same with sb_trainer_step (host waits for each loss)"}

        # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
        res["cpu_baseline"] = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import shifu_oracle as so
            from oracle.torch_cpu_worker import time_train
            net = so.NetDesc(F, hidden, [so.ACT_RELU] * len(hidden))
            nbc = min(4, nb)
            batches = [(X[i * B:(i + 1) * B], y[i * B:(i + 1) * B], w[i * B:(i + 1) * B]) for i in range(nbc)]
            r = time_train(net, so.xavier_init(net, 1), so.OptConfig(kind=OPT_ID[c["optimizer"]], lr=c["lr"]), batches,
                           min_seconds=10.0, max_steps=400, threads=host_threads())
            res["cpu_baseline"] = {"value": r["rows_per_sec"], "unit": "rows/s", "cores": r["cores"], "kind": "port",
                                   "sample": "%d steps (%.1f s) of %s on torch-CPU fp32 = reference-equivalent worker loop "
                                             "(TF-1.x absent)" % (r["steps"], r["seconds"], name)}
        t.close()
        return res

    main_res = measure(args.config, True)
    # the data-parallel target of BASELINE.json is stated on cfg2 (2000 cols x 8192 rows/GPU): report it in the same line
    second = None
    if args.also and args.also != args.config:
        second = measure(args.also, False)

    if rank == 0:
        out = {
            "metric": "rows/sec tabular-DNN train", "value": main_res["value"], "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec == sb.PREC_BF16 else "f32", "data": "synthetic",
            "config": main_res["config"], "clocks": main_res["clocks"], "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"], "cpu_baseline": main_res["cpu_baseline"],
            "last_loss": main_res["last_loss"], "gradient_exchange": main_res["gradient_exchange"],
        }
        if second is not None:
            out["also"] = {k: second[k] for k in ("value", "ms_per_step", "config", "roofline", "last_loss", "gpu_launches")}
            out["also"]["unit"] = "rows/s"
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
