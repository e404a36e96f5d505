"""Host-side mirrors of the reference's two plug-in surfaces: the worker script (trainer.py <- ssgd_monitor.py) and the
scorer (scorer.py <- TensorflowModel.java).  CPU tests cover the contract (names, env vars, errors, loader semantics,
socket line, rendezvous); the gpu tests run the whole worker end to end."""
import gzip
import json
import os
import socket
import threading

import numpy as np
import pytest

from oracle import shifu_oracle as so


def _write_gz(path, X, y, w=None):
    with gzip.open(path, "wb") as f:
        for i in range(len(X)):
            cols = [str(int(y[i]))] + [repr(float(v)) for v in X[i]]
            if w is not None:
                cols.append(repr(float(w[i])))
            f.write(("|".join(cols) + "\n").encode())


class _Seq:
    def __init__(self, seed):
        self.r = np.random.RandomState(seed)

    def random(self):
        return float(self.r.rand())


def test_get_activation_fun_and_modelconf(sb):
    from shifu_tensorflow_b200 import trainer as tr
    assert [tr.get_activation_fun(n) for n in ("sigmoid", "TANH", "ReLU", "leakyrelu", "nope", None)] == [0, 1, 2, 3, 3, 3]
    conf = {"train": {"params": {"NumHiddenLayers": 2, "NumHiddenNodes": [10, 5, 99], "ActivationFunc": ["tanh", "relu", "x"],
                                  "LearningRate": 0.1}, "numTrainEpochs": 3, "validSetRate": 0.2}}
    assert tr.generate_from_modelconf(conf) == ([10, 5], [1, 2])
    d = tr.model(7, conf, 100)
    assert (d.n_features, d.n_hidden, d.hidden[0], d.hidden[1], d.optimizer, d.loss) == (7, 2, 10, 5, sb.OPT_ADADELTA, sb.LOSS_MSE)
    assert abs(d.learning_rate - 0.1) < 1e-7 and d.max_batch == 100


def test_load_data_equals_oracle_loader(sb, tmp_path):
    from shifu_tensorflow_b200 import trainer as tr
    X, y, w = so.synth_batch(57, 6, 3, weights="mixed")
    w = w.ravel(); w[5] = -2.0                      # negative weight -> 1.0
    p = str(tmp_path / "part-0.gz")
    _write_gz(p, X, y.ravel(), w)
    a = tr.load_data(p, [1, 2, 3, 4, 5, 6], 0, 7, 0.25, rng=_Seq(1))
    b = so.load_data([p], [1, 2, 3, 4, 5, 6], 0, 7, 0.25, rng=_Seq(1))
    for k in a:
        assert a[k] == b[k], k
    assert len(a["train_data"]) + len(a["valid_data"]) == 57 and [1.0] in a["train_data_sample_weight"] + a["valid_data_sample_weight"]


def test_missing_env_var_is_a_keyerror_like_the_reference(sb):
    from shifu_tensorflow_b200 import trainer as tr
    with pytest.raises(KeyError):
        tr.main(env={"CLUSTER_SPEC": "{}"})


def test_nccl_id_rendezvous_over_cluster_spec_address(sb, monkeypatch):
    """worker 0 serves the 128-byte id on its CLUSTER_SPEC address, the others fetch it (replaces tf.train.Server)"""
    from shifu_tensorflow_b200 import trainer as tr
    fake = bytes(range(128))
    monkeypatch.setattr(tr.capi, "nccl_unique_id", lambda: fake)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    spec = {"ps": ["127.0.0.1:1"], "worker": ["127.0.0.1:%d" % port, "127.0.0.1:2", "127.0.0.1:3"]}
    got = {}
    ths = [threading.Thread(target=lambda r=r: got.__setitem__(r, tr._exchange_nccl_id(spec, r, 3))) for r in range(3)]
    [t.start() for t in ths]; [t.join(30) for t in ths]
    assert got == {0: fake, 1: fake, 2: fake}
    assert tr._exchange_nccl_id(spec, 0, 1) is None


def test_scorer_init_errors_match_tensorflowmodel(sb):
    from shifu_tensorflow_b200.scorer import TensorflowModel, IllegalStateException, IllegalArgumentException
    m = TensorflowModel()
    with pytest.raises(IllegalStateException, match="TF model not initialized."):
        m.compute([0.0])
    with pytest.raises(RuntimeError, match="Config is null"):
        m.init(None)
    with pytest.raises(RuntimeError, match="Properties is null"):
        m.init({"inputnames": ["a"], "properties": {}})
    base = {"modelpath": "/x", "outputnames": "o", "tags": ["serve"]}
    for drop, msg in (("modelpath", "Model path is null"), ("outputnames", "Output names is null"), ("tags", "Tags is null")):
        props = dict(base); props.pop(drop)
        with pytest.raises(RuntimeError, match=msg):
            TensorflowModel().init({"inputnames": ["a"], "properties": props})
    with pytest.raises(RuntimeError, match="Input names is null"):
        TensorflowModel().init({"inputnames": [], "properties": base})
    with pytest.raises(IllegalArgumentException):
        TensorflowModel().init({"inputnames": ["a"], "properties": dict(base, outputnames=["o1", "o2"])})


def _run_worker(sb, tmp_path, n_rows, epochs, params_extra, seed=5, env_extra=None):
    from shifu_tensorflow_b200 import trainer as tr
    F = 12
    X, y, w = so.synth_batch(n_rows, F, 2, weights="ones")
    data = str(tmp_path / "part-00000.gz")
    _write_gz(data, X, y.ravel())
    conf = {"train": {"params": dict({"NumHiddenLayers": 2, "NumHiddenNodes": [8, 4], "ActivationFunc": ["tanh", "relu"],
                                      "LearningRate": 0.5, "Precision": "fp32"}, **params_extra),
                      "numTrainEpochs": epochs, "validSetRate": 0.2}}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    json.dump(conf, open("ModelConfig.json", "w"))
    srv = socket.socket(); srv.bind(("127.0.0.1", 0)); srv.listen(1)
    lines = []

    def serve():
        c, _ = srv.accept()
        buf = b""
        while True:
            d = c.recv(4096)
            if not d:
                break
            buf += d
        lines.extend(buf.decode().splitlines())

    th = threading.Thread(target=serve); th.start()
    env = {"CLUSTER_SPEC": json.dumps({"ps": ["127.0.0.1:1"], "worker": ["127.0.0.1:2"]}), "WORKER_CNT": "1", "JOB_NAME": "worker",
           "TASK_ID": "0", "SOCKET_SERVER_PORT": str(srv.getsockname()[1]), "TOTAL_TRAINING_DATA_NUMBER": str(n_rows),
           "SELECTED_COLUMN_NUMS": " ".join(str(i) for i in range(1, F + 1)), "WEIGHT_COLUMN_NUM": "-1", "TARGET_COLUMN_NUM": "0",
           "TMP_MODEL_PATH": str(tmp_path / "tmp_model"), "FINAL_MODEL_PATH": str(tmp_path / "final_model"),
           "TRAINING_DATA_PATH": data, "SB_SEED": "11"}
    env.update(env_extra or {})
    try:
        rc = tr.main(env=env, rng=_Seq(seed))
    finally:
        os.chdir(cwd)
    th.join(10); srv.close()
    return rc, lines, env, (X, y, w, F, conf)


@pytest.mark.gpu
def test_worker_end_to_end_reference_schedule(sb, tmp_path):
    """the whole plug-in: env contract -> load_data -> epoch-sync Adadelta training -> socket lines -> SavedModel"""
    rc, lines, env, (X, y, w, F, conf) = _run_worker(sb, tmp_path, 1000, 3, {})
    assert rc == 0 and len(lines) >= 1
    import re
    pat = re.compile(r"^worker_index:(\d+),time:([0-9.e+-]+),current_epoch:(\d+),training_loss:([0-9.e+-]+),valid_loss:([0-9.e+-]+)$")
    parsed = [pat.match(l) for l in lines]
    assert all(parsed), lines                       # exactly what SocketServer.java:71-89 splits on ',' and ':'
    assert int(parsed[-1].group(3)) == 3            # StopAtStepHook(num_steps=EPOCH)
    final = env["FINAL_MODEL_PATH"]
    assert sorted(os.listdir(final)) == ["GenericModelConfig.json", "saved_model.pb", "variables"]
    assert os.path.exists(os.path.join(env["TMP_MODEL_PATH"], "model.ckpt"))
    # the exported model scores through the scorer mirror exactly like the oracle does on the exported weights
    from shifu_tensorflow_b200.scorer import TensorflowModel
    cfg = json.load(open(os.path.join(final, "GenericModelConfig.json")))
    cfg["properties"]["modelpath"] = final
    m = TensorflowModel(); m.init(cfg); m.init(cfg)          # second init is a no-op
    Fn, hidden, acts, out_act, flat = sb.capi.savedmodel_read(final, "shifu_input_0", "shifu_output_0")
    net = so.NetDesc(Fn, hidden, acts)
    want = so.score_rows(net, so.unflatten_params(net, flat), X[:50].astype(np.float64))
    got = m.computeBatch(X[:50].astype(np.float64))
    assert np.abs(got - want).max() <= 1e-5
    assert abs(m.compute(X[3].astype(np.float64)) - want[3]) <= 1e-5
    m.releaseResource()


@pytest.mark.gpu
def test_worker_matches_oracle_epoch_sync_trajectory(sb, tmp_path):
    """same data, same split (seeded), same init -> the loss the worker reports per epoch equals the oracle's
    epoch-sync trajectory (mean of the R mini-batch gradients, one Adadelta update per epoch) within 1e-4"""
    from shifu_tensorflow_b200 import trainer as tr
    rc, lines, env, (X, y, w, F, conf) = _run_worker(sb, tmp_path, 1000, 3, {})
    ctx = tr.load_data(env["TRAINING_DATA_PATH"], list(range(1, F + 1)), 0, -1, 0.2, rng=_Seq(5))
    tx = np.asarray(ctx["train_data"], np.float32); ty = np.asarray(ctx["train_target"], np.float32)
    tw = np.asarray(ctx["train_data_sample_weight"], np.float32)
    vx = np.asarray(ctx["valid_data"], np.float32); vy = np.asarray(ctx["valid_target"], np.float32)
    vw = np.asarray(ctx["valid_data_sample_weight"], np.float32)
    net = so.NetDesc(F, [8, 4], [so.ACT_TANH, so.ACT_RELU])
    # the worker initialised with sb_trainer_init_xavier(seed 11): read the same start back from a fresh trainer
    with sb.Trainer(tr.model(F, conf, 128)) as t0:
        t0.init_xavier(11)
        theta = t0.get_params()
    opt = so.Optimizer(so.OptConfig(kind=so.OPT_ADADELTA, lr=0.5), theta.size)
    batches = so.split_batches(len(tx), 100)
    R = so.replicas_to_aggregate(1000, 0.2, 100)
    valid_losses, pend, gsum = [], 0, np.zeros_like(theta)
    steps = 0
    while steps < 3:
        for idx in batches:
            P = so.unflatten_params(net, theta)
            L, g, _ = so.loss_and_grads(net, P, tx[idx], ty[idx], tw[idx])
            gsum += so.flatten_params(g); pend += 1
            if pend >= R:
                theta = opt.apply(theta, gsum / np.float32(pend)); gsum[:] = 0; pend = 0; steps += 1
                if steps >= 3:
                    break
        A, z, yh = so.forward(net, so.unflatten_params(net, theta), vx)
        valid_losses.append(float(so.loss_value(z, yh, vy, vw, so.LOSS_MSE)[0]))
    got = [float(l.split("valid_loss:")[1]) for l in lines]
    assert len(got) == len(valid_losses)
    assert np.abs(np.array(got) - np.array(valid_losses)).max() <= 1e-4


def test_equal_size_runs_cover_array_split_batches(sb):
    """the per-batch schedule hands runs of equally sized mini-batches to sb_trainer_run_resident: np.array_split
    (ssgd_monitor.py:189-192) yields sizes that differ by at most one row, i.e. at most two runs"""
    from shifu_tensorflow_b200 import trainer as tr
    for n_rows, total_batch in [(1000, 10), (1003, 10), (7, 7), (95, 4), (5, 1)]:
        bounds = [b[0] for b in np.array_split(np.arange(n_rows), total_batch)] + [n_rows]
        runs = tr.equal_size_runs(bounds)
        assert len(runs) <= 2
        covered = []
        for first, count, rows in runs:
            for k in range(count):
                assert bounds[first + k + 1] - bounds[first + k] == rows
                covered.append(first + k)
        assert covered == list(range(total_batch))


class _FakeTrainer:
    """stands in for capi.Trainer so that the worker's control flow (schedules, stop condition, metrics lines,
    checkpoint / export calls) runs without a GPU; every call is recorded"""
    instances = []

    def __init__(self, desc, device=0, nccl_id=None, rank=0, world=1):
        self.desc, self.rank, self.world = desc, rank, world
        self.calls, self._gs, self._acc, self.n_rows = [], 0, 0, 0
        _FakeTrainer.instances.append(self)

    global_step = property(lambda self: self._gs)

    def init_xavier(self, seed): self.calls.append(("init_xavier", seed))
    def load_checkpoint(self, path): self.calls.append(("load_checkpoint", path))
    def load_dataset(self, X, y, w=None):
        assert X.dtype == np.float32 and X.ndim == 2 and len(X) == len(y) == len(w)
        self.n_rows = len(X); self.calls.append(("load_dataset", len(X)))
    def _check(self, off, rows):
        assert 0 <= off and rows > 0 and off + rows <= self.n_rows and rows <= self.desc.max_batch
    def step_resident(self, off, rows):
        self._check(off, rows); self._gs += 1; self.calls.append(("step", off, rows)); return 0.5
    def run_resident(self, offs, rows):
        for o in offs: self._check(o, rows)
        self._gs += len(offs); self.calls.append(("run", list(offs), rows))
    def last_loss(self): return 0.25
    def accumulate_resident(self, off, rows):
        self._check(off, rows); self._acc += 1; self.calls.append(("acc", off, rows)); return 0.3
    def apply_accumulated(self):
        assert self._acc > 0
        self.calls.append(("apply", self._acc)); self._acc = 0; self._gs += 1
    def eval_loss(self, X, y, w=None): self.calls.append(("eval", len(X))); return 0.125
    def save_checkpoint(self, path): open(path, "w").write("ckpt"); self.calls.append(("save", path))
    def export_savedmodel(self, d): os.makedirs(d); self.calls.append(("export", d))
    def close(self): self.calls.append(("close",))


@pytest.fixture
def fake_trainer(sb, monkeypatch):
    from shifu_tensorflow_b200 import trainer as tr
    _FakeTrainer.instances = []
    monkeypatch.setattr(tr.capi, "Trainer", _FakeTrainer)
    return _FakeTrainer


def test_worker_control_flow_per_batch_schedule(sb, tmp_path, fake_trainer):
    """Schedule "batch": numTrainEpochs counts update steps (StopAtStepHook on global_step, ssgd_monitor.py:235); the
    epoch's batch loop goes through run_resident in runs of equally sized batches and stops exactly at the step limit"""
    rc, lines, env, _ = _run_worker(sb, tmp_path, 1000, 11, {"Schedule": "batch", "MiniBatchs": 100},
                                    env_extra={"SB_HOST_LOADER": "1"})
    assert rc == 0
    t = fake_trainer.instances[0]
    n_train = t.n_rows
    total_batch = n_train // 100
    runs = [c for c in t.calls if c[0] == "run"]
    steps = [(o, c[2]) for c in runs for o in c[1]]
    assert len(steps) == 11 == t.global_step                       # stopped at the limit, inside the second epoch
    bounds = [b[0] for b in np.array_split(np.arange(n_train), total_batch)] + [n_train]
    want = [(int(bounds[i]), int(bounds[i + 1] - bounds[i])) for i in range(total_batch)]
    assert steps == (want + want)[:11]                             # batch order of np.array_split, epoch after epoch
    assert not [c for c in t.calls if c[0] in ("acc", "apply", "step")]
    assert len(lines) == 2 and lines[0].startswith("worker_index:0,time:") and ",current_epoch:%d," % total_batch in lines[0]
    assert ",current_epoch:11,training_loss:0.25,valid_loss:0.125" in lines[1]
    assert [c[0] for c in t.calls].count("save") == 2 and t.calls[-2][0] == "export" and t.calls[-1] == ("close",)
    assert os.path.isdir(env["FINAL_MODEL_PATH"])


def test_worker_control_flow_epoch_schedule_and_restart(sb, tmp_path, fake_trainer):
    """reference schedule: R = N_train / batch mini-batch gradients per update, one update (= one global step) per
    epoch; a checkpoint left in TMP_MODEL_PATH is restored instead of a fresh init (ssgd_monitor.py:251-257)"""
    rc, lines, env, _ = _run_worker(sb, tmp_path, 1000, 3, {}, env_extra={"SB_HOST_LOADER": "1"})
    assert rc == 0
    t = fake_trainer.instances[0]
    applies = [c for c in t.calls if c[0] == "apply"]
    accs = [c for c in t.calls if c[0] == "acc"]
    assert len(applies) == 3 == t.global_step and len(lines) == 3
    assert len(accs) == sum(a[1] for a in applies)
    assert ("init_xavier", 11) in t.calls and not [c for c in t.calls if c[0] == "load_checkpoint"]
    assert [l.split(",")[2] for l in lines] == ["current_epoch:1", "current_epoch:2", "current_epoch:3"]
    # second run over the same TMP_MODEL_PATH: restores
    import shutil
    shutil.rmtree(env["FINAL_MODEL_PATH"])
    rc, _, _, _ = _run_worker(sb, tmp_path, 1000, 3, {}, env_extra={"SB_HOST_LOADER": "1"})
    t2 = fake_trainer.instances[1]
    assert rc == 0 and [c for c in t2.calls if c[0] == "load_checkpoint"] and not [c for c in t2.calls if c[0] == "init_xavier"]


def test_worker_row_shard_from_the_launcher(sb, tmp_path, fake_trainer):
    rc, _, _, _ = _run_worker(sb, tmp_path, 1000, 1, {}, env_extra={"SB_HOST_LOADER": "1"})
    whole = fake_trainer.instances[0].n_rows
    rc2, _, _, _ = _run_worker(sb, tmp_path, 1000, 1, {}, env_extra={"SB_HOST_LOADER": "1", "SB_ROW_SHARD": "1/4"})
    assert rc == 0 and rc2 == 0 and fake_trainer.instances[1].n_rows == whole // 4


def test_allgather_bytes_and_peer_exchange_decision(sb):
    """TCP hub all-gather on worker 0's CLUSTER_SPEC address (rank order, ragged payloads), and the rule that the
    peer-memory exchange is only switched on when every rank reports the same host"""
    from shifu_tensorflow_b200 import trainer as tr
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    spec = {"ps": [], "worker": ["127.0.0.1:%d" % port] + ["10.0.0.%d:1" % i for i in range(1, 4)]}
    n = 4
    out, errs = [None] * n, []

    def rank(r):
        try:
            out[r] = tr.allgather_bytes(spec, r, n, b"x" * r + bytes([r]))
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=rank, args=(r,)) for r in (3, 1, 0, 2)]
    [t.start() for t in th]; [t.join(20) for t in th]
    assert not errs, errs
    want = [b"x" * r + bytes([r]) for r in range(n)]
    assert all(o == want for o in out)

    class T:                                     # records what the trainer is asked to do
        def __init__(self, r): self.r, self.peers = r, None
        def ipc_handle(self): return bytes([self.r]) * 64
        def set_peer_handles(self, hs): self.peers = list(hs)

    for same_host, expect in ((True, True), (False, False)):
        ts, res = [T(r) for r in range(n)], [None] * n
        names = iter(["nodeA"] * n if same_host else ["nodeA", "nodeA", "nodeB", "nodeA"])
        lock = threading.Lock()
        real = socket.gethostname

        def fake_hostname():
            with lock:
                return next(names)

        def run(r):
            res[r] = tr.enable_peer_exchange(ts[r], spec, r, n)

        tr.socket.gethostname = fake_hostname
        try:
            th = [threading.Thread(target=run, args=(r,)) for r in range(n)]
            [t.start() for t in th]; [t.join(30) for t in th]
        finally:
            tr.socket.gethostname = real
        assert res == [expect] * n
        for t in ts:
            assert (t.peers == [bytes([q]) * 64 for q in range(n)]) if expect else (t.peers is None)
