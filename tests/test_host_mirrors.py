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


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _in_threads(n, fn, order=None):
    out, errs = [None] * n, []

    def run(r):
        try:
            out[r] = fn(r)
        except Exception as e:      # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=run, args=(r,)) for r in (order or range(n))]
    [t.start() for t in th]; [t.join(60) for t in th]
    assert not errs, errs
    return out


def test_nccl_id_rendezvous_over_cluster_spec_address(sb, monkeypatch):
    """worker 0 serves the 128-byte id on its CLUSTER_SPEC address, the others fetch it (replaces tf.train.Server)"""
    from shifu_tensorflow_b200 import trainer as tr
    fake = bytes(range(128))
    monkeypatch.setattr(tr.capi, "nccl_unique_id", lambda: fake)
    spec = {"ps": ["127.0.0.1:1"], "worker": ["127.0.0.1:%d" % _free_port(), "127.0.0.1:2", "127.0.0.1:3"]}

    def rank(r):
        rdv = tr.Rendezvous(spec, r, 3, timeout=30)
        try:
            return tr.exchange_nccl_id(rdv)
        finally:
            rdv.close()

    assert _in_threads(3, rank) == [fake, fake, fake]
    assert tr.exchange_nccl_id(tr.Rendezvous(spec, 0, 1)) is None


def test_rendezvous_times_out_instead_of_hanging(sb):
    """a worker that never shows up is an error on worker 0 within the deadline, not an accept() that blocks forever"""
    from shifu_tensorflow_b200 import trainer as tr
    spec = {"worker": ["127.0.0.1:%d" % _free_port(), "x:1"]}
    with pytest.raises(RuntimeError, match="did not connect"):
        tr.Rendezvous(spec, 0, 2, timeout=0.5)
    with pytest.raises(RuntimeError, match="unreachable"):
        tr.Rendezvous({"worker": ["127.0.0.1:%d" % _free_port(), "x:1"]}, 1, 2, timeout=0.5)


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
    # extra named inputs (TensorflowModel.java:73-83; TensorflowModelTest.java:44-47 feeds a Keras learning-phase bool):
    # accepted when they select the inference branch, rejected when they ask for training or are not a phase switch
    from shifu_tensorflow_b200.scorer import check_phase_switch
    for ok in (False, 0, 0.0, None, np.bool_(False)):
        check_phase_switch("dropout_1/keras_learning_phase", ok)
    for bad in (True, 1, 0.5, "false", [0]):
        with pytest.raises(IllegalArgumentException):
            check_phase_switch("dropout_1/keras_learning_phase", bad)
    with pytest.raises(IllegalArgumentException, match="training branch"):
        TensorflowModel().init({"inputnames": ["a", "phase"], "properties": dict(base, phase=True)})


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
def test_worker_matches_oracle_sync_replicas_trajectory(sb, tmp_path):
    """same data, same split (seeded), same init -> the losses the worker reports per epoch (training loss of the last
    mini-batch, full validation loss, global step) equal oracle.SyncReplicasTrainer driven by the reference's loop
    (ssgd_monitor.py:268-285: for i in range(total_batch): sess.run(train_step); then the validation run) within 1e-4"""
    from shifu_tensorflow_b200 import trainer as tr
    rc, lines, env, (X, y, w, F, conf) = _run_worker(sb, tmp_path, 1000, 4, {"Precision": "fp32"})
    assert rc == 0
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
    R = so.replicas_to_aggregate(1000, 0.2, 100)
    ref = so.SyncReplicasTrainer(net, so.unflatten_params(net, theta), so.OptConfig(kind=so.OPT_ADADELTA, lr=0.5), R)
    batches = so.split_batches(len(tx), 100)
    want = []
    while ref.global_step < 4:
        for idx in batches:
            L, gs = ref.run(tx[idx], ty[idx], tw[idx])
            if gs >= 4:
                break
        A, z, yh = so.forward(net, so.unflatten_params(net, ref.theta), vx)
        want.append((gs, float(L), float(so.loss_value(z, yh, vy, vw, so.LOSS_MSE)[0])))
    got = [(int(l.split("current_epoch:")[1].split(",")[0]), float(l.split("training_loss:")[1].split(",")[0]),
            float(l.split("valid_loss:")[1])) for l in lines]
    assert [g[0] for g in got] == [x[0] for x in want]
    assert np.abs(np.array(got)[:, 1:] - np.array(want)[:, 1:]).max() <= 1e-4


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
    def loss_resident(self, off, rows):
        self._check(off, rows); self.calls.append(("loss", off, rows)); return 0.3
    def broadcast_state(self, root=0): self.calls.append(("bcast", root))
    def apply_accumulated(self, total_pushes=None):
        assert self._acc > 0 and (total_pushes is None or total_pushes == self._acc)
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
    R = so.replicas_to_aggregate(1000, 0.2, 100)
    assert len(applies) == 3 == t.global_step
    assert [a[1] for a in applies] == [R] * 3 and len(accs) == 3 * R       # every update is the mean of R accepted pushes
    # the first push after the first update still carries local_step 0 and is dropped (it only reports its loss); the
    # update points therefore drift against the for-loop's epoch boundaries exactly like in the reference
    assert len([c for c in t.calls if c[0] == "loss"]) == 1
    assert ("init_xavier", 11) in t.calls and not [c for c in t.calls if c[0] == "load_checkpoint"]
    epochs_seen = [int(l.split(",")[2].split(":")[1]) for l in lines]
    assert epochs_seen == sorted(epochs_seen) and epochs_seen[-1] == 3 and len(lines) >= 3
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


def test_rendezvous_allgather_rounds_and_peer_exchange_decision(sb):
    """one persistent hub on worker 0's CLUSTER_SPEC address: several all-gather rounds over the same connections (rank
    order, ragged payloads); the peer-memory exchange is switched on only when every rank reports the same host AND every
    rank mapped its peers - one failing rank sends everybody back to NCCL"""
    from shifu_tensorflow_b200 import trainer as tr
    n = 4
    spec = {"ps": [], "worker": ["127.0.0.1:%d" % _free_port()] + ["10.0.0.%d:1" % i for i in range(1, 4)]}

    def rounds(r):
        rdv = tr.Rendezvous(spec, r, n, timeout=30)
        try:
            return [rdv.allgather(b"x" * r + bytes([r, k])) for k in range(5)], rdv.bcast(b"root" if r == 0 else None)
        finally:
            rdv.close()

    for got, root in _in_threads(n, rounds, order=(3, 1, 0, 2)):
        assert got == [[b"x" * r + bytes([r, k]) for r in range(n)] for k in range(5)] and root == b"root"

    class T:                                     # records what the trainer is asked to do
        def __init__(self, r, fail): self.r, self.peers, self.fail, self.cleared = r, None, fail, False
        def ipc_handle(self): return bytes([self.r]) * 64
        def set_peer_handles(self, hs):
            if self.fail:
                raise RuntimeError("cudaIpcOpenMemHandle failed")
            self.peers = list(hs)
        def clear_peer_handles(self): self.cleared, self.peers = True, None

    for hosts, failing, expect in ((["nodeA"] * n, None, True), (["nodeA", "nodeA", "nodeB", "nodeA"], None, False),
                                   (["nodeA"] * n, 2, False)):
        spec = {"ps": [], "worker": ["127.0.0.1:%d" % _free_port()] + ["10.0.0.%d:1" % i for i in range(1, 4)]}
        ts = [T(r, r == failing) for r in range(n)]
        names = iter(hosts)
        lock = threading.Lock()
        real = socket.gethostname

        def fake_hostname():
            with lock:
                return next(names)

        def run(r):
            rdv = tr.Rendezvous(spec, r, n, timeout=30)
            try:
                return tr.enable_peer_exchange(ts[r], rdv)
            finally:
                rdv.close()

        tr.socket.gethostname = fake_hostname
        try:
            res = _in_threads(n, run)
        finally:
            tr.socket.gethostname = real
        assert res == [expect] * n
        for t in ts:
            if expect:
                assert t.peers == [bytes([q]) * 64 for q in range(n)]
            else:
                assert t.peers is None and (failing is None or t.cleared or t.fail)


def test_sync_replicas_schedule_is_the_oracle_state_machine(sb):
    """trainer.SyncReplicasSchedule (host bookkeeping the worker runs) against oracle.SyncReplicasTrainer (the restatement
    of SyncReplicasOptimizer, ssgd_monitor.py:136-142,218,259-260): same accepted / dropped pushes, same update points"""
    from shifu_tensorflow_b200 import trainer as tr
    net = so.NetDesc(6, [4], [so.ACT_TANH])
    params = so.xavier_init(net, 3)
    X, y, w = so.synth_batch(40, 6, 1)
    for R in (1, 2, 3, 7):
        ref = so.SyncReplicasTrainer(net, params, so.OptConfig(kind=so.OPT_SGD, lr=0.1), R)
        sched = tr.SyncReplicasSchedule(R, 1)
        for k in range(6 * R + 5):
            acc_before, gs_before = ref.acc_n, ref.global_step
            fresh = ref.local_step >= ref.global_step
            ref.run(X, y, w)
            ran, accepted, apply_now, pushes = sched.round()
            assert ran == [True] and accepted == [fresh], (R, k)
            assert apply_now == (ref.global_step == gs_before + 1), (R, k)
            if apply_now:
                assert pushes == R
            assert sched.global_step == ref.global_step and sched.local_step[0] == ref.local_step
    # several workers in lock-step, rank order = arrival order: R accepted pushes per update, at most one update per
    # round, a push behind the update in the same round is stale
    for R, n in ((4, 2), (5, 2), (3, 4), (1, 4), (8, 8)):
        sched = tr.SyncReplicasSchedule(R, n)
        acc, updates = 0, 0
        for k in range(60):
            ran, accepted, apply_now, pushes = sched.round()
            assert any(ran) and all(ran[r] or not accepted[r] for r in range(n))
            acc += sum(accepted)
            if apply_now:
                assert pushes == R == acc
                acc, updates = 0, updates + 1
            assert acc < R
        assert updates >= 60 * min(n, R) // (R + n) and sched.global_step == updates


_FAKE_HDFS = """#!/bin/bash
# stand-in for the `hdfs dfs` command line of a YARN container: hdfs://nn/<p> lives under $FAKE_HDFS_ROOT/<p>
shift
map() { echo "${1/hdfs:\\/\\/nn/$FAKE_HDFS_ROOT}"; }
case "$1" in
  -cat) cat "$(map "$2")" ;;
  -test) [ -e "$(map "$3")" ] ;;
  -get) cp -r "$(map "$3")" "$4" ;;
  -rm) rm -rf "$(map "$4")" ;;
  -mkdir) mkdir -p "$(map "$3")" ;;
  -put) rm -rf "$(map "$4")"; cp -r "$3" "$(map "$4")" ;;
  *) echo "unsupported $1" >&2; exit 2 ;;
esac
"""


def test_worker_reads_and_writes_scheme_paths_through_the_hdfs_cli(sb, tmp_path, fake_trainer, monkeypatch):
    """the stock AM hands out hdfs:// URIs (TrainingDataSet.java:74) and HDFS model paths, which the reference reads with
    tf.gfile: training files are fetched with `hdfs dfs -cat`, checkpoint and SavedModel are staged locally and uploaded;
    without the CLI the worker fails fast with a message that says so"""
    from shifu_tensorflow_b200 import trainer as tr
    root = tmp_path / "hdfs_root"
    (root / "data").mkdir(parents=True)
    cli = tmp_path / "hdfs"
    cli.write_text(_FAKE_HDFS); cli.chmod(0o755)
    monkeypatch.setenv("FAKE_HDFS_ROOT", str(root))
    monkeypatch.setattr(tr._Fs, "CLI", str(cli))
    work = tmp_path / "w"; work.mkdir()
    rc, lines, env, (X, y, w, F, conf) = _run_worker(sb, work, 400, 2, {}, env_extra={"SB_HOST_LOADER": "1"})   # writes the gz file
    os.replace(env["TRAINING_DATA_PATH"], root / "data" / "part-00000.gz")
    fake_trainer.instances.clear()
    work2 = tmp_path / "w2"; work2.mkdir()
    extra = {"SB_HOST_LOADER": "1", "TRAINING_DATA_PATH": "hdfs://nn/data/part-00000.gz", "TMP_MODEL_PATH": "hdfs://nn/tmp_model",
             "FINAL_MODEL_PATH": "hdfs://nn/final_model"}

    real_run = _run_worker.__globals__["_write_gz"]
    _run_worker.__globals__["_write_gz"] = lambda *a, **k: None      # the data already lives on "HDFS"
    try:
        rc, lines, env, _ = _run_worker(sb, work2, 400, 2, {}, env_extra=extra)
    finally:
        _run_worker.__globals__["_write_gz"] = real_run
    assert rc == 0 and len(lines) >= 2
    t = fake_trainer.instances[0]
    assert t.n_rows > 0 and ("init_xavier", 11) in t.calls
    assert (root / "tmp_model" / "model.ckpt").read_text() == "ckpt"           # staged locally, uploaded
    assert (root / "final_model").is_dir()
    assert not os.path.exists(work2 / "hdfs:")                                  # nothing written into a local 'hdfs:' directory
    # second run: the checkpoint on "HDFS" is fetched and restored by worker 0
    fake_trainer.instances.clear()
    _run_worker.__globals__["_write_gz"] = lambda *a, **k: None
    try:
        rc, _, _, _ = _run_worker(sb, tmp_path / "w2", 400, 2, {}, env_extra=extra)
    finally:
        _run_worker.__globals__["_write_gz"] = real_run
    assert rc == 0 and [c for c in fake_trainer.instances[0].calls if c[0] == "load_checkpoint"]
    # no CLI -> clear failure
    monkeypatch.setattr(tr._Fs, "CLI", str(tmp_path / "no_such_hdfs"))
    with pytest.raises(RuntimeError, match="needs the"):
        tr._Fs.read_bytes("hdfs://nn/data/part-00000.gz")


@pytest.mark.gpu
def test_worker_wide_deep_from_the_numeric_and_category_column_lists(sb, tmp_path):
    """SELECTED_COLUMN_NUMS blank + SELECTED_NUMERIC_/CATEGORY_COLUMN_NUMS (TensorflowTaskExecutor.java:213-223): the worker
    trains the wide+deep model through sparse steps and exports an ordinary SavedModel whose first layer has
    n_dense + sum(vocab) inputs; its losses follow the dense oracle on the materialised one-hot matrix"""
    from oracle import wide_deep as wd
    from shifu_tensorflow_b200 import trainer as tr
    n_dense, vocab, n_rows = 6, [4, 7, 3], 900
    Xd, idx, y, w = wd.synth_wide_deep_batch(n_rows, n_dense, vocab, 3)
    offs = np.concatenate([[0], np.cumsum(vocab)[:-1]])
    codes = np.where(idx >= 0, idx - offs[None, :], -1)
    data = str(tmp_path / "part-00000.gz")
    with gzip.open(data, "wb") as f:
        for i in range(n_rows):
            cols = [str(int(y[i, 0]))] + [repr(float(v)) for v in Xd[i]] + [str(int(c)) for c in codes[i]]
            f.write(("|".join(cols) + "\n").encode())
    conf = {"train": {"params": {"NumHiddenLayers": 2, "NumHiddenNodes": [16, 8], "ActivationFunc": ["relu", "tanh"], "LearningRate": 0.1,
                                 "Optimizer": "sgd", "Precision": "fp32", "MiniBatchs": 100}, "numTrainEpochs": 14, "validSetRate": 0.2}}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    json.dump(conf, open("ModelConfig.json", "w"))
    env = {"CLUSTER_SPEC": json.dumps({"ps": ["127.0.0.1:1"], "worker": ["127.0.0.1:2"]}), "WORKER_CNT": "1", "JOB_NAME": "worker",
           "TASK_ID": "0", "SOCKET_SERVER_PORT": "1", "SB_REQUIRE_SOCKET": "0", "TOTAL_TRAINING_DATA_NUMBER": str(n_rows),
           "SELECTED_COLUMN_NUMS": "", "SELECTED_NUMERIC_COLUMN_NUMS": " ".join(str(1 + i) for i in range(n_dense)),
           "SELECTED_CATEGORY_COLUMN_NUMS": " ".join(str(1 + n_dense + i) for i in range(len(vocab))),
           "SB_CATEGORY_VOCAB": " ".join(str(v) for v in vocab), "WEIGHT_COLUMN_NUM": "-1", "TARGET_COLUMN_NUM": "0",
           "TMP_MODEL_PATH": str(tmp_path / "tmp_model"), "FINAL_MODEL_PATH": str(tmp_path / "final_model"),
           "TRAINING_DATA_PATH": data, "SB_SEED": "11"}
    try:
        rc = tr.main(env=env, rng=_Seq(5))
    finally:
        os.chdir(cwd)
    assert rc == 0
    Fn, hidden, acts, out_act, flat = sb.capi.savedmodel_read(env["FINAL_MODEL_PATH"], "shifu_input_0", "shifu_output_0")
    assert Fn == n_dense + sum(vocab) and hidden == [16, 8]
    # replay with the dense oracle on the one-hot matrix: same split, same init, one SGD update per mini-batch
    coins = np.array([_c >= 0.2 for _c in (lambda r: [r.random() for _ in range(n_rows)])(_Seq(5))])
    Xfull = np.concatenate([Xd, wd.onehot_matrix(idx, sum(vocab))], axis=1).astype(np.float32)
    tx, ty = Xfull[coins], y[coins]
    net = so.NetDesc(Fn, [16, 8], [so.ACT_RELU, so.ACT_TANH])
    with sb.Trainer(sb.make_desc(Fn, [16, 8], [so.ACT_RELU, so.ACT_TANH], max_batch=128)) as t0:
        t0.init_xavier(11)
        theta = t0.get_params()
    ref = so.CleanTrainer(net, so.unflatten_params(net, theta), so.OptConfig(kind=so.OPT_SGD, lr=0.1))
    steps = 0
    while steps < 14:
        for b in so.split_batches(len(tx), 100):
            ref.step([(tx[b], ty[b], np.ones((len(b), 1), np.float32))]); steps += 1
            if steps >= 14:
                break
    assert np.abs(flat - ref.theta).max() <= 1e-4
