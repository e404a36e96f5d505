"""Node launcher (shifu-tensorflow_b200/launcher.py): env rewriting for G local ranks, metrics relay, exit codes.
CPU only: the worker command is replaced by a tiny script that speaks the worker's side of the contract."""
import json
import os
import socket
import sys
import textwrap
import threading

import numpy as np
import pytest


def _env(tmp_path, port, task=1, workers=2):
    return {
        "JOB_NAME": "worker", "TASK_ID": str(task), "WORKER_CNT": str(workers),
        "CLUSTER_SPEC": json.dumps({"ps": ["10.0.0.1:2000"], "worker": ["10.0.0.2:3000", "10.0.0.3:3001"][:workers]}),
        "SOCKET_SERVER_PORT": str(port), "TRAINING_DATA_PATH": "a.gz,b.gz", "TOTAL_TRAINING_DATA_NUMBER": "1000",
        "SELECTED_COLUMN_NUMS": "1 2 3", "WEIGHT_COLUMN_NUM": "-1", "TARGET_COLUMN_NUM": "0",
        "TMP_MODEL_PATH": str(tmp_path / "tmp"), "FINAL_MODEL_PATH": str(tmp_path / "final"),
        "SB_LOCAL_GPUS": "4", "PATH": os.environ.get("PATH", ""), "OUT_DIR": str(tmp_path),
    }


FAKE_WORKER = textwrap.dedent('''
    import json, os, socket, sys
    g = int(os.environ["LOCAL_RANK"])
    json.dump({k: os.environ[k] for k in ("WORKER_CNT", "TASK_ID", "CLUSTER_SPEC", "LOCAL_RANK", "SB_DEVICE", "SB_ROW_SHARD",
                                         "SOCKET_SERVER_PORT", "TRAINING_DATA_PATH")},
              open(os.path.join(os.environ["OUT_DIR"], "env_%d.json" % g), "w"))
    s = socket.create_connection(("127.0.0.1", int(os.environ["SOCKET_SERVER_PORT"])))
    for epoch in (1, 2):
        s.sendall(("worker_index:%s,time:%s,current_epoch:%d,training_loss:%s,valid_loss:%s\\n"
                   % (os.environ["TASK_ID"], 1.0 + g, epoch, 0.1 * (g + 1) * epoch, 0.2 * (g + 1))).encode())
    s.close()
    sys.exit(int(os.environ.get("FAIL_RANK", "-1")) == g and 7 or 0)
''')


class _Upstream(threading.Thread):
    """stands in for the executor's SocketServer (SocketServer.java:56-112): one client, newline-separated lines"""

    def __init__(self):
        super().__init__(daemon=True)
        self.srv = socket.socket(); self.srv.bind(("127.0.0.1", 0)); self.srv.listen(1)
        self.port = self.srv.getsockname()[1]
        self.data = b""

    def run(self):
        self.srv.settimeout(20)
        try:
            conn, _ = self.srv.accept()
        except OSError:
            return
        with conn:
            while True:
                chunk = conn.recv(4096)
                if not chunk:
                    break
                self.data += chunk


def test_aggregate_lines_and_parse(sb):
    from shifu_tensorflow_b200 import launcher as la
    lines = ["worker_index:%d,time:%s,current_epoch:3,training_loss:%s,valid_loss:%s\n" % (4 + g, 1.0 + g, 0.1 * (g + 1), 0.5)
             for g in range(4)]
    out = la.parse_metrics_line(la.aggregate_lines(1, lines))
    assert out["worker_index"] == 1 and out["current_epoch"] == 3 and out["time"] == 4.0
    assert abs(out["training_loss"] - 0.25) < 1e-12 and out["valid_loss"] == 0.5
    with pytest.raises(ValueError):
        la.aggregate_lines(0, [lines[0], lines[1].replace("current_epoch:3", "current_epoch:4")])
    with pytest.raises(ValueError):
        la.parse_metrics_line("worker_index:0,time:1")


def test_launcher_fans_out_relays_and_returns_zero(sb, tmp_path):
    from shifu_tensorflow_b200 import launcher as la
    up = _Upstream(); up.start()
    script = tmp_path / "fake_worker.py"
    script.write_text(FAKE_WORKER)
    rc = la.main(env=_env(tmp_path, up.port), worker_cmd=[sys.executable, str(script)])
    up.join(timeout=10)
    assert rc == 0
    envs = [json.load(open(tmp_path / ("env_%d.json" % g))) for g in range(4)]
    assert [e["TASK_ID"] for e in envs] == ["4", "5", "6", "7"]          # container 1 of 2, 4 GPUs each
    assert all(e["WORKER_CNT"] == "8" for e in envs)
    assert [e["SB_ROW_SHARD"] for e in envs] == ["0/4", "1/4", "2/4", "3/4"]
    assert [e["SB_DEVICE"] for e in envs] == ["0", "1", "2", "3"]
    spec = json.loads(envs[0]["CLUSTER_SPEC"])
    assert len(spec["worker"]) == 8 and spec["worker"][0] == "10.0.0.2:3000" and spec["ps"] == ["10.0.0.1:2000"]
    assert all(e["SOCKET_SERVER_PORT"] != str(up.port) for e in envs)      # the ranks talk to the relay
    lines = [l for l in up.data.decode().split("\n") if l]
    assert len(lines) == 2                                                 # ONE line per epoch for the container
    recs = sorted((la.parse_metrics_line(l) for l in lines), key=lambda r: r["current_epoch"])
    for ep, r in zip((1, 2), recs):
        assert r["worker_index"] == 1 and r["current_epoch"] == ep and r["time"] == 4.0
        assert abs(r["training_loss"] - np.mean([0.1 * (g + 1) * ep for g in range(4)])) < 1e-9
        assert abs(r["valid_loss"] - 0.5) < 1e-9


def test_launcher_propagates_the_first_failure(sb, tmp_path):
    from shifu_tensorflow_b200 import launcher as la
    up = _Upstream(); up.start()
    script = tmp_path / "fake_worker.py"
    script.write_text(FAKE_WORKER)
    env = _env(tmp_path, up.port, task=0)
    env["FAIL_RANK"] = "2"
    assert la.main(env=env, worker_cmd=[sys.executable, str(script)]) == 7
    up.srv.close()


def test_row_shard_gives_every_local_rank_the_same_count(sb):
    from shifu_tensorflow_b200 import trainer as tr
    x = np.arange(103 * 2, dtype=np.float32).reshape(103, 2); y = np.arange(103, dtype=np.float32)
    parts = [tr.row_shard("%d/4" % g, x, y) for g in range(4)]
    assert all(len(px) == len(py) == 25 for px, py in parts)
    seen = np.concatenate([py for _, py in parts])
    assert len(set(seen.tolist())) == 100 and set(seen.tolist()) <= set(range(103))
    np.testing.assert_array_equal(parts[1][0][:, 0] // 2, parts[1][1])
    with pytest.raises(ValueError):
        tr.row_shard("4/4", x)
