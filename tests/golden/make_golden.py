"""Generates tests/golden/*.json|npz from the reference's OWN artefacts, in the build container only
(/root/reference does not exist on the GPU box).  Run:  python tests/golden/make_golden.py

dummydl_known_answers.json : forward outputs of the reference's SavedModel fixture
    (shifu-tensorflow-eval/src/test/resources/dummydl, loaded by TensorflowModelTest.java:35-60) computed by the
    oracle reader + fp32 numpy forward.  NOT TF-verified (TF is not installable here); they agree with the values
    SURVEY.md section 8c lists, which were derived independently.
dummydl_op_attrs.json : for every op type in the fixture's GraphDef (written by a real TF 1.x), the attribute keys its
    nodes carry, plus the attribute keys of the SignatureDef / SaverDef plumbing - the structural reference our own
    SavedModel writer is linted against (tests/test_formats.py).
dummydl_head.npz : the first 3 and the last layer of that model + a 16-row input/output pair, small enough to
    commit, so the GPU box can check the scorer kernels against the fixture's real weights.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tf_formats as tff  # noqa: E402

FIXTURE = "/root/reference/shifu-tensorflow-eval/src/test/resources/dummydl"


def main():
    layers, names = tff.extract_mlp(FIXTURE, "dense_46_input", "dense_66/Sigmoid")
    cases = []
    for value in (0.5, 0.0):
        X = np.full((1, 1522), value, np.float32)
        cases.append({"input_fn": "const", "value": value, "expected": [float(v) for v in tff.mlp_forward(layers, X).ravel()]})
    X = np.random.RandomState(0).rand(4, 1522).astype(np.float32)
    cases.append({"input_fn": "rand", "seed": 0, "rows": 4, "expected": [float(v) for v in tff.mlp_forward(layers, X).ravel()]})
    json.dump({"source": "dummydl fixture via oracle/tf_formats.py", "input": "dense_46_input", "output": "dense_66/Sigmoid",
               "cases": cases}, open(os.path.join(HERE, "dummydl_known_answers.json"), "w"), indent=1)
    # reduced model: layers 0,1,2 + output layer (1522->100->100->100->1), fp16-free, ~650 KB compressed
    sub = [layers[0], layers[1], layers[2], layers[-1]]
    X = np.random.RandomState(1).rand(16, 1522).astype(np.float32)
    Y = tff.mlp_forward(sub, X)
    np.savez_compressed(os.path.join(HERE, "dummydl_head.npz"), X=X, Y=Y,
                        **{"W%d" % i: l[0] for i, l in enumerate(sub)}, **{"b%d" % i: l[1] for i, l in enumerate(sub)})
    nodes, sigs = tff.read_graph_nodes(os.path.join(FIXTURE, "saved_model.pb"))
    ops = {}
    for _name, (op, _inputs, attrs) in nodes.items():
        ops.setdefault(op, set()).update(attrs.keys())
    json.dump({"source": "dummydl/saved_model.pb (TF-written GraphDef), via oracle/tf_formats.read_graph_nodes",
               "op_attr_keys": {op: sorted(keys) for op, keys in sorted(ops.items())}, "signatures": sorted(sigs)},
              open(os.path.join(HERE, "dummydl_op_attrs.json"), "w"), indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
