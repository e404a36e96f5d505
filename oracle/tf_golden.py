#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY - generator of TF-written golden vectors for the oracle.

No TensorFlow exists in the build container or on the GPU box, so this script has NOT been run here and the oracle stays
"parity unpinned" against real TF until somebody runs

    python oracle/tf_golden.py --out tests/golden/tf_golden.npz          # any machine with tensorflow 1.x or 2.x

and commits the file; tests/test_oracle.py::test_oracle_matches_tf_golden picks it up automatically (it is skipped while
the file is absent).  The graph below is a py3 / tf.compat.v1 transcription of the reference's `model()`,
`nn_layer`, `get_activation_fun` and loss/optimizer construction (res/ssgd_monitor.py:57-88, 110-144) with the two
py2-isms replaced (tf.contrib.layers.xavier_initializer -> injected initial values, so the run is deterministic; the
SyncReplicasOptimizer wrapper is dropped because a single local session cannot host its token queue - its bookkeeping
is checked separately in tests/test_host_mirrors.py).  What it pins:

    fwd/loss/grads   per-layer activations, loss (MSE SUM_BY_NONZERO_WEIGHTS and sigmoid-CE), every gradient
    optimizers       three updates each of tf.train.{Adadelta,Adam,GradientDescent,Momentum}Optimizer on those gradients
                     -> resolves the ApplyAdadelta evaluation-order switch (OptConfig.adadelta_var_uses_new_accum_update)
    leaky_relu       TF's default alpha
"""
import argparse
import sys

import numpy as np


def build_and_run(out_path):
    import tensorflow as tf
    tf1 = tf.compat.v1 if hasattr(tf, "compat") and hasattr(tf.compat, "v1") else tf
    if hasattr(tf1, "disable_eager_execution"):
        tf1.disable_eager_execution()
    sys.path.insert(0, ".")
    from oracle import shifu_oracle as so

    acts_tf = {so.ACT_SIGMOID: tf.nn.sigmoid, so.ACT_TANH: tf.nn.tanh, so.ACT_RELU: tf.nn.relu, so.ACT_LEAKYRELU: tf.nn.leaky_relu}
    F, hidden, acts = 20, [16, 8], [so.ACT_LEAKYRELU, so.ACT_TANH]
    net = so.NetDesc(F, hidden, acts)
    params = so.xavier_init(net, 123)
    X, y, w = so.synth_batch(64, F, 7, weights="mixed")
    out = {"F": F, "hidden": np.array(hidden), "acts": np.array(acts), "X": X, "y": y, "w": w, "tf_version": tf.__version__}
    for i, p in enumerate(params):
        out["param%d" % i] = p

    for loss_name, loss_id in (("mse", so.LOSS_MSE), ("ce", so.LOSS_SIGMOID_CE)):
        for opt_name, make in (("adadelta", lambda: tf1.train.AdadeltaOptimizer(0.5, rho=0.95, epsilon=1e-8)),
                               ("adam", lambda: tf1.train.AdamOptimizer(0.01)),
                               ("sgd", lambda: tf1.train.GradientDescentOptimizer(0.1)),
                               ("momentum", lambda: tf1.train.MomentumOptimizer(0.1, 0.9))):
            g = tf1.Graph()
            with g.as_default():
                x_ = tf1.placeholder(tf.float32, [None, F], name="shifu_input_0")        # ssgd_monitor.py:207
                y_ = tf1.placeholder(tf.float32, [None, 1])
                w_ = tf1.placeholder(tf.float32, [None, 1])
                vs, a, layers = [], x_, []
                for l, act in enumerate(acts):                                            # nn_layer, :57-71
                    W = tf1.Variable(params[2 * l], name="weight_hidden_layer%d" % l)
                    b = tf1.Variable(params[2 * l + 1], name="biases_hidden_layer%d" % l)
                    a = acts_tf[act](tf.matmul(a, W) + b)
                    vs += [W, b]; layers.append(a)
                Wo = tf1.Variable(params[-2], name="weight_shifu_output_0")
                bo = tf1.Variable(params[-1], name="biases_shifu_output_0")
                vs += [Wo, bo]
                z = tf.matmul(a, Wo) + bo
                yhat = tf.nn.sigmoid(z, name="shifu_output_0")                            # :121
                if loss_id == so.LOSS_MSE:
                    loss = tf1.losses.mean_squared_error(predictions=yhat, labels=y_, weights=w_)   # :129
                else:
                    loss = tf1.losses.sigmoid_cross_entropy(multi_class_labels=y_, logits=z, weights=w_)
                opt = make()
                gv = opt.compute_gradients(loss, var_list=vs)
                train = opt.apply_gradients(gv)
                with tf1.Session() as sess:
                    sess.run(tf1.global_variables_initializer())
                    fd = {x_: X, y_: y, w_: w}
                    key = "%s_%s_" % (loss_name, opt_name)
                    vals = sess.run([loss, yhat] + layers + [gi for gi, _ in gv], fd)
                    out[key + "loss"], out[key + "yhat"] = vals[0], vals[1]
                    for l in range(len(layers)):
                        out[key + "act%d" % l] = vals[2 + l]
                    for i, gi in enumerate(vals[2 + len(layers):]):
                        out[key + "grad%d" % i] = gi
                    for step in range(3):
                        sess.run(train, fd)
                        for i, v in enumerate(sess.run(vs)):
                            out[key + "step%d_param%d" % (step + 1, i)] = v
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "with TF", tf.__version__)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/tf_golden.npz")
    build_and_run(ap.parse_args().out)
