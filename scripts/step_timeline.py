"""In-graph timeline of one steady-state training step: %globaltimer stamps written by CTA 0 of every GEMM of the
captured step (SB_STEP_TRACE=1).  Prints, per GEMM, when it entered / resolved its dependencies / finished, relative
to the first kernel's entry, and the gap to the previous kernel's exit.  Usage: python scripts/step_timeline.py [cfg1|cfg2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB_STEP_TRACE"] = "1"
import numpy as np
import shifu_tensorflow_b200 as sb

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
F, hidden, B, opt = dict(cfg1=(1000, [512, 256, 128], 4096, sb.OPT_ADAM), cfg2=(2000, [1024, 512, 256], 8192, sb.OPT_MOMENTUM))[name]
desc = sb.make_desc(F, hidden, [sb.ACT_RELU] * len(hidden), optimizer=opt, learning_rate=0.01, max_batch=B, precision=sb.PREC_BF16)
nb = 8
rng = np.random.RandomState(0)
X = rng.standard_normal((nb * B, F)).astype(np.float32); y = (rng.uniform(size=(nb * B, 1)) < 0.2).astype(np.float32)
with sb.Trainer(desc) as t:
    t.init_xavier(1); t.load_dataset(X, y, None)
    for rep in range(3):
        for i in range(50): t.step_resident_async((i % nb) * B, B)
        t.sync()
        names, st = t.debug_step_trace()
        st = st.astype(np.int64)
        order = np.argsort(st[:, 0])
        t0 = st[order[0], 0]
        print("--- %s step timeline (us since first GEMM entry), repetition %d" % (name, rep))
        print("%-8s %8s %8s %8s %8s %8s %8s | %s" % ("kernel", "entry", "deps_ok", "tma0", "acc0", "epi0", "exit", "entry - latest earlier exit"))
        def rel(v):
            return (v - t0) / 1e3 if v > 0 else float("nan")
        for k in order:
            s = st[k]
            earlier = [st[j, 8] for j in order if 0 < st[j, 8] <= s[0] and j != k]
            gap = (s[0] - max(earlier)) / 1e3 if earlier else float("nan")
            print("%-8s %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f | %6.2f" % (names[k], rel(s[0]), rel(s[2]), rel(s[3]), rel(s[6]), rel(s[7]), rel(s[8]), gap))
