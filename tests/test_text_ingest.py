"""Text ingest (sb_text_parse): the parsing state machine against Python's float() - on the CPU through
the host test hook (identical code, text_parse.cuh), on the GPU through the product entry point, and end to end against
the oracle's load_data restatement."""
import gzip
import random

import numpy as np
import pytest

from oracle import shifu_oracle as so


def _f32(s):
    return np.float32(float(s))


FAST = ["0", "-0", "0.0", "1", "-1", "+1.5", "3.14159", "-2.718281828", "1e5", "1E-5", "-4.25e+3", "000123.4500", ".5", "5.",
        "0.000001234", "123456789012345", "9007199254740992", "1e22", "1e-22", "  7.5  ", "7.5\r",
        "0.1", "0.2", "16777217", "33554433.0", "1e0", "-0e5", "1234567.890123", "-0.000000000000001"]
SLOW = ["4.9406564584124654", "1.17549435e-38", "3.4028235e38", "0.30000000000000004", "12345678901234567890123", "1e23", "1e-23", "nan", "inf", "-inf", "1_000", "0x10", "", ".", "e5", "1e", "--1", "1 2",
        "9007199254740993", "1.7976931348623157e308", "123456789.123456789123456789"]


def _parse_cells(sb, cells, host_debug, **kw):
    text = ("|".join(["1"] + cells) + "\n").encode()
    col_map = [sb.capi.COL_TARGET] + list(range(len(cells)))
    return sb.capi.text_parse(text, col_map, len(cells), host_debug=host_debug, **kw)


def test_number_parser_fast_path_is_bit_exact_vs_python_float(sb):
    X, y, w, flags, _ = _parse_cells(sb, FAST, True)
    assert flags == []
    want = np.array([_f32(c) for c in FAST], np.float32)
    np.testing.assert_array_equal(X[0].view(np.uint32), want.view(np.uint32))
    assert y[0] == 1.0 and w[0] == 1.0


def test_number_parser_declines_what_it_cannot_do_exactly(sb):
    X, y, w, flags, text = _parse_cells(sb, SLOW, True)
    assert sorted(f[1] for f in flags) == list(range(len(SLOW)))          # every slow cell flagged, none guessed
    for row, slot, off, ln in flags:
        assert text[off:off + ln].decode() == SLOW[slot]                 # and the flag points at exactly that cell


def test_number_parser_random_decimals_bit_exact(sb):
    rnd = random.Random(5)
    cells = []
    for _ in range(4000):
        kind = rnd.randrange(4)
        if kind == 0:
            cells.append(repr(rnd.gauss(0, 1)))                          # 17 significant digits -> mostly slow path
        elif kind == 1:
            cells.append("%.6f" % rnd.gauss(0, 3))
        elif kind == 2:
            cells.append("%.8e" % rnd.uniform(-1e6, 1e6))
        else:
            cells.append(str(rnd.randrange(-10 ** 9, 10 ** 9)))
    X, y, w, flags, text = _parse_cells(sb, cells, True)
    flagged = {f[1] for f in flags}
    assert 0 < len(flagged) < len(cells)
    for j, c in enumerate(cells):
        if j not in flagged:
            assert X[0, j].view(np.uint32) == _f32(c).view(np.uint32), c


def _make_text(n_rows, F, seed, with_weight):
    rnd = np.random.RandomState(seed)
    lines = []
    for i in range(n_rows):
        feats = ["%.6f" % v for v in rnd.randn(F)]
        if i % 7 == 3:
            feats[1] = repr(float(rnd.randn()))                          # long decimal: slow path through the flag list
        cols = [str(int(rnd.rand() < 0.3))] + feats + ["extra%d" % i]
        if with_weight:
            cols.append("%.3f" % (rnd.rand() * 4 - 1))                   # some negative weights -> 1.0
        lines.append("|".join(cols))
    return ("\n".join(lines) + "\n").encode()


def test_host_hook_full_lines_match_oracle_load_data(sb, tmp_path):
    F = 9
    raw = _make_text(301, F, 2, True)
    p = str(tmp_path / "part.gz")
    with gzip.open(p, "wb") as f:
        f.write(raw)
    want = so.load_data([p], list(range(1, F + 1)), 0, F + 2, 0.0, rng=random.Random(1))
    col_map = [sb.capi.COL_TARGET] + list(range(F)) + [sb.capi.COL_SKIP, sb.capi.COL_WEIGHT]
    X, y, w, flags, text = sb.capi.text_parse(raw, col_map, F, host_debug=True)
    for row, slot, off, ln in flags:
        assert slot == 1
        X[row, slot] = float(text[off:off + ln])
    np.testing.assert_array_equal(X, np.asarray(want["train_data"], np.float32))
    np.testing.assert_array_equal(y, np.asarray(want["train_target"], np.float32).ravel())
    np.testing.assert_array_equal(w, np.asarray(want["train_data_sample_weight"], np.float32).ravel())
    assert (w == 1.0).sum() > 40                                          # the negative ones were clamped


def test_short_line_is_reported_not_guessed(sb):
    col_map = [sb.capi.COL_TARGET, 0, 1, 2]
    X, y, w, flags, _ = sb.capi.text_parse(b"1|0.5|0.25|2\n0|0.5\n", col_map, 3, host_debug=True)
    assert [(f[0], f[1]) for f in flags] == [(1, -100)]


@pytest.mark.gpu
def test_gpu_parser_equals_host_hook_and_python(sb):
    F = 37
    raw = _make_text(5000, F, 3, True) + ("|".join(["1"] + (FAST + SLOW)[:F] + ["x", "2.5"]) + "\n").encode()
    col_map = [sb.capi.COL_TARGET] + list(range(F)) + [sb.capi.COL_SKIP, sb.capi.COL_WEIGHT]
    Xh, yh, wh, fh, _ = sb.capi.text_parse(raw, col_map, F, host_debug=True)
    Xg, yg, wg, fg, _ = sb.capi.text_parse(raw, col_map, F)
    np.testing.assert_array_equal(Xg.view(np.uint32), Xh.view(np.uint32))
    np.testing.assert_array_equal(yg, yh); np.testing.assert_array_equal(wg, wh)
    assert sorted(fg) == sorted(fh)
    assert Xg.shape == (5001, F) and yg[-1] == 1.0 and wg[-1] == 2.5


@pytest.mark.gpu
def test_load_data_gpu_equals_load_data(sb, tmp_path):
    from shifu_tensorflow_b200 import trainer as tr
    F = 12
    files = []
    for k in range(2):
        p = str(tmp_path / ("part-%d.gz" % k))
        with gzip.open(p, "wb") as f:
            f.write(_make_text(400 + 31 * k, F, 10 + k, True))
        files.append(p)
    a = tr.load_data(",".join(files), list(range(1, F + 1)), 0, F + 2, 0.25, rng=random.Random(9))
    b = tr.load_data_gpu(",".join(files), list(range(1, F + 1)), 0, F + 2, 0.25, rng=random.Random(9))
    for k in ("train_data", "valid_data", "train_target", "valid_target", "train_data_sample_weight", "valid_data_sample_weight"):
        assert isinstance(b[k], sb.capi.DeviceArray)                  # the parsed set stays on the device ...
        got = b[k].numpy()
        np.testing.assert_array_equal(np.asarray(a[k], np.float32).reshape(got.shape), got, err_msg=k)   # ... with the same bits
    assert a["feature_count"] == b["feature_count"] == F


@pytest.mark.gpu
def test_device_resident_set_feeds_the_trainer_without_a_host_copy(sb):
    """sb_text_parse_device -> device gather (split) -> sb_trainer_load_dataset / eval_loss on DEVICE pointers: same losses as
    the host-array path"""
    from oracle import shifu_oracle as so
    F = 12
    text = _make_text(600, F, 3, True)
    col_map = [sb.capi.COL_TARGET] + list(range(F)) + [sb.capi.COL_SKIP, sb.capi.COL_WEIGHT]
    Xh, yh, wh, fl_h, _ = sb.capi.text_parse(text, col_map, F)
    Xd, yd, wd, fl_d, _, kms = sb.capi.text_parse_device(text, col_map, F)
    assert sorted(fl_h) == sorted(fl_d) and kms > 0          # (the flag list is appended with an atomic counter: order varies)
    np.testing.assert_array_equal(Xd.numpy(), Xh); np.testing.assert_array_equal(yd.numpy(), yh); np.testing.assert_array_equal(wd.numpy(), wh)
    rows = np.arange(0, 600, 3)
    np.testing.assert_array_equal(Xd.take_rows(rows).numpy(), Xh[rows])
    desc = sb.make_desc(F, [16, 8], [so.ACT_RELU, so.ACT_TANH], optimizer=so.OPT_SGD, learning_rate=0.1, max_batch=128, precision=sb.PREC_BF16)
    with sb.Trainer(desc) as a, sb.Trainer(desc) as b:
        for t in (a, b):
            t.init_xavier(5)
        a.load_dataset(Xh, yh, wh)
        b.load_dataset(Xd, yd, wd)
        for k in range(4):
            assert a.step_resident(k * 128, 128) == b.step_resident(k * 128, 128)
        # (the evaluation loss is summed with fp32 atomics across CTAs: equal up to the summation order)
        assert abs(a.eval_loss(Xh, yh, wh) - b.eval_loss(Xd, yd, wd)) <= 1e-6
