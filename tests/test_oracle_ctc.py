"""Pins the CPU oracle's CTC aligner to the reference's own known-answer tests
(/root/reference/test-ctc.cc:47-109; stored in tests/golden/ctc_kat.json)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ctc_kat.json")


@pytest.mark.parametrize("double", [False, True])
def test_ctc_kats(oracle, double):
    kat = json.load(open(GOLD))
    for c in kat["cases"]:
        r = oracle.ctc_align_dense(np.array(c["outputs"]), np.array(c["targets"]), double=double)
        err = np.abs(r - np.array(c["expected"])).max()
        assert err < kat["tolerance"], (c["name"], err)


def test_labels_equals_dense(oracle):
    # ctc_align_targets(Sequence&, Sequence&, Classes&)-style one-hot targets == mktargets path (ctc.cc:136-157)
    rng = np.random.default_rng(0)
    T, nc, labels = 40, 11, [3, 3, 7, 1]
    out = rng.random((T, nc)).astype(np.float32)
    out /= out.sum(1, keepdims=True)
    S = 2 * len(labels) + 1
    tg = np.zeros((S, nc), np.float32)
    for s in range(S):
        tg[s, labels[(s - 1) // 2] if s % 2 else 0] = 1
    a = oracle.ctc_align_dense(out, tg)
    b = oracle.ctc_align_labels(out, labels)
    assert np.array_equal(a, b)
    assert np.allclose(a.sum(1), 1, atol=1e-5)


def test_argmax_ties_and_decode(oracle):
    # tensor.h:357-366: ties resolve to the LAST maximal index
    m = np.array([[0.2, 0.5, 0.5], [1, 1, 1], [0.9, 0.05, 0.05]], np.float32)
    assert oracle.argmax_rows(m).tolist() == [2, 2, 0]
    # ctc.cc:159-194: a run is emitted at its highest-probability frame; trailing unclosed run dropped
    o = np.zeros((8, 4), np.float32)
    for t, (c, p) in enumerate([(0, .9), (2, .6), (2, .8), (0, .9), (3, .7), (1, .75), (0, .9), (2, .9)]):
        o[t, c] = p
    cs, locs = oracle.trivial_decode(o)
    assert cs.tolist() == [2, 1] and locs.tolist() == [2, 5]
