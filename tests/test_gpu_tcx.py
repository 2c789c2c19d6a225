"""GPU tests of the cluster-resident tensor-core recurrence (clstm_b200/csrc/lstm_tcx.cu: a thread-block cluster owns 16 text
lines of one direction, the recurrent matrix is split over the CTAs' shared memories, tcgen05 products with the lines on the UMMA
N dimension, h / partial sums exchanged through distributed shared memory with st.async + mbarrier).
Reference semantics: GenericNPLSTM::forward / backward, /root/reference/clstm.cc:600-653.
Forced through CLSTM_B200_LSTM=tcx (read when a net is created); compared with the CPU oracle at the 1e-4 bar of BASELINE.json
and, on the device, with the fp32 SIMT kernels."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from clstm_b200 import synth  # noqa: E402

TOL = 1e-4


@pytest.fixture(scope="module")
def ffi():
    import clstm_b200
    clstm_b200.lib()
    return clstm_b200


class forced:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.old = os.environ.get("CLSTM_B200_LSTM")
        os.environ["CLSTM_B200_LSTM"] = self.mode

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("CLSTM_B200_LSTM", None)
        else:
            os.environ["CLSTM_B200_LSTM"] = self.old


def split(a, T):
    offs = np.concatenate([[0], np.cumsum(T)])
    return [a[offs[i]:offs[i + 1]] for i in range(len(T))]


# nhidden, lines, tmin, tmax: clusters of 2..15 CTAs (32 units each, the last one partly filled), weight slice in shared memory
# (nhidden <= 256) and in tensor memory (wider), fewer lines than one group, several groups per cluster (more groups than
# resident clusters), ragged down to T = 1
AB_CASES = [(200, 16, 3, 8), (200, 40, 1, 20), (100, 20, 5, 12), (256, 33, 4, 9), (128, 128, 10, 30), (96, 7, 2, 15),
            (160, 50, 1, 25), (200, 700, 2, 6), (64, 130, 3, 9), (400, 40, 1, 20), (320, 33, 4, 9), (480, 17, 2, 7), (288, 16, 5, 9)]


@pytest.mark.parametrize("no,B,t0,t1", AB_CASES)
def test_tcx_recurrence_matches_simt_kernels(ffi, no, B, t0, t1):
    r = ffi.selftest_lstm(no, B, t0, t1, seed=7, cluster_resident=True)
    assert max(r["d_gates"], r["d_cell"], r["d_h"], r["d_hprev"]) < 2e-5, r
    assert r["d_delta_rel"] < 1e-4, r


PARITY = [
    # ni, nh, nc, B, T, weights
    (48, 200, 83, 5, (1, 60), "trained"),     # BASELINE config 3 width, ragged incl. T = 1
    (48, 200, 83, 35, (30, 50), "init"),      # three line groups
    (48, 100, 83, 20, (20, 45), "trained"),   # config 2 width
    (48, 256, 83, 3, (10, 25), "trained"),
    (48, 400, 83, 5, (20, 45), "trained"),    # config 4 width: weight slice in tensor memory
]


@pytest.mark.parametrize("ni,nh,nc,B,T,weights", PARITY)
def test_tcx_parity_with_oracle(ffi, oracle, ni, nh, nc, B, T, weights):
    x, Ts, labels, L = synth.make_lines(B, T, ni, nc, seed=3)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    if weights == "trained":
        onet.set_params(synth.trained_like(onet.nparams, 0.3 if nh <= 200 else 0.3 * (200.0 / nh) ** 0.5, seed=7))
    with forced("tcx"):
        gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    out = gnet.forward(x, Ts)
    assert gnet.lstm_variant == "tcx"
    xs, outs = split(x, Ts), split(out, Ts)
    for b in range(B):
        assert np.abs(onet.forward(xs[b]) - outs[b]).max() < TOL
    rng = np.random.default_rng(5)
    deltas = (rng.standard_normal(out.shape) * 0.1).astype(np.float32)
    gnet.clear_derivs()
    din = gnet.backward(deltas)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for b in range(B):
        onet.forward(xs[b])
        o_din = onet.backward(split(deltas, Ts)[b])
        assert np.abs(o_din - split(din, Ts)[b]).max() < TOL * max(1.0, np.abs(o_din).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < TOL * max(1.0, np.abs(od).max())
    # CTC on these outputs: alignment indices bit-exact against the oracle run on the same outputs
    al = gnet.ctc_align(labels, L)
    amax = gnet.argmax(1)
    dec = gnet.decode(1)
    labs = split(labels, L)
    for b in range(B):
        o_al = oracle.ctc_align_labels(outs[b], labs[b])
        assert np.abs(o_al - split(al, Ts)[b]).max() < 2e-5
        assert np.array_equal(oracle.argmax_rows(o_al), split(amax, Ts)[b])
        cs, locs = oracle.trivial_decode(o_al)
        assert np.array_equal(cs, dec[b][0]) and np.array_equal(locs, dec[b][1])


def test_tcx_training_steps_track_oracle(ffi, oracle):
    ni, nh, nc, B = 48, 200, 83, 6
    x, Ts, labels, L = synth.make_lines(B, (30, 60), ni, nc, seed=11)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    onet.set_params(synth.trained_like(onet.nparams, 0.3, seed=7))
    with forced("tcx"):
        gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    for _ in range(3):   # the fp16 hi/lo weight slices must follow every update
        gnet.train_step(x, Ts, labels, L, 1e-3, 0.9)
        onet.train_lines(x, Ts, labels, L, 1e-3, 0.9, threads=1, reps=1)
    assert gnet.lstm_variant == "tcx"
    assert np.abs(gnet.get_params() - onet.get_params()).max() < 1e-4


# other prefab wirings (clstm_prefab.cc:22-129): one direction only (the launch covers direction slot 0 or 1 alone, H rows are
# nhidden wide) and two stacked blocks of different widths
PREFABS = [("lstm1", 64, 0), ("revlstm1", 64, 0), ("bidi2", 64, 96)]


@pytest.mark.parametrize("prefab,nh,nh2", PREFABS)
def test_tcx_recurrence_in_other_topologies(ffi, oracle, prefab, nh, nh2):
    ni, nc = 12, 9
    rng = np.random.default_rng(5)
    onet = oracle.PrefabOracle(prefab, ni, nh, nc, nh2=nh2, cell="NPLSTM", output="SoftmaxLayer", seed=0.21)
    with forced("tcx"):
        gnet = ffi.Net(ni, nh, nc, prefab=prefab, nhidden2=nh2)
    p = rng.normal(0, 0.25, onet.nparams).astype(np.float32)
    onet.set_params(p); gnet.set_params(p)
    T = np.array([11, 1, 23, 7, 16], np.int32)
    x = rng.uniform(-1, 1, (int(T.sum()), ni)).astype(np.float32)
    out = gnet.forward(x, T)
    assert gnet.lstm_variant == "tcx"
    probe = rng.normal(0, 1, out.shape).astype(np.float32)
    gnet.clear_derivs()
    din = gnet.backward(probe)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for xx, oo, pp, dd in zip(split(x, T), split(out, T), split(probe, T), split(din, T)):
        assert np.abs(onet.forward(xx) - oo).max() < TOL
        o_din = onet.backward(pp)
        assert np.abs(o_din - dd).max() < 2e-4 * max(1.0, np.abs(o_din).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < 3e-4 * max(1.0, np.abs(od).max())
