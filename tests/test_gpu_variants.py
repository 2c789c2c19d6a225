"""The other 1-D prefabs (lstm1, revlstm1, bidi0, bidi2), LSTM cell variants and Full<F> output layers of the reference
(clstm_prefab.cc:22-129, clstm.cc:382-389, 655-668 -- SURVEY.md section 8(f) rank 4) through clstm_b200_create_ex, against
the oracle's generic layer tree on the same seeded inputs.  fp32 within 1e-4 like the main parity tests."""
import numpy as np
import pytest

from clstm_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ffi():
    import clstm_b200
    clstm_b200.lib()
    return clstm_b200


def split(a, T):
    o = np.concatenate([[0], np.cumsum(T)])
    return [a[o[i]:o[i + 1]] for i in range(len(T))]


CASES = [
    # prefab, cell, output, ni, nh, nh2, nc
    ("lstm1", "NPLSTM", "SoftmaxLayer", 12, 16, 0, 9),            # register kernels, one direction
    ("revlstm1", "NPLSTM", "SoftmaxLayer", 12, 16, 0, 9),         # register kernels, reversed only
    ("bidi2", "NPLSTM", "SoftmaxLayer", 12, 16, 32, 9),           # two stacked blocks, both on register kernels
    ("bidi2", "NPLSTM", "SigmoidLayer", 10, 200, 16, 3),          # cluster block feeding a register block
    ("bidi", "LINNPLSTM", "TanhLayer", 8, 12, 0, 5),              # generic kernels, linear cell output
    ("lstm1", "RELUTANHNPLSTM", "LinearLayer", 8, 9, 0, 4),
    ("revlstm1", "RELUNPLSTM", "ReluLayer", 8, 9, 0, 4),
    ("bidi0", "RELU2NPLSTM", None, 8, 6, 0, 12),
    ("bidi", "NPLSTM", "SigmoidLayer", 8, 7, 0, 1),               # perplstm-style single sigmoid output
]


@pytest.mark.parametrize("prefab,cell,output,ni,nh,nh2,nc", CASES)
def test_prefab_forward_backward_parity(ffi, oracle, prefab, cell, output, ni, nh, nh2, nc):
    rng = np.random.default_rng(5)
    onet = oracle.PrefabOracle(prefab, ni, nh, nc, nh2=nh2, cell=cell, output=output, seed=0.21)
    gnet = ffi.Net(ni, nh, nc, prefab=prefab, nhidden2=nh2, cell=cell, output=output)
    assert gnet.nparams == onet.nparams
    scale = 0.15 if nh >= 100 else 0.4
    p = rng.normal(0, scale, onet.nparams).astype(np.float32)
    onet.set_params(p); gnet.set_params(p)
    assert np.array_equal(gnet.get_params(), p)                   # layout conversion round trip
    T = np.array([11, 1, 23], np.int32)
    x = rng.uniform(-1, 1, (int(T.sum()), ni)).astype(np.float32)
    out = gnet.forward(x, T)
    probe = rng.normal(0, 1, out.shape).astype(np.float32)
    gnet.clear_derivs()
    din = gnet.backward(probe)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for xx, oo, pp, dd in zip(split(x, T), split(out, T), split(probe, T), split(din, T)):
        o_out = onet.forward(xx)
        assert o_out.shape == oo.shape
        assert np.abs(o_out - oo).max() < TOL
        o_din = onet.backward(pp)
        assert np.abs(o_din - dd).max() < 2e-4 * max(1.0, np.abs(o_din).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < 3e-4 * max(1.0, np.abs(od).max())


def test_bidi2_trains_like_the_oracle(ffi, oracle):
    ni, nh, nh2, nc = 12, 16, 32, 9
    x, T, labels, L = synth.make_lines(3, (20, 40), ni, nc, seed=8)
    onet = oracle.PrefabOracle("bidi2", ni, nh, nc, nh2=nh2, seed=0.5)
    gnet = ffi.Net(ni, nh, nc, prefab="bidi2", nhidden2=nh2)
    p = np.random.default_rng(1).normal(0, 0.3, onet.nparams).astype(np.float32)
    onet.set_params(p); gnet.set_params(p)
    for _ in range(3):
        dec, out, al = gnet.train_step(x, T, labels, L, 1e-3, 0.9, want_out=True, want_aligned=True)
        for xx, ll, oo in zip(split(x, T), split(labels, L), split(out, T)):
            o_out, _ = onet.fwdbwd(xx, ll)
            assert np.abs(o_out - oo).max() < TOL
        onet.sgd_update(1e-3, 0.9)
    assert np.abs(onet.get_params() - gnet.get_params()).max() < TOL


def test_variant_errors(ffi):
    net = ffi.Net(8, 6, 4, prefab="lstm1", output="SigmoidLayer")
    x = np.zeros((5, 8), np.float32)
    net.forward(x, [5])
    with pytest.raises(ffi.Error, match="SoftmaxLayer"):
        net.ctc_align(np.array([1], np.int32), [1])
    with pytest.raises(ffi.Error, match="nblocks"):
        import ctypes as C
        from clstm_b200 import _ffi
        ex = _ffi.CfgEx(8, 4, 0, 3, (C.c_int * 2)(4, 4), (C.c_int * 2)(2, 2), 0, 0)
        h = C.c_void_p()
        _ffi._chk(_ffi.lib().clstm_b200_create_ex(C.byref(ex), C.byref(h)))
