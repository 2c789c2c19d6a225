"""Checks the CPU oracle's LSTM / bidi restatement with the reference's own gradient-check recipe
(/root/reference/test-deriv.cc:103-176: loss = sum(outputs * targets_delta), central differences,
inputs from the cos(3.7k) generator :26-33) and the LCG initialiser (batches.cc:11-52)."""
import math

import numpy as np
import pytest


def cosgen(shape, k0=0):
    n = int(np.prod(shape))
    return np.cos(3.7 * (np.arange(n) + k0)).reshape(shape)


@pytest.mark.parametrize("kind,ni,nh,no", [(0, 7, 3, 3), (1, 7, 3, 3), (2, 7, 5, 3)])
def test_gradcheck(oracle, kind, ni, nh, no):
    T, bs = 11, 2
    net = oracle.Net64(kind, ni, nh, no, seed=0.222)
    p0 = net.get_params()
    rng = np.random.default_rng(1)
    net.set_params(p0 + 0.3 * rng.standard_normal(p0.size))  # leave the tiny-init regime
    p0 = net.get_params()
    x = cosgen((T, bs, ni))
    out = net.forward(x)
    dout = cosgen(out.shape, 1000)
    din, dp = net.backward(dout)

    def loss(xx, pp):
        net.set_params(pp)
        return float((net.forward(xx) * dout).sum())

    h = 1e-5
    worst = 0.0
    for idx in rng.choice(x.size, 12, replace=False):
        xp = x.copy().ravel(); xm = x.copy().ravel()
        xp[idx] += h; xm[idx] -= h
        num = (loss(xp.reshape(x.shape), p0) - loss(xm.reshape(x.shape), p0)) / (2 * h)
        ana = din.ravel()[idx]
        worst = max(worst, abs(num - ana) / max(1e-8, abs(num) + abs(ana)))
    for idx in rng.choice(p0.size, 40, replace=False):
        pp = p0.copy(); pm = p0.copy()
        pp[idx] += h; pm[idx] -= h
        num = (loss(x, pp) - loss(x, pm)) / (2 * h)
        ana = dp[idx]
        if abs(num) + abs(ana) < 1e-9:
            continue
        worst = max(worst, abs(num - ana) / (abs(num) + abs(ana)))
    if kind == 2:
        # SoftmaxLayer backward deliberately omits the Jacobian (clstm_compute.cc:346-356; test-deriv.cc:207
        # does not test it either) so only the LSTM part can be checked: done by kinds 0 and 1.
        return
    assert worst < 1e-6, worst


def test_bidi_equals_manual_reverse(oracle):
    # Reversed{NPLSTM} == NPLSTM run on the time-reversed input, outputs reversed back (clstm.cc:461-469)
    T, bs, ni, nh = 9, 1, 5, 4
    a = oracle.Net64(0, ni, nh, nh, seed=0.3)
    b = oracle.Net64(1, ni, nh, nh, seed=0.3)
    x = cosgen((T, bs, ni))
    assert np.allclose(a.forward(x[::-1])[::-1], b.forward(x), atol=0)


def test_lcg_and_init(oracle):
    L = oracle.lib()
    L.oracle_seed(0.1)
    s = 0.1
    for _ in range(5):
        s = 189843.9384938 * s + 0.328340981343
        s -= math.floor(s)
        assert L.oracle_randu() == s
    # negbiased, scale 0.01 => uniform in [-0.02, 0.01] (clstm.cc:30-36, batches.cc:39-41)
    net = oracle.BidiOracle(48, 10, 12, seed=0.222)
    p = net.get_params()
    assert p.size == 2 * 4 * 10 * (1 + 48 + 10) + 12 * (1 + 20)
    assert p.min() >= -0.02 - 1e-9 and p.max() <= 0.01 + 1e-9
    # rinit draw order: i outer, j inner (batches.cc:37-38) while storage is column-major
    L.oracle_seed(0.5)
    draws = [L.oracle_randu() for _ in range(6)]
    L.oracle_seed(0.5)
    import ctypes
    a = np.zeros(6, np.float32)
    L.oracle_rinit(a.ctypes.data_as(oracle.f32p), 2, 3, 1.0, b"pos", 0.0)
    m = a.reshape(3, 2).T  # col-major 2x3
    assert np.allclose(m.ravel(), np.array(draws, np.float32))


def test_fwdbwd_and_update_semantics(oracle):
    # Params.d accumulates across fwdbwd calls and doubles as momentum buffer (clstm.cc:201-217)
    net = oracle.BidiOracle(8, 6, 5, seed=0.222)
    rng = np.random.default_rng(3)
    net.set_params(net.get_params() + 0.2 * rng.standard_normal(net.nparams).astype(np.float32))
    img = rng.random((30, 8)).astype(np.float32)
    out, al = net.fwdbwd(img, [1, 2, 4])
    assert np.allclose(out.sum(1), 1, atol=1e-5) and np.allclose(al.sum(1), 1, atol=1e-5)
    d1 = net.get_derivs()
    net.fwdbwd(img, [1, 2, 4])
    assert np.allclose(net.get_derivs(), 2 * d1, rtol=1e-5, atol=1e-7)
    p0 = net.get_params()
    net.sgd_update(1e-2, 0.9)
    assert np.allclose(net.get_params(), p0 + 1e-2 * np.clip(2 * d1, -100, 100), atol=1e-6)
    assert np.allclose(net.get_derivs(), 0.9 * 2 * d1, rtol=1e-5, atol=1e-7)


def test_training_reduces_ctc_error(oracle):
    # convergence smoke in the spirit of test-lstm.cc: repeated training on one line must align & decode it
    net = oracle.BidiOracle(8, 12, 6, seed=0.222)
    rng = np.random.default_rng(5)
    labels = [1, 3, 2, 5, 4]
    T = 40
    img = np.zeros((T, 8), np.float32)
    for k, c in enumerate(labels):
        img[4 + 7 * k: 9 + 7 * k, c] = 1.0
        img[4 + 7 * k: 9 + 7 * k, (c + 3) % 8] = 0.5
    for it in range(400):
        out, al = net.fwdbwd(img, labels)
        net.sgd_update(1e-2, 0.9)
    out = net.forward(img)
    cs, _ = oracle.trivial_decode(out)
    assert cs.tolist() == labels


def test_product_side_init_matches_reference_lcg(oracle):
    # clstm_b200.synth.reference_init (host side of the product) must reproduce the reference init bit for bit
    from clstm_b200 import synth
    net = oracle.BidiOracle(9, 6, 5, seed=0.222)
    assert np.array_equal(net.get_params(), synth.reference_init(9, 6, 5, seed=0.222))


@pytest.mark.parametrize("prefab,cell,output", [
    ("lstm1", "NPLSTM", "SigmoidLayer"), ("revlstm1", "LINNPLSTM", "SigmoidLayer"), ("bidi", "RELUTANHNPLSTM", "TanhLayer"),
    ("bidi2", "RELUNPLSTM", "LinearLayer"), ("bidi0", "RELU2NPLSTM", None), ("bidi2", "NPLSTM", "ReluLayer")])
def test_prefab_variants_gradcheck(oracle, prefab, cell, output):
    # the reference's gradient-check recipe (test-deriv.cc:103-176) on every prefab / cell / output-layer variant of the
    # oracle: central differences of sum(out * probe) against backward(), float32 => loose tolerance.  (SoftmaxLayer is
    # left out on purpose: its backward has no Jacobian upstream, clstm_compute.cc:346-356, so it is not a gradient.)
    ni, nh, nh2, nc, T = 5, 4, 3, 6, 7
    rng = np.random.default_rng(11)
    net = oracle.PrefabOracle(prefab, ni, nh, nc, nh2=nh2, cell=cell, output=output, seed=0.37)
    p = rng.normal(0, 0.4, net.nparams).astype(np.float32)
    net.set_params(p)
    x = rng.uniform(-1, 1, (T, ni)).astype(np.float32)
    probe = rng.normal(0, 1, (T, net.nc)).astype(np.float32)
    out = net.forward(x)
    assert out.shape == (T, net.nc)
    net.clear_derivs()
    din = net.backward(probe)
    g = net.get_derivs()
    eps = 2e-3
    for idx in rng.choice(net.nparams, 12, replace=False):
        q = p.copy(); q[idx] += eps; net.set_params(q); lp = float((net.forward(x) * probe).sum())
        q[idx] -= 2 * eps; net.set_params(q); lm = float((net.forward(x) * probe).sum())
        num = (lp - lm) / (2 * eps)
        assert abs(num - g[idx]) < 3e-2 * max(1.0, abs(num)), (prefab, idx, num, g[idx])
    net.set_params(p)
    for t, i in [(0, 0), (T - 1, ni - 1), (3, 2)]:
        xx = x.copy(); xx[t, i] += eps; lp = float((net.forward(xx) * probe).sum())
        xx[t, i] -= 2 * eps; lm = float((net.forward(xx) * probe).sum())
        num = (lp - lm) / (2 * eps)
        assert abs(num - din[t, i]) < 3e-2 * max(1.0, abs(num)), (prefab, t, i, num, din[t, i])


def test_bidi_forward_matches_independent_numpy_restatement(oracle):
    # SURVEY Appendix A.1/A.3 written out in numpy float64, straight from the prose (not from the oracle's loops):
    # per direction  src=[x_t; h_{t-1}], gi/gf/go = sigmoid, ci = tanh, c = ci*gi + gf*c_prev, h = tanh(c)*go,
    # second LSTM on the time-reversed input, outputs concatenated, softmax = limexp(W [1; h]) normalised.
    ni, nh, nc, T = 6, 5, 4, 9
    rng = np.random.default_rng(3)
    net = oracle.BidiOracle(ni, nh, nc, seed=0.2)
    p = rng.normal(0, 0.5, net.nparams)
    net.set_params(p.astype(np.float32))
    p = net.get_params().astype(np.float64)            # exactly the float32 values the oracle holds
    x = rng.uniform(-1, 1, (T, ni)).astype(np.float32)
    nf = ni + nh
    msz = nh * (1 + nf)

    def mats(block):                                   # walk_params order inside one LSTM: WCI, WGF, WGI, WGO (col-major)
        W = [block[k * msz:(k + 1) * msz].reshape(1 + nf, nh).T for k in range(4)]
        return {"ci": W[0], "gf": W[1], "gi": W[2], "go": W[3]}
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))

    def lstm(W, xs):
        h = np.zeros(nh); c = np.zeros(nh); out = []
        for t in range(len(xs)):
            src = np.concatenate([[1.0], xs[t], h])
            gi, gf, go = sig(W["gi"] @ src), sig(W["gf"] @ src), sig(W["go"] @ src)
            ci = np.tanh(W["ci"] @ src)
            c = ci * gi + (gf * c if t > 0 else 0.0)
            h = np.tanh(c) * go
            out.append(h)
        return np.array(out)
    xs = x.astype(np.float64)
    hf = lstm(mats(p[0:4 * msz]), xs)
    hb = lstm(mats(p[4 * msz:8 * msz]), xs[::-1])[::-1]
    W1 = p[8 * msz:].reshape(1 + 2 * nh, nc).T
    z = np.exp(np.clip(np.concatenate([np.ones((T, 1)), hf, hb], 1) @ W1.T, -30, 30))
    ref = z / z.sum(1, keepdims=True)
    assert np.abs(net.forward(x) - ref).max() < 2e-6    # float32 oracle vs float64 restatement
