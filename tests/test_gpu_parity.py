"""GPU parity tests (run on the B200 box with -m gpu).  Every test drives the CUDA path through the C ABI
(ctypes -> libclstm_b200.so) and compares with the CPU oracle on the same seeded inputs.

Tolerances: fp32 values within 1e-4 absolute of the oracle (BASELINE.json north_star); CTC alignment indices
(per-column argmax of `aligned`, trivial_decode classes and locations) bit-exact."""
import json
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


def split(a, T):
    offs = np.concatenate([[0], np.cumsum(T)])
    return [a[offs[i]:offs[i + 1]] for i in range(len(T))]


def make_pair(ffi, oracle, ni, nh, nc, weights, seed=0.222):
    onet = oracle.BidiOracle(ni, nh, nc, seed=seed)
    if weights == "trained":
        onet.set_params(synth.trained_like(onet.nparams, 0.3, seed=7))
    gnet = ffi.Net(ni, nh, nc)
    assert gnet.nparams == onet.nparams
    gnet.set_params(onet.get_params())
    return onet, gnet


CASES = [
    # ni, nh, nc, B, T, weights
    (48, 100, 83, 3, (40, 70), "init"),      # register-resident kernels
    (48, 100, 83, 3, (40, 70), "trained"),
    (48, 50, 83, 2, (30, 50), "trained"),     # regs, K padded to a multiple of 4
    (7, 5, 4, 2, (9, 14), "trained"),         # generic kernels, test-deriv.cc-sized net
    (48, 24, 30, 4, (1, 40), "trained"),      # generic, ragged incl. very short lines
    (48, 200, 83, 3, (1, 60), "trained"),     # 4-CTA cluster kernels (BASELINE config 3 width)
    (48, 400, 83, 2, (20, 45), "trained"),    # 16-CTA cluster kernels (BASELINE config 4 width)
]


@pytest.mark.parametrize("ni,nh,nc,B,T,weights", CASES)
def test_forward_ctc_backward_parity(ffi, oracle, ni, nh, nc, B, T, weights):
    x, Ts, labels, L = synth.make_lines(B, T, ni, nc, seed=3)
    # transcripts must fit the line: the aligner needs no minimum, but keep L <= T
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, weights)
    assert np.array_equal(gnet.get_params(), onet.get_params())           # layout round trip is exact
    out = gnet.forward(x, Ts)
    aligned = gnet.ctc_align(labels, L)
    din = gnet.backward()
    gd = gnet.get_derivs()
    amax = gnet.argmax(1)
    dec_al = gnet.decode(1)
    dec_out = gnet.decode(0)
    xs, outs, als, dins = split(x, Ts), split(out, Ts), split(aligned, Ts), split(din, Ts)
    labs = split(labels, L)
    onet.clear_derivs()
    for b in range(B):
        o_out, o_al = onet.fwdbwd(xs[b], labs[b])
        assert np.abs(o_out - outs[b]).max() < TOL
        # the aligner amplifies output differences; compare it on identical inputs (the GPU's own outputs)
        o_al_same = oracle.ctc_align_labels(outs[b], labs[b])
        assert np.abs(o_al_same - als[b]).max() < 2e-5
        assert np.abs(o_al - als[b]).max() < 5e-3
        assert np.array_equal(oracle.argmax_rows(o_al_same), split(amax, Ts)[b])
        cs, locs = oracle.trivial_decode(o_al_same)
        assert np.array_equal(cs, dec_al[b][0]) and np.array_equal(locs, dec_al[b][1])
        cs, locs = oracle.trivial_decode(outs[b])
        assert np.array_equal(cs, dec_out[b][0]) and np.array_equal(locs, dec_out[b][1])
    od = onet.get_derivs()
    scale = max(1.0, np.abs(od).max())
    assert np.abs(od - gd).max() < 5e-3 * scale      # end to end incl. the aligner's amplification
    # backward in isolation: same deltas into both
    rng = np.random.default_rng(5)
    deltas = (rng.standard_normal(out.shape) * 0.1).astype(np.float32)
    gnet.clear_derivs()
    gnet.forward(x, Ts)
    din = gnet.backward(deltas)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for b in range(B):
        onet.forward(xs[b])
        o_din = onet.backward(split(deltas, Ts)[b])
        assert np.abs(o_din - split(din, Ts)[b]).max() < TOL * max(1.0, np.abs(o_din).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < TOL * max(1.0, np.abs(od).max())


def test_ctc_known_answers_on_gpu(ffi):
    # the reference's own golden vectors (test-ctc.cc:47-109) through the CUDA aligner
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_kat.json")))
    for c in kat["cases"]:
        outs = np.array(c["outputs"], np.float32)
        tg = np.array(c["targets"])
        states = tg.argmax(1).astype(np.int32)
        net = ffi.Net(8, 4, outs.shape[1])
        al = net.ctc_align_states(outs, [outs.shape[0]], states, [states.size])
        assert np.abs(al - np.array(c["expected"])).max() < kat["tolerance"], c["name"]


def test_training_steps_track_oracle(ffi, oracle):
    ni, nh, nc, B = 48, 100, 83, 4
    x, Ts, labels, L = synth.make_lines(B, (40, 60), ni, nc, seed=11)
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, "trained")
    lr, mom = 1e-3, 0.9
    for step in range(3):
        dec, _, _ = gnet.train_step(x, Ts, labels, L, lr, mom)
        onet.train_lines(x, Ts, labels, L, lr, mom, threads=1, reps=1)
    gp, op = gnet.get_params(), onet.get_params()
    assert np.abs(gp - op).max() < 1e-4
    gd, od = gnet.get_derivs(), onet.get_derivs()
    assert np.abs(gd - od).max() < 5e-3 * max(1.0, np.abs(od).max())


def test_derivs_accumulate_and_momentum(ffi):
    # Params.d accumulates over backward calls and doubles as the momentum buffer (clstm.cc:201-217)
    ni, nh, nc = 48, 16, 10
    x, Ts, labels, L = synth.make_lines(2, 30, ni, nc, seed=2)
    net = ffi.Net(ni, nh, nc)
    net.set_params(synth.trained_like(net.nparams, 0.3))
    net.forward(x, Ts); net.ctc_align(labels, L); net.backward()
    d1 = net.get_derivs()
    net.forward(x, Ts); net.ctc_align(labels, L); net.backward()
    assert np.allclose(net.get_derivs(), 2 * d1, rtol=1e-5, atol=1e-6)
    p0 = net.get_params()
    net.sgd_update(1e-2, 0.9, 0.05)
    clipped = np.clip(2 * d1, -0.05, 0.05)
    assert np.allclose(net.get_params(), p0 + 1e-2 * clipped, atol=1e-6)
    assert np.allclose(net.get_derivs(), 0.9 * clipped, rtol=1e-5, atol=1e-7)


def test_full_size_properties(ffi):
    # BASELINE config 2 at full size (nh=100, B=32, T=500): size-independent properties instead of the oracle
    ni, nh, nc, B, T = 48, 100, 83, 32, 500
    x, Ts, labels, L = synth.make_lines(B, T, ni, nc, seed=4)
    net = ffi.Net(ni, nh, nc)
    net.set_params(synth.trained_like(net.nparams, 0.1))
    out = net.forward(x, Ts)
    al = net.ctc_align(labels, L)
    assert np.isfinite(out).all() and np.isfinite(al).all()
    assert np.abs(out.sum(1) - 1).max() < 1e-5 and np.abs(al.sum(1) - 1).max() < 1e-5   # check_normalized batches.h:159
    # aligned mass only on blank and the line's own transcript classes
    for b, (a, lab) in enumerate(zip(split(al, Ts), split(labels, L))):
        mask = np.ones(nc, bool); mask[0] = False; mask[lab] = False
        assert a[:, mask].max() == 0.0
    # batch independence: line 5 alone gives the same outputs as inside the batch
    o5 = net.forward(split(x, Ts)[5], [T])
    assert np.abs(o5 - split(out, Ts)[5]).max() < 1e-6
    # reversal symmetry of the wiring: swapping the two directions' weights == time-reversing the input
    p = net.get_params()
    blk = 4 * nh * (1 + ni + nh)
    swapped = p.copy(); swapped[:blk] = p[blk:2 * blk]; swapped[blk:2 * blk] = p[:blk]
    w1 = p[2 * blk:].reshape(1 + 2 * nh, nc).copy()          # col-major nc x (1+2nh) => rows here are columns
    w1s = w1.copy(); w1s[1:1 + nh] = w1[1 + nh:]; w1s[1 + nh:] = w1[1:1 + nh]
    swapped[2 * blk:] = w1s.ravel()
    net.set_params(swapped)
    xr = np.concatenate([xx[::-1] for xx in split(x, Ts)], 0)
    outr = net.forward(xr, Ts)
    outr = np.concatenate([oo[::-1] for oo in split(outr, Ts)], 0)
    assert np.abs(outr - out).max() < 1e-5


def test_error_behaviour(ffi):
    net = ffi.Net(48, 16, 10)
    with pytest.raises(ffi.Error):
        net.set_params(np.zeros(3, np.float32))                      # size mismatch (clstm.cc:871)
    with pytest.raises(ffi.Error):
        net.backward()                                                # backward before forward
    x, Ts, labels, L = synth.make_lines(1, 20, 48, 10, seed=0)
    net.forward(x, Ts)
    with pytest.raises(ffi.Error):
        net.ctc_align(np.array([12], np.int32), [1])                  # label out of range
    with pytest.raises(ffi.Error):
        net.forward(x[:0], [0])                                       # empty line


@pytest.mark.parametrize("nh", [100, 24])
def test_tcgen05_dense_products_match_simt(ffi, nh):
    # 3xTF32 on the 5th-gen tensor cores must reproduce the fp32 SIMT products to fp32-level accuracy
    net = ffi.Net(48, nh, 83)
    err = net.selftest_gemm()
    assert len(err) == 6
    assert err.max() < 1e-5, err


@pytest.mark.parametrize("nh", [100, 400])
def test_tma_fed_gemm_matches_simt(ffi, nh, monkeypatch):
    # the persistent TMA-fed GEMM on fp16 hi/lo planes (gemm_x.cu), forced for every product size (CLSTM_B200_GEMM=x is read
    # when the net is created): forward products with bias, both derivative products with column blocks + ones column
    monkeypatch.setenv("CLSTM_B200_GEMM", "x")
    net = ffi.Net(48, nh, 83)
    err = net.selftest_gemm()
    assert len(err) == 6
    # cases 0, 1 (forward products) and 4, 5 (derivative products) run on gemm_x.cu; 2, 3 (weights used untransposed) stay on the
    # 3xTF32 kernels, whose truncation split reaches 1.3e-5 at K = 1600
    assert err[[0, 1, 4, 5]].max() < 1e-5 and err.max() < 2e-5, err


def test_tma_fed_gemm_training_steps_track_oracle(ffi, oracle, monkeypatch):
    # whole training steps with every dense product on the TMA-fed GEMM (ragged batch, K and N tails, split-K derivative products)
    monkeypatch.setenv("CLSTM_B200_GEMM", "x")
    ni, nh, nc, B = 48, 100, 83, 7
    x, Ts, labels, L = synth.make_lines(B, (5, 70), ni, nc, seed=13)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    onet.set_params(synth.trained_like(onet.nparams, 0.3, seed=7))
    gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    for _ in range(3):
        gnet.train_step(x, Ts, labels, L, 1e-3, 0.9)
        onet.train_lines(x, Ts, labels, L, 1e-3, 0.9, threads=1, reps=1)
    assert np.abs(gnet.get_params() - onet.get_params()).max() < 1e-4


def test_long_lines_wide_net(ffi, oracle):
    # BASELINE config 3 shape in miniature: nhidden=200 (cluster recurrent kernels), ragged lines up to T=1500,
    # transcripts up to 75 labels (lattice lanes own 5 states each); alignment indices must be bit-exact.
    ni, nh, nc = 48, 200, 83
    x, Ts, labels, L = synth.make_lines(3, (600, 1500), ni, nc, seed=9)
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, "init")
    out = gnet.forward(x, Ts)
    aligned = gnet.ctc_align(labels, L)
    gnet.backward()
    amax = gnet.argmax(1)
    dec_al = gnet.decode(1)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for b, (xx, oo, aa, ll, am) in enumerate(zip(split(x, Ts), split(out, Ts), split(aligned, Ts), split(labels, L), split(amax, Ts))):
        o_out, _ = onet.fwdbwd(xx, ll)
        assert np.abs(o_out - oo).max() < TOL
        o_al = oracle.ctc_align_labels(oo, ll)
        # with untrained (near-uniform) outputs the lattice values reach ~ -4.4*T ~ -6000 here; one Float ulp at that
        # magnitude is 4.9e-4, so the reference's own Float recursion is only defined to a few 1e-4 in the posteriors
        assert np.abs(o_al - aa).max() < 5e-4
        assert np.array_equal(oracle.argmax_rows(o_al), am)
        cs, locs = oracle.trivial_decode(o_al)
        assert np.array_equal(cs, dec_al[b][0]) and np.array_equal(locs, dec_al[b][1])
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < 5e-3 * max(1.0, np.abs(od).max())


def test_input_pipeline_matches_sequential_steps(ffi):
    # prefetch_batch(i+1) on the copy stream while step i runs, second input set; batches of changing geometry
    # (growing and shrinking) must give bit-identical decodes and parameters to one train_step per batch
    ni, nh, nc = 48, 100, 30
    batches = [synth.make_lines(B, T, ni, nc, seed=40 + k) for k, (B, T) in
               enumerate([(4, (30, 60)), (6, (50, 90)), (2, (10, 20)), (8, (80, 120)), (5, (40, 70))])]
    p0 = synth.reference_init(ni, nh, nc, seed=0.4)
    seq = ffi.Net(ni, nh, nc); seq.set_params(p0)
    ref_dec = []
    for x, T, labels, L in batches:
        dec, _, _ = seq.train_step(x, T, labels, L, 1e-3, 0.9)
        ref_dec.append(dec)
    pipe = ffi.Net(ni, nh, nc); pipe.set_params(p0)
    pipe.prefetch_batch(*batches[0])
    for k in range(len(batches)):
        pipe.step_prefetched(1e-3, 0.9)
        if k + 1 < len(batches):
            pipe.prefetch_batch(*batches[k + 1])
        dec = pipe.fetch_decoded(int(batches[k][1].max()) // 2 + 1)
        assert len(dec) == len(ref_dec[k])
        for (c0, l0), (c1, l1) in zip(ref_dec[k], dec):
            assert np.array_equal(c0, c1) and np.array_equal(l0, l1)
    assert np.array_equal(seq.get_params(), pipe.get_params())
    with pytest.raises(ffi.Error, match="no prefetched batch"):
        pipe.step_prefetched(1e-3, 0.9)


@pytest.mark.parametrize("mode", ["pair", "pair-barrier", "barrier"])
@pytest.mark.parametrize("nh,B,T", [(200, 5, (1, 70)), (400, 3, (20, 45)), (200, 2, (33, 33))])
def test_cluster_exchange_variants_parity(ffi, oracle, monkeypatch, nh, B, T, mode):
    # the cluster kernels exchange through mbarrier + st.async, one line per cluster by default (covered by CASES above) or
    # two lines software pipelined ("pair"); the cluster-barrier variants of both stay selectable for A/B runs and stay
    # tested: odd line counts and very different lengths inside a pair included
    monkeypatch.setenv("CLSTM_B200_LSTM", "simt")      # (the library itself prefers the tensor-core recurrence at these widths)
    if mode.startswith("pair"):
        monkeypatch.setenv("CLSTM_B200_CLUSTER_PAIR", "1")
    else:
        monkeypatch.setenv("CLSTM_B200_CLUSTER_PAIR", "0")
    if mode.endswith("barrier"):
        monkeypatch.setenv("CLSTM_B200_CLUSTER_MBAR", "0")
    ni, nc = 48, 83
    x, Ts, labels, L = synth.make_lines(B, T, ni, nc, seed=77)
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, "trained")
    out = gnet.forward(x, Ts)
    assert gnet.lstm_variant == "cluster"
    probe = np.random.default_rng(3).normal(0, 1, out.shape).astype(np.float32)
    gnet.clear_derivs()
    din = gnet.backward(probe)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for xx, oo, pp, dd in zip(split(x, Ts), split(out, Ts), split(probe, Ts), split(din, Ts)):
        assert np.abs(onet.forward(xx) - oo).max() < TOL
        assert np.abs(onet.backward(pp) - dd).max() < 2e-4 * max(1.0, np.abs(dd).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < 3e-4 * max(1.0, np.abs(od).max())


def test_ctc_limits_max_states_and_classes(ffi, oracle):
    # the device aligner's documented limits: 511 labels (S = 1023 lattice states, 32 per lane) and 512 classes;
    # alignment indices stay bit-exact there, one label more is refused (it must not silently truncate)
    ni, nh, nc = 16, 16, 512
    rng = np.random.default_rng(12)
    T = np.array([1100, 40], np.int32)
    x = rng.uniform(0, 1, (int(T.sum()), ni)).astype(np.float32)
    L = np.array([511, 0], np.int32)                                # second line: empty transcript (blank only)
    labels = rng.integers(1, nc, int(L.sum())).astype(np.int32)
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, "init")
    out = gnet.forward(x, T)
    aligned = gnet.ctc_align(labels, L)
    amax = gnet.argmax(1)
    for xx, oo, aa, ll, am in zip(split(x, T), split(out, T), split(aligned, T), split(labels, L), split(amax, T)):
        assert np.abs(onet.forward(xx) - oo).max() < TOL
        o_al = oracle.ctc_align_labels(oo, ll)
        assert np.abs(o_al - aa).max() < 5e-4                        # lattice values reach ~ -6.2*T here (see long-line test)
        assert np.array_equal(oracle.argmax_rows(o_al), am)
    assert np.array_equal(amax[T[0]:], np.zeros(T[1], np.int32))     # empty transcript aligns to blank everywhere
    with pytest.raises(ffi.Error, match="too long"):
        gnet.ctc_align(rng.integers(1, nc, 512).astype(np.int32), [512, 0])
    with pytest.raises(ffi.Error, match="not supported"):
        ffi.Net(ni, nh, 513)


def test_single_column_single_line(ffi, oracle):
    # B = 1, T = 1: every kernel's degenerate case (no recurrence step, lattice of one row)
    ni, nh, nc = 48, 100, 11
    onet, gnet = make_pair(ffi, oracle, ni, nh, nc, "trained")
    x = np.random.default_rng(2).uniform(0, 1, (1, ni)).astype(np.float32)
    dec, out, al = gnet.train_step(x, [1], np.array([3], np.int32), [1], 1e-3, 0.9, want_out=True, want_aligned=True)
    o_out, o_al = onet.fwdbwd(x, np.array([3], np.int32))
    onet.sgd_update(1e-3, 0.9)
    assert np.abs(o_out - out).max() < TOL and np.abs(o_al - al).max() < TOL
    assert np.abs(onet.get_params() - gnet.get_params()).max() < TOL
