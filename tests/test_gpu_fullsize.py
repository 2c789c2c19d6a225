"""Parity at the FULL sizes of the BASELINE configs, against the CPU oracle:
  cfg2  nhidden=100, 32 lines x 500 columns : outputs, input deltas and every parameter derivative within 1e-4
  cfg3  nhidden=200, 128 ragged lines (200..2000 columns): per-line argmax(aligned) and trivial_decode(aligned) classes
        and locations bit-exact (the "CTC alignment indices"), outputs of sampled lines within 1e-4
Reference semantics: CLSTMOCR::train, /root/reference/clstmhl.h:201-217; ctc_align_targets / trivial_decode,
/root/reference/ctc.cc:57-112, 159-194."""
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


def check_indices(oracle, o_al, g_al, g_amax, g_dec, b, margin=1e-3):
    """argmax(aligned) per column and trivial_decode(aligned) of the device against the oracle's on the same outputs.
    Returns the number of near-tie decisions (top-2 / best-frame margin below `margin`) that came out differently;
    any other difference fails."""
    ties = 0
    o_amax = oracle.argmax_rows(o_al)
    for t in np.nonzero(o_amax != g_amax)[0]:
        top2 = np.sort(o_al[t])[-2:]
        assert top2[1] - top2[0] < margin, (b, int(t), top2)
        ties += 1
    cs, locs = oracle.trivial_decode(o_al)
    if ties == 0:
        assert np.array_equal(cs, g_dec[0]), b
        for k in np.nonzero(locs != g_dec[1])[0]:            # same run, another frame: only if the two frames tie
            c = cs[k]
            assert abs(o_al[locs[k], c] - o_al[g_dec[1][k], c]) < margin, (b, int(k))
            ties += 1
    return ties


def test_cfg2_full_size_values(ffi, oracle):
    ni, nh, nc, B, T = 48, 100, 83, 32, 500
    x, Ts, labels, L = synth.make_lines(B, T, ni, nc, seed=4)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    onet.set_params(synth.trained_like(onet.nparams, 0.2, seed=7))
    gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    out = gnet.forward(x, Ts)
    rng = np.random.default_rng(5)
    deltas = (rng.standard_normal(out.shape) * 0.1).astype(np.float32)     # identical deltas into both (clstmhl.h:211-212 injects them)
    gnet.clear_derivs()
    din = gnet.backward(deltas)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for xx, oo, dd, di in zip(split(x, Ts), split(out, Ts), split(deltas, Ts), split(din, Ts)):
        assert np.abs(onet.forward(xx) - oo).max() < TOL
        o_din = onet.backward(dd)
        assert np.abs(o_din - di).max() < TOL * max(1.0, np.abs(o_din).max())
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < TOL * max(1.0, np.abs(od).max())


@pytest.mark.parametrize("recurrence", ["auto", "tc", "simt"])
def test_cfg3_full_size_alignment_indices(ffi, oracle, monkeypatch, recurrence):
    # "auto": the variant the library picks for 128 lines of nhidden 200 (the cluster-resident tensor-core recurrence, lstm_tcx.cu);
    # "tc": the lock-step tensor-core recurrence forced through CLSTM_B200_LSTM=tc (read when the net is created);
    # "simt": the fp32 thread-block cluster kernels (CLSTM_B200_LSTM=simt)
    if recurrence != "auto":
        monkeypatch.setenv("CLSTM_B200_LSTM", recurrence)
    ni, nh, nc, B = 48, 200, 83, 128
    x, Ts, labels, L = synth.make_lines(B, (200, 2000), ni, nc, seed=1000)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    onet.set_params(synth.trained_like(onet.nparams, 0.2, seed=7))
    gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    out = gnet.forward(x, Ts)
    assert gnet.lstm_variant == {"auto": "tcx", "tc": "tc", "simt": "cluster"}[recurrence]
    aligned = gnet.ctc_align(labels, L)
    amax = gnet.argmax(1)
    dec = gnet.decode(1)
    outs, als, ams, labs = split(out, Ts), split(aligned, Ts), split(amax, Ts), split(labels, L)
    near_ties = 0
    for b in range(B):
        o_al = oracle.ctc_align_labels(outs[b], labs[b])      # the oracle aligner on the device's own outputs
        # lattice values reach several thousand for T ~ 2000, where one Float ulp is ~5e-4 (see test_long_lines_wide_net)
        assert np.abs(o_al - als[b]).max() < 5e-4, b
        near_ties += check_indices(oracle, o_al, als[b], ams[b], dec[b], b)
    # indices are bit-exact wherever the reference's own Float arithmetic defines them: the only admissible differences are
    # decisions between two posteriors that differ by less than the 5e-4 resolution above; they must be rare (4 of 138 566 columns on B200)
    assert near_ties <= max(4, int(1e-4 * Ts.sum())), near_ties
    # every line through the oracle net: outputs, then input deltas and parameter derivatives for identical injected deltas
    rng = np.random.default_rng(5)
    deltas = (rng.standard_normal(out.shape) * 0.1).astype(np.float32)
    gnet.clear_derivs()
    din = gnet.backward(deltas)
    gd = gnet.get_derivs()
    onet.clear_derivs()
    for b, (xx, dd, di) in enumerate(zip(split(x, Ts), split(deltas, Ts), split(din, Ts))):
        assert np.abs(onet.forward(xx) - outs[b]).max() < TOL, b
        o_din = onet.backward(dd)
        assert np.abs(o_din - di).max() < TOL * max(1.0, np.abs(o_din).max()), b
    od = onet.get_derivs()
    assert np.abs(od - gd).max() < TOL * max(1.0, np.abs(od).max())


def test_prefetch_of_a_larger_batch_keeps_pending_results(ffi):
    # pipelined order of clstmocrtrain (batch > 1): step(i), prefetch(i+1), fetch_decoded(i).  Batch i+1 has more lines and
    # more labels than anything before, which regrows the per-line scratch; the decoded result of step i must survive it.
    ni, nh, nc = 48, 32, 20
    small = synth.make_lines(2, (20, 30), ni, nc, seed=1)
    big = synth.make_lines(40, (60, 90), ni, nc, seed=2)
    p0 = synth.reference_init(ni, nh, nc, seed=0.3)
    seq = ffi.Net(ni, nh, nc); seq.set_params(p0)
    ref_small, _, _ = seq.train_step(*small, 1e-3, 0.9)
    ref_big, _, _ = seq.train_step(*big, 1e-3, 0.9)
    pipe = ffi.Net(ni, nh, nc); pipe.set_params(p0)
    pipe.prefetch_batch(*small)
    pipe.step_prefetched(1e-3, 0.9)
    pipe.prefetch_batch(*big)                                  # regrow happens here, before the fetch
    got_small = pipe.fetch_decoded(int(small[1].max()) // 2 + 1)
    pipe.step_prefetched(1e-3, 0.9)
    got_big = pipe.fetch_decoded(int(big[1].max()) // 2 + 1)
    for ref, got in ((ref_small, got_small), (ref_big, got_big)):
        assert len(ref) == len(got)
        for (c0, l0), (c1, l1) in zip(ref, got):
            assert np.array_equal(c0, c1) and np.array_equal(l0, l1)
    assert np.array_equal(seq.get_params(), pipe.get_params())


def test_cfg4_width_large_batch_uses_tensor_core_recurrence(ffi, oracle):
    # BASELINE config 4 width (nhidden 400) with 96 ragged lines: the library picks the cluster-resident tcgen05 recurrence by itself;
    # outputs of every line and the alignment indices against the oracle
    ni, nh, nc, B = 48, 400, 83, 96
    x, Ts, labels, L = synth.make_lines(B, (60, 240), ni, nc, seed=21)
    onet = oracle.BidiOracle(ni, nh, nc, seed=0.222)
    onet.set_params(synth.trained_like(onet.nparams, 0.15, seed=7))
    gnet = ffi.Net(ni, nh, nc)
    gnet.set_params(onet.get_params())
    out = gnet.forward(x, Ts)
    assert gnet.lstm_variant == "tcx"
    aligned = gnet.ctc_align(labels, L)
    amax = gnet.argmax(1)
    dec = gnet.decode(1)
    for b, (xx, oo, aa, am, ll) in enumerate(zip(split(x, Ts), split(out, Ts), split(aligned, Ts), split(amax, Ts), split(labels, L))):
        assert np.abs(onet.forward(xx) - oo).max() < TOL, b
        o_al = oracle.ctc_align_labels(oo, ll)
        assert np.abs(o_al - aa).max() < 1e-4, b
        assert np.array_equal(oracle.argmax_rows(o_al), am), b
        cs, locs = oracle.trivial_decode(o_al)
        assert np.array_equal(cs, dec[b][0]) and np.array_equal(locs, dec[b][1]), b
