"""Device normalizers (clstm_b200/csrc/normalize.cu) against the oracle restatement of extras.cc -- BIT-EXACT:
the normalised width, r and the centre-line indices are integer decisions, so every float/double operation is
replicated (same types, same order); pixels, centre lines and r must be identical, not merely close."""
import numpy as np
import pytest

from clstm_b200 import synth

pytestmark = pytest.mark.gpu



@pytest.fixture(scope="module")
def ffi():
    import clstm_b200
    clstm_b200.lib()
    return clstm_b200


SHAPES = [(300, 60), (1, 48), (17, 31), (1200, 97), (64, 48), (33, 200)]   # (w, h): ragged, tiny, tall


def lines(shapes, seed=0):
    return [synth.make_raw_line(w, h, seed=seed + k) for k, (w, h) in enumerate(shapes)]


def split(a, T):
    o = np.concatenate([[0], np.cumsum(T)])
    return [a[o[i]:o[i + 1]] for i in range(len(T))]


def test_center_normalizer_bit_exact(ffi, oracle):
    imgs = lines(SHAPES)
    net = ffi.Net(48, 16, 10)
    T = net.normalize_batch(imgs, "center")
    x = net.get_inputs()
    center, r = net.normalizer_state()
    cs = split(center, [im.shape[1] for im in imgs])
    for b, (im, xb) in enumerate(zip(imgs, split(x, T))):
        oc, orr, _ = oracle.center_measure(im)
        assert np.array_equal(oc, cs[b]), (b, np.abs(oc - cs[b]).max())
        assert orr == r[b]
        ox = oracle.center_line(im, 48)
        assert ox.shape[0] == T[b]
        assert np.array_equal(ox, xb), (b, np.abs(ox - xb).max())


def test_center_normalizer_custom_params_and_height(ffi, oracle):
    imgs = lines([(240, 52), (90, 77)], seed=20)
    net = ffi.Net(32, 8, 5)                                   # target_height 32
    T = net.normalize_batch(imgs, "center", params=[3.0, 0.7, 0.5, 1.0])
    for im, xb in zip(imgs, split(net.get_inputs(), T)):
        assert np.array_equal(oracle.center_line(im, 32, 3.0, 0.7, 0.5), xb)


def test_mean_and_none_normalizers_bit_exact(ffi, oracle):
    imgs = lines([(150, 40), (400, 66), (9, 48)], seed=40)
    net = ffi.Net(48, 8, 5)
    T = net.normalize_batch(imgs, "mean")
    for im, xb, t in zip(imgs, split(net.get_inputs(), T), T):
        ox, _, _ = oracle.mean_line(im, 48)
        assert ox.shape[0] == t and np.array_equal(ox, xb)
    imgs48 = lines([(70, 48), (5, 48)], seed=50)
    T = net.normalize_batch(imgs48, "none")
    assert list(T) == [70, 5]
    for im, xb in zip(imgs48, split(net.get_inputs(), T)):
        assert np.array_equal(np.ascontiguousarray(im.T), xb)  # x[t][j] = raw(t, j)
    with pytest.raises(ffi.Error, match="NoNormalizer"):
        net.normalize_batch(imgs, "none")                      # extras.cc:149 asserts the height


def test_normalized_batch_feeds_the_network(ffi, oracle):
    # normalize_batch + forward_resident == forward on the oracle-normalised lines (same resident input, same kernels)
    imgs = lines([(300, 60), (180, 44), (520, 71)], seed=70)
    ni, nh, nc = 48, 100, 20
    net = ffi.Net(ni, nh, nc)
    net.set_params(synth.reference_init(ni, nh, nc, seed=0.3))
    labels = np.array([1, 2, 3, 4, 5, 6], np.int32)
    T = net.normalize_batch(imgs, "center", labels=labels, L=[2, 2, 2])
    out = net.forward_resident()
    xo = np.concatenate([oracle.center_line(im) for im in imgs])
    out2 = net.forward(xo, T)
    assert np.array_equal(out, out2)
    net.upload_batch(xo, T, labels, [2, 2, 2])
    net.step_resident(1e-4, 0.9)
    p_ref = net.get_params()
    net.set_params(synth.reference_init(ni, nh, nc, seed=0.3))
    net.clear_derivs()
    net.normalize_batch(imgs, "center", labels=labels, L=[2, 2, 2])
    net.step_resident(1e-4, 0.9)
    assert np.array_equal(net.get_params(), p_ref)


def test_normalizer_errors(ffi):
    net = ffi.Net(48, 8, 5)
    with pytest.raises(ffi.Error, match="rows high"):
        net.normalize_batch([np.ones((2000, 3), np.float32)], "center")
    with pytest.raises(KeyError):
        net.normalize_batch([np.ones((48, 3), np.float32)], "fancy")


def test_raw_input_pipeline_matches_sequential(ffi):
    # prefetch_raw_batch normalises batch i+1 on the copy stream while step i runs; results must be bit-identical to
    # normalize_batch + step_resident per batch (batches of different geometry)
    ni, nh, nc = 48, 100, 12
    batches = []
    for k, shapes in enumerate([[(200, 60), (120, 44)], [(90, 52), (300, 66), (40, 30)], [(150, 48)]]):
        imgs = lines(shapes, seed=100 + 10 * k)
        L = [2] * len(imgs)
        labels = np.arange(1, 2 * len(imgs) + 1, dtype=np.int32) % (nc - 1) + 1
        batches.append((imgs, labels, L))
    p0 = synth.reference_init(ni, nh, nc, seed=0.3)
    seq = ffi.Net(ni, nh, nc); seq.set_params(p0)
    ref = []
    for imgs, labels, L in batches:
        T = seq.normalize_batch(imgs, "center", labels=labels, L=L)
        seq.step_resident(1e-3, 0.9)
        ref.append((T, seq.fetch_decoded(int(T.max()) // 2 + 1)))
    pipe = ffi.Net(ni, nh, nc); pipe.set_params(p0)
    Ts = [pipe.prefetch_raw_batch(batches[0][0], batches[0][1], batches[0][2])]
    for k in range(len(batches)):
        pipe.step_prefetched(1e-3, 0.9)
        if k + 1 < len(batches):
            Ts.append(pipe.prefetch_raw_batch(*batches[k + 1]))
        dec = pipe.fetch_decoded(int(Ts[k].max()) // 2 + 1)
        assert np.array_equal(Ts[k], ref[k][0])
        for (c0, l0), (c1, l1) in zip(ref[k][1], dec):
            assert np.array_equal(c0, c1) and np.array_equal(l0, l1)
    assert np.array_equal(seq.get_params(), pipe.get_params())
