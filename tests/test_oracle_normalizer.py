"""CPU properties of the normalizer oracle (oracle/normalizer_oracle.cc restates extras.cc:57-285).
The reference has no tests or golden vectors for its normalizers ("parity unpinned" upstream), so the restatement is
pinned by the properties the algorithm must have."""
import numpy as np
import pytest

from clstm_b200 import synth


def test_gauss_mask_matches_definition(oracle):
    for sigma in (0.3, 1.0, 14.4, 30.0, 48.0):
        m, r = oracle.gauss_mask(sigma)
        assert r == 1 + int(3.0 * np.float32(sigma)) and len(m) == 2 * r + 1
        assert np.array_equal(m, m[::-1])
        assert abs(float(m.sum()) - 1.0) < 1e-5
        k = np.arange(-r, r + 1)
        ref = np.exp(-k * k / 2.0 / float(np.float32(sigma)) ** 2)
        assert np.allclose(m, ref / ref.sum(), rtol=1e-5, atol=1e-9)


def test_center_normalizer_centres_a_straight_band(oracle):
    h, w, c0, half = 64, 400, 37, 6
    img = np.zeros((h, w), np.float32)
    img[c0 - half:c0 + half + 1, :] = 1.0
    center, r, smooth = oracle.center_measure(img)
    assert np.allclose(center, c0, atol=0.51)                  # argmax of the smoothed band, ties -> last row
    mad = np.abs(np.arange(c0 - half, c0 + half + 1) - center.mean()).mean()
    assert r == int(4.0 * np.float32(mad) + 1)
    x = oracle.center_line(img, 48)
    assert x.shape[1] == 48 and x.shape[0] == max(int(w / np.float32(2.0 * r / 48)), 1)
    rows = x.mean(axis=0)                                      # the band sits around row target_height/2
    assert abs(float((rows * np.arange(48)).sum() / rows.sum()) - 24.0) < 1.0


def test_center_normalizer_follows_a_sloped_baseline(oracle):
    img = synth.make_raw_line(700, 60, seed=3)
    center, r, _ = oracle.center_measure(img)
    assert center.shape == (700,) and 0 < center.min() and center.max() < 60
    assert r == int(r) and r >= 1
    x = oracle.center_line(img)
    assert x.shape[1] == 48 and np.isfinite(x).all() and 0 <= x.min() and x.max() <= 1.0 + 1e-6


def test_normalised_width_scales_with_the_line(oracle):
    a = synth.make_raw_line(500, 40, seed=5)
    b = np.kron(a, np.ones((2, 2), np.float32))                # the same line at twice the resolution
    xa, xb = oracle.center_line(a), oracle.center_line(b)
    assert abs(xa.shape[0] - xb.shape[0]) <= 0.15 * xa.shape[0]


def test_mean_normalizer_measures_centroid(oracle):
    h, w = 50, 120
    img = np.zeros((h, w), np.float32)
    img[20:31, 10:100] = 0.5
    x, ym, yd = oracle.mean_line(img, 48)
    assert abs(ym - 25.0) < 1e-9 and abs(yd - np.abs(np.arange(20, 31) - 25).mean()) < 1e-9
    assert x.shape == (int(w / np.float32(np.float32(2 * yd) / 48)), 48)


def _np_gauss1d(v, sigma):
    """independent numpy restatement of gauss1d (extras.cc:57-87), written from the comment "FIR filter, mask to 3 sigma,
    edges clamped", without looking at the oracle's loop structure: float32 products, float64 accumulation in tap order"""
    sigma = np.float32(sigma)
    r = 1 + int(3.0 * sigma)
    k = np.arange(-r, r + 1)
    mask = np.exp(-(k * k) / 2.0 / float(sigma) / float(sigma)).astype(np.float32)
    total = np.float32(0)
    for m in mask:
        total = np.float32(total + m)
    mask = (mask / total).astype(np.float32)
    n = len(v)
    idx = np.clip(np.arange(n)[:, None] + k[None, :], 0, n - 1)
    prod = (v[idx].astype(np.float32) * mask[None, :]).astype(np.float32)          # float products
    acc = np.zeros(n, np.float64)
    for j in range(prod.shape[1]):                                                 # sequential double accumulation
        acc += prod[:, j].astype(np.float64)
    return acc.astype(np.float32)


def test_center_measure_matches_independent_numpy_restatement(oracle):
    img = synth.make_raw_line(90, 33, seed=21)
    h, w = img.shape
    # gauss2d: along y with sigma h*0.5 inside every column, then along x with sigma h*smooth2d inside every row
    sm = img.copy()
    for i in range(w):
        sm[:, i] = _np_gauss1d(sm[:, i], np.float32(h * 0.5))
    for j in range(h):
        sm[j, :] = _np_gauss1d(sm[j, :], np.float32(h * np.float32(1.0)))
    # add_smear: v = 0.9 v + line, smooth += min(1, v) * 1e-3 (double arithmetic, stored float)
    for j in range(h):
        v = 0.0
        for i in range(w):
            v = v * 0.9 + float(img[j, i])
            sm[j, i] = np.float32(float(sm[j, i]) + min(1.0, v) * 1e-3)
    # argmax per column, ties -> last row; then gauss1d with sigma h*smooth1d
    a = np.array([max(range(h), key=lambda j: (sm[j, i], j)) for i in range(w)], np.float32)
    center = _np_gauss1d(a, np.float32(h * np.float32(0.3)))
    oc, orr, osm = oracle.center_measure(img)
    assert np.array_equal(osm, sm)
    assert np.array_equal(oc, center)
    s1 = np.float32(0); sy = np.float32(0)
    for i in range(w):
        for j in range(h):
            s1 = np.float32(s1 + img[j, i])
            sy = np.float32(sy + np.float32(img[j, i] * np.float32(abs(np.float32(j) - center[i]))))
    assert orr == float(int(np.float32(np.float32(4.0) * np.float32(sy / s1)) + np.float32(1)))


def test_center_normalize_matches_independent_numpy_restatement(oracle):
    # bilin (extras.cc:133-145) with its mixed arithmetic spelled out: l, m float; (1.0 - l), (1.0 - m) double;
    # m*s01 and m*s11 are float products; everything else double; result rounded to float
    img = synth.make_raw_line(120, 41, seed=31)
    h, w = img.shape
    center, r, _ = oracle.center_measure(img)
    th = 48
    scale = np.float32((2.0 * float(np.float32(r))) / th)
    tw = max(int(np.float32(w) / scale), 1)
    out = np.zeros((th, tw), np.float32)
    cl = lambda v, n: min(max(v, 0), n - 1)
    for i in range(tw):
        x = np.float32(scale * np.float32(i))
        for j in range(th):
            y = np.float32(np.float32(scale * np.float32(j - th // 2)) + center[int(x)])
            ii, jj = int(np.floor(x)), int(np.floor(y))
            l, m = np.float32(x - np.float32(ii)), np.float32(y - np.float32(jj))
            s00, s01 = img[cl(jj, h), cl(ii, w)], img[cl(jj + 1, h), cl(ii, w)]
            s10, s11 = img[cl(jj, h), cl(ii + 1, w)], img[cl(jj + 1, h), cl(ii + 1, w)]
            t0 = (1.0 - float(m)) * float(s00) + float(np.float32(m * s01))
            t1 = (1.0 - float(m)) * float(s10) + float(np.float32(m * s11))
            out[j, i] = np.float32((1.0 - float(l)) * t0 + float(l) * t1)
    assert np.array_equal(oracle.center_normalize(img, center, r, th), out)
