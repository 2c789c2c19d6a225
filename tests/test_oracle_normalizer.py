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
