"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle/clstm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (clstm_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_randu.restype = C.c_double
        L.oracle_seed.argtypes = [C.c_double]
        L.oracle_rinit.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_float]
        L.oracle_bidi_create.restype = C.c_void_p
        L.oracle_bidi_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_prefab_create.restype = C.c_void_p
        L.oracle_prefab_create.argtypes = [C.c_char_p] + [C.c_int] * 6
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_nparams.restype = C.c_size_t
        L.oracle_nparams.argtypes = [C.c_void_p]
        for n in ("get_params", "get_derivs"):
            getattr(L, "oracle_" + n).argtypes = [C.c_void_p, f32p]
        for n in ("set_params", "set_derivs"):
            getattr(L, "oracle_" + n).argtypes = [C.c_void_p, f32p]
        L.oracle_clear_derivs.argtypes = [C.c_void_p]
        L.oracle_forward.argtypes = [C.c_void_p, f32p, C.c_int, f32p]
        L.oracle_backward.argtypes = [C.c_void_p, f32p, f32p]
        L.oracle_sgd_update.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.oracle_fwdbwd.argtypes = [C.c_void_p, f32p, C.c_int, i32p, C.c_int, f32p, f32p]
        L.oracle_ctc_align_dense.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_ctc_align_dense_f64.argtypes = [f64p, f64p, f64p, C.c_int, C.c_int, C.c_int]
        L.oracle_ctc_align_labels.argtypes = [f32p, f32p, C.c_int, C.c_int, i32p, C.c_int]
        L.oracle_trivial_decode.restype = C.c_int
        L.oracle_trivial_decode.argtypes = [f32p, C.c_int, C.c_int, i32p, i32p]
        L.oracle_argmax_rows.argtypes = [f32p, C.c_int, C.c_int, i32p]
        L.oracle64_create.restype = C.c_void_p
        L.oracle64_create.argtypes = [C.c_int] * 4
        L.oracle64_destroy.argtypes = [C.c_void_p]
        L.oracle64_nparams.restype = C.c_size_t
        L.oracle64_nparams.argtypes = [C.c_void_p]
        L.oracle64_noutput.restype = C.c_int
        L.oracle64_noutput.argtypes = [C.c_void_p]
        L.oracle64_get_params.argtypes = [C.c_void_p, f64p]
        L.oracle64_set_params.argtypes = [C.c_void_p, f64p]
        L.oracle64_forward.argtypes = [C.c_void_p, f64p, C.c_int, C.c_int, f64p]
        L.oracle64_backward.argtypes = [C.c_void_p, f64p, f64p, f64p]
        L.oracle_forward_batched.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, f32p]
        L.oracle_backward_batched.argtypes = [C.c_void_p, f32p, f32p]
        L.oracle_train_lines.restype = C.c_double
        L.oracle_train_lines.argtypes = [C.c_void_p, f32p, i32p, C.c_int, i32p, i32p, C.c_float,
                                         C.c_float, C.c_int, C.c_int]
        L.oracle_gauss_mask.restype = C.c_int
        L.oracle_gauss_mask.argtypes = [C.c_float, f32p, C.c_int]
        L.oracle_center_measure.restype = C.c_float
        L.oracle_center_measure.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p, f32p]
        L.oracle_center_width.restype = C.c_int
        L.oracle_center_width.argtypes = [C.c_int, C.c_float, C.c_int]
        L.oracle_center_normalize.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_float, C.c_int, f32p]
        L.oracle_mean_measure.argtypes = [f32p, C.c_int, C.c_int, f64p, f64p]
        L.oracle_mean_width.restype = C.c_int
        L.oracle_mean_width.argtypes = [C.c_int, C.c_double, C.c_float, C.c_float, C.c_int]
        L.oracle_mean_normalize.argtypes = [f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float, C.c_float,
                                            C.c_int, f32p]
        _LIB = L
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(f64p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(i32p)


class BidiOracle:
    """float32 bidi net: Stacked{Parallel{NPLSTM, Reversed{NPLSTM}}, SoftmaxLayer} (clstm_prefab.cc:52-68)."""

    def __init__(self, ni, nh, nc, seed=None):
        L = lib()
        if seed is not None:
            L.oracle_seed(float(seed))
        self.ni, self.nh, self.nc = ni, nh, nc
        self.h = L.oracle_bidi_create(ni, nh, nc)
        self.nparams = L.oracle_nparams(self.h)
        self.T = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def get_params(self):
        a = np.empty(self.nparams, np.float32)
        lib().oracle_get_params(self.h, a.ctypes.data_as(f32p))
        return a

    def set_params(self, flat):
        a, p = _f32(flat)
        assert a.size == self.nparams
        lib().oracle_set_params(self.h, p)

    def get_derivs(self):
        a = np.empty(self.nparams, np.float32)
        lib().oracle_get_derivs(self.h, a.ctypes.data_as(f32p))
        return a

    def set_derivs(self, flat):
        a, p = _f32(flat)
        lib().oracle_set_derivs(self.h, p)

    def clear_derivs(self):
        lib().oracle_clear_derivs(self.h)

    def forward(self, image):
        a, p = _f32(image)
        T = a.shape[0]
        self.T = T
        out = np.empty((T, self.nc), np.float32)
        lib().oracle_forward(self.h, p, T, out.ctypes.data_as(f32p))
        return out

    def backward(self, deltas):
        a, p = _f32(deltas)
        din = np.empty((self.T, self.ni), np.float32)
        lib().oracle_backward(self.h, p, din.ctypes.data_as(f32p))
        return din

    def fwdbwd(self, image, labels):
        a, p = _f32(image)
        lab, lp = _i32(labels)
        T = a.shape[0]
        self.T = T
        out = np.empty((T, self.nc), np.float32)
        al = np.empty((T, self.nc), np.float32)
        lib().oracle_fwdbwd(self.h, p, T, lp, lab.size, out.ctypes.data_as(f32p), al.ctypes.data_as(f32p))
        return out, al

    def sgd_update(self, lr, momentum, gc=100.0):
        lib().oracle_sgd_update(self.h, lr, momentum, gc)

    def train_lines(self, x, T, labels, L, lr, momentum, threads=1, reps=1):
        xa, xp = _f32(x)
        Ta, Tp = _i32(T)
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        return lib().oracle_train_lines(self.h, xp, Tp, Ta.size, lp, Lp, lr, momentum, threads, reps)


CELLS = {"NPLSTM": 0, "LINNPLSTM": 1, "RELUTANHNPLSTM": 2, "RELUNPLSTM": 3, "RELU2NPLSTM": 4}
OUTPUTS = {"SoftmaxLayer": 0, "SigmoidLayer": 1, "LinearLayer": 2, "TanhLayer": 3, "ReluLayer": 4, None: -1}


class PrefabOracle(BidiOracle):
    """the 1-D prefabs of clstm_prefab.cc:22-129 (lstm1, revlstm1, bidi, bidi0, bidi2) with the LSTM cell variants of
    clstm.cc:655-668 and the Full<F> / Softmax output layers of clstm.cc:382-419"""

    def __init__(self, prefab, ni, nh, nc, nh2=0, cell="NPLSTM", output="SoftmaxLayer", seed=None):
        L = lib()
        if seed is not None:
            L.oracle_seed(float(seed))
        if prefab == "bidi0":
            output, nc = None, 2 * nh
        self.ni, self.nh, self.nc = ni, nh, nc
        self.h = L.oracle_prefab_create(prefab.encode(), ni, nh, nh2, nc, CELLS[cell], OUTPUTS[output])
        assert self.h, prefab
        self.nparams = L.oracle_nparams(self.h)
        self.T = 0


def ctc_align_dense(outputs, targets, double=False):
    if double:
        o, op = _f64(outputs)
        t, tp = _f64(targets)
        res = np.empty_like(o)
        lib().oracle_ctc_align_dense_f64(res.ctypes.data_as(f64p), op, tp, o.shape[0], t.shape[0], o.shape[1])
        return res
    o, op = _f32(outputs)
    t, tp = _f32(targets)
    res = np.empty_like(o)
    lib().oracle_ctc_align_dense(res.ctypes.data_as(f32p), op, tp, o.shape[0], t.shape[0], o.shape[1])
    return res


def ctc_align_labels(outputs, labels):
    o, op = _f32(outputs)
    lab, lp = _i32(labels)
    res = np.empty_like(o)
    lib().oracle_ctc_align_labels(res.ctypes.data_as(f32p), op, o.shape[0], o.shape[1], lp, lab.size)
    return res


def trivial_decode(outputs):
    o, op = _f32(outputs)
    T, nc = o.shape
    cs = np.empty(max(T, 1), np.int32)
    locs = np.empty(max(T, 1), np.int32)
    n = lib().oracle_trivial_decode(op, T, nc, cs.ctypes.data_as(i32p), locs.ctypes.data_as(i32p))
    return cs[:n].copy(), locs[:n].copy()


def argmax_rows(m):
    o, op = _f32(m)
    idx = np.empty(o.shape[0], np.int32)
    lib().oracle_argmax_rows(op, o.shape[0], o.shape[1], idx.ctypes.data_as(i32p))
    return idx


class Net64:
    """double nets for the gradient-check recipe of test-deriv.cc. kind 0 NPLSTM, 1 Reversed{NPLSTM}, 2 bidi."""

    def __init__(self, kind, ni, nh, no, seed=0.1):
        L = lib()
        L.oracle_seed(float(seed))
        self.ni = ni
        self.h = L.oracle64_create(kind, ni, nh, no)
        self.no = L.oracle64_noutput(self.h)
        self.nparams = L.oracle64_nparams(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle64_destroy(self.h)
            self.h = None

    def get_params(self):
        a = np.empty(self.nparams, np.float64)
        lib().oracle64_get_params(self.h, a.ctypes.data_as(f64p))
        return a

    def set_params(self, flat):
        a, p = _f64(flat)
        lib().oracle64_set_params(self.h, p)

    def forward(self, x):  # x [T][bs][ni]
        a, p = _f64(x)
        T, bs, _ = a.shape
        self.T, self.bs = T, bs
        out = np.empty((T, bs, self.no), np.float64)
        lib().oracle64_forward(self.h, p, T, bs, out.ctypes.data_as(f64p))
        return out

    def backward(self, dout):
        a, p = _f64(dout)
        din = np.empty((self.T, self.bs, self.ni), np.float64)
        dp = np.empty(self.nparams, np.float64)
        lib().oracle64_backward(self.h, p, din.ctypes.data_as(f64p), dp.ctypes.data_as(f64p))
        return din, dp


# ---- text-line normalizers (oracle/normalizer_oracle.cc).  Images are numpy [h][w] (row j, column i) = Tensor2 (i, j).
def gauss_mask(sigma):
    cap = 2 * (1 + int(3.0 * sigma)) + 8
    m = np.zeros(cap, np.float32)
    r = lib().oracle_gauss_mask(float(np.float32(sigma)), m.ctypes.data_as(f32p), cap)
    assert r >= 0
    return m[:2 * r + 1].copy(), r


def center_measure(img, range_=4.0, smooth2d=1.0, smooth1d=0.3):
    a, pa = _f32(img)
    h, w = a.shape
    center = np.zeros(w, np.float32)
    smooth = np.zeros((h, w), np.float32)
    r = lib().oracle_center_measure(pa, w, h, range_, smooth2d, smooth1d, center.ctypes.data_as(f32p),
                                    smooth.ctypes.data_as(f32p))
    return center, float(r), smooth


def center_normalize(img, center, r, target_height=48):
    a, pa = _f32(img)
    h, w = a.shape
    c, pc = _f32(center)
    tw = lib().oracle_center_width(w, r, target_height)
    out = np.zeros((target_height, tw), np.float32)
    lib().oracle_center_normalize(pa, w, h, pc, r, target_height, out.ctypes.data_as(f32p))
    return out


def center_line(img, target_height=48, range_=4.0, smooth2d=1.0, smooth1d=0.3):
    """measure + normalize; returns the network input [T][target_height] (x[t][j] = out(t, j))."""
    center, r, _ = center_measure(img, range_, smooth2d, smooth1d)
    return np.ascontiguousarray(center_normalize(img, center, r, target_height).T)


def mean_line(img, target_height=48, range_=1.0, vscale=1.0):
    a, pa = _f32(img)
    h, w = a.shape
    ym, yd = C.c_double(), C.c_double()
    lib().oracle_mean_measure(pa, w, h, C.byref(ym), C.byref(yd))
    tw = lib().oracle_mean_width(w, yd.value, vscale, range_, target_height)
    out = np.zeros((target_height, max(tw, 0)), np.float32)
    if tw > 0:
        lib().oracle_mean_normalize(pa, w, h, ym.value, yd.value, vscale, range_, target_height, out.ctypes.data_as(f32p))
    return np.ascontiguousarray(out.T), ym.value, yd.value
