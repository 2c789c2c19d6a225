"""ctypes binding of clstm_b200/lib/libclstm_b200.so (C ABI: include/clstm_b200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclstm_b200.so")
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_longlong)


class Error(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [("ninput", C.c_int), ("nhidden", C.c_int), ("nclasses", C.c_int), ("device", C.c_int)]


class CfgEx(C.Structure):
    _fields_ = [("ninput", C.c_int), ("noutput", C.c_int), ("device", C.c_int), ("nblocks", C.c_int),
                ("nhidden", C.c_int * 2), ("direction", C.c_int * 2), ("cell", C.c_int), ("output", C.c_int)]


CELLS = {"NPLSTM": 0, "LINNPLSTM": 1, "RELUTANHNPLSTM": 2, "RELUNPLSTM": 3, "RELU2NPLSTM": 4}
OUTPUTS = {"SoftmaxLayer": 0, "SigmoidLayer": 1, "LinearLayer": 2, "TanhLayer": 3, "ReluLayer": 4, None: -1}
# prefab name (clstm_prefab.cc:152-160) -> (nblocks, directions)
PREFABS = {"lstm1": (1, (0, 0)), "revlstm1": (1, (1, 0)), "bidi": (1, (2, 0)), "bidi0": (1, (2, 0)), "bidi2": (2, (2, 2))}


# every symbol include/clstm_b200.h declares: (restype, argtypes)
EXPORTS = {
    "clstm_b200_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(C.c_void_p)]),
    "clstm_b200_create_ex": (C.c_int, [C.POINTER(CfgEx), C.POINTER(C.c_void_p)]),
    "clstm_b200_destroy": (None, [C.c_void_p]),
    "clstm_b200_nparams": (C.c_size_t, [C.c_void_p]),
    "clstm_b200_set_params": (C.c_int, [C.c_void_p, f32p, C.c_size_t]),
    "clstm_b200_get_params": (C.c_int, [C.c_void_p, f32p, C.c_size_t]),
    "clstm_b200_get_derivs": (C.c_int, [C.c_void_p, f32p, C.c_size_t]),
    "clstm_b200_set_derivs": (C.c_int, [C.c_void_p, f32p, C.c_size_t]),
    "clstm_b200_clear_derivs": (C.c_int, [C.c_void_p]),
    "clstm_b200_forward": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int, f32p]),
    "clstm_b200_ctc_align": (C.c_int, [C.c_void_p, i32p, i32p, f32p]),
    "clstm_b200_ctc_align_states": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int, i32p, i32p, f32p]),
    "clstm_b200_backward": (C.c_int, [C.c_void_p, f32p, f32p]),
    "clstm_b200_decode": (C.c_int, [C.c_void_p, C.c_int, i32p, i32p, i32p, C.c_int]),
    "clstm_b200_argmax": (C.c_int, [C.c_void_p, C.c_int, i32p]),
    "clstm_b200_sgd_update": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "clstm_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "clstm_b200_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "clstm_b200_allreduce_derivs": (C.c_int, [C.c_void_p]),
    "clstm_b200_p2p_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clstm_b200_p2p_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "clstm_b200_train_step": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int, i32p, i32p, C.c_float, C.c_float, C.c_float,
                                        f32p, f32p, i32p, i32p, i32p, C.c_int]),
    "clstm_b200_upload_batch": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int, i32p, i32p]),
    "clstm_b200_step_resident": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "clstm_b200_fetch_decoded": (C.c_int, [C.c_void_p, C.c_int, i32p, i32p, i32p, C.c_int]),
    "clstm_b200_synchronize": (C.c_int, [C.c_void_p]),
    "clstm_b200_prefetch_batch": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int, i32p, i32p]),
    "clstm_b200_prefetch_raw_batch": (C.c_int, [C.c_void_p, f32p, i32p, i32p, C.c_int, C.c_int, f32p, i32p, i32p, i32p]),
    "clstm_b200_step_prefetched": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "clstm_b200_normalize_batch": (C.c_int, [C.c_void_p, f32p, i32p, i32p, C.c_int, C.c_int, f32p, i32p, i32p, i32p]),
    "clstm_b200_normalizer_state": (C.c_int, [C.c_void_p, f32p, f32p]),
    "clstm_b200_get_inputs": (C.c_int, [C.c_void_p, f32p]),
    "clstm_b200_forward_resident": (C.c_int, [C.c_void_p, f32p]),
    "clstm_b200_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "clstm_b200_num_phases": (C.c_int, []),
    "clstm_b200_phase_name": (C.c_char_p, [C.c_int]),
    "clstm_b200_phase_stats": (C.c_int, [C.c_void_p, f32p, i64p, C.c_int]),
    "clstm_b200_stream": (C.c_void_p, [C.c_void_p]),
    "clstm_b200_lstm_variant": (C.c_char_p, [C.c_void_p]),
    "clstm_b200_selftest_gemm": (C.c_int, [C.c_void_p, f32p, C.c_int]),
    "clstm_b200_peer_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "clstm_b200_selftest_lstm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_float, f32p]),
    "clstm_b200_selftest_lstm_x": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_float, f32p]),
    "clstm_b200_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "clstm_b200_free_pinned": (None, [C.c_void_p]),
    "clstm_b200_last_error": (C.c_char_p, []),
    "clstm_b200_version": (C.c_char_p, []),
}


def build():
    """Compile libclstm_b200.so for sm_100a with nvcc (no GPU needed)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise Error("libclstm_b200.so is not built (%s); run __graft_entry__.build(). "
                        "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise Error(lib().clstm_b200_last_error().decode())


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(i32p)


def selftest_lstm(nhidden, nlines, tmin, tmax, seed=1, wscale=None, device=0, cluster_resident=False):
    """Device A/B of the batched tensor-core recurrence against the fp32 SIMT kernels (clstm_b200_selftest_lstm).
    Returns dict(d_gates, d_cell, d_h, d_hprev, d_delta_rel, ms_tc_fwd, ms_tc_bwd, ms_simt_fwd, ms_simt_bwd)."""
    out = np.zeros(9, np.float32)
    if wscale is None:
        wscale = 0.5 / np.sqrt(nhidden)
    fn = lib().clstm_b200_selftest_lstm_x if cluster_resident else lib().clstm_b200_selftest_lstm
    _chk(fn(device, nhidden, nlines, tmin, tmax, seed, float(wscale), out.ctypes.data_as(f32p)))
    keys = ["d_gates", "d_cell", "d_h", "d_hprev", "d_delta_rel", "ms_tc_fwd", "ms_tc_bwd", "ms_simt_fwd", "ms_simt_bwd"]
    return dict(zip(keys, (float(v) for v in out)))


def pinned_array(shape, dtype):
    """numpy array backed by page-locked host memory (cudaHostAlloc via the C ABI)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().clstm_b200_alloc_pinned(max(n, 1))
    if not p:
        raise Error(lib().clstm_b200_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr  # intentionally never freed: lives for the process (bench buffers)


class Net:
    """Device-resident bidi net: Stacked{Parallel{NPLSTM, Reversed{NPLSTM}}, SoftmaxLayer}."""

    def __init__(self, ninput, nhidden, nclasses, device=0, prefab="bidi", nhidden2=0, cell="NPLSTM",
                 output="SoftmaxLayer"):
        L = lib()
        self.ni, self.nh, self.nc = ninput, nhidden, nclasses
        h = C.c_void_p()
        if prefab == "bidi" and cell == "NPLSTM" and output == "SoftmaxLayer":
            cfg = Cfg(ninput, nhidden, nclasses, device)
            _chk(L.clstm_b200_create(C.byref(cfg), C.byref(h)))
        else:   # the other prefabs / layer variants of clstm_prefab.cc, clstm.cc:382-389, 655-668
            nblocks, dirs = PREFABS[prefab]
            if prefab == "bidi0":
                output = None
                self.nc = 2 * nhidden
            ex = CfgEx(ninput, nclasses, device, nblocks, (C.c_int * 2)(nhidden, nhidden2), (C.c_int * 2)(*dirs),
                       CELLS[cell], OUTPUTS[output])
            _chk(L.clstm_b200_create_ex(C.byref(ex), C.byref(h)))
        self.h = h
        self.nparams = L.clstm_b200_nparams(h)
        self.N = 0
        self.B = 0

    def close(self):
        if getattr(self, "h", None):
            lib().clstm_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters (reference flat order)
    def set_params(self, flat):
        a, p = _f32(flat)
        _chk(lib().clstm_b200_set_params(self.h, p, a.size))

    def get_params(self):
        a = np.empty(self.nparams, np.float32)
        _chk(lib().clstm_b200_get_params(self.h, a.ctypes.data_as(f32p), a.size))
        return a

    def get_derivs(self):
        a = np.empty(self.nparams, np.float32)
        _chk(lib().clstm_b200_get_derivs(self.h, a.ctypes.data_as(f32p), a.size))
        return a

    def set_derivs(self, flat):
        a, p = _f32(flat)
        _chk(lib().clstm_b200_set_derivs(self.h, p, a.size))

    def clear_derivs(self):
        _chk(lib().clstm_b200_clear_derivs(self.h))

    # ---- passes
    def forward(self, x, T, want_out=True):
        xa, xp = _f32(x)
        Ta, Tp = _i32(T)
        self.N, self.B = int(Ta.sum()), Ta.size
        assert xa.size == self.N * self.ni
        out = np.empty((self.N, self.nc), np.float32) if want_out else None
        _chk(lib().clstm_b200_forward(self.h, xp, Tp, Ta.size, out.ctypes.data_as(f32p) if want_out else None))
        return out

    def ctc_align(self, labels, L, want=True):
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        al = np.empty((self.N, self.nc), np.float32) if want else None
        _chk(lib().clstm_b200_ctc_align(self.h, lp, Lp, al.ctypes.data_as(f32p) if want else None))
        return al

    def ctc_align_states(self, outputs, T, states, S):
        oa, op = _f32(outputs)
        Ta, Tp = _i32(T)
        sa, sp = _i32(states)
        Sa, Sp = _i32(S)
        al = np.empty_like(oa)
        _chk(lib().clstm_b200_ctc_align_states(self.h, op, Tp, Ta.size, sp, Sp, al.ctypes.data_as(f32p)))
        return al

    def backward(self, deltas=None, want_din=True):
        dp = None
        if deltas is not None:
            da, dp = _f32(deltas)
        din = np.empty((self.N, self.ni), np.float32) if want_din else None
        _chk(lib().clstm_b200_backward(self.h, dp, din.ctypes.data_as(f32p) if want_din else None))
        return din

    def decode(self, which=0, max_per_line=None):
        m = max_per_line or max(1, self.N)
        cls = np.zeros((self.B, m), np.int32)
        locs = np.zeros((self.B, m), np.int32)
        cnt = np.zeros(self.B, np.int32)
        _chk(lib().clstm_b200_decode(self.h, which, cls.ctypes.data_as(i32p), locs.ctypes.data_as(i32p),
                                     cnt.ctypes.data_as(i32p), m))
        return [(cls[b, :cnt[b]].copy(), locs[b, :cnt[b]].copy()) for b in range(self.B)]

    def argmax(self, which=0):
        idx = np.empty(self.N, np.int32)
        _chk(lib().clstm_b200_argmax(self.h, which, idx.ctypes.data_as(i32p)))
        return idx

    def sgd_update(self, lr, momentum, clip=100.0):
        _chk(lib().clstm_b200_sgd_update(self.h, lr, momentum, clip))

    def train_step(self, x, T, labels, L, lr, momentum, clip=100.0, want_out=False, want_aligned=False,
                   max_per_line=None):
        xa, xp = _f32(x)
        Ta, Tp = _i32(T)
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        self.N, self.B = int(Ta.sum()), Ta.size
        m = max_per_line or int(Ta.max()) // 2 + 1
        out = np.empty((self.N, self.nc), np.float32) if want_out else None
        al = np.empty((self.N, self.nc), np.float32) if want_aligned else None
        cls = np.zeros((self.B, m), np.int32)
        locs = np.zeros((self.B, m), np.int32)
        cnt = np.zeros(self.B, np.int32)
        _chk(lib().clstm_b200_train_step(self.h, xp, Tp, self.B, lp, Lp, lr, momentum, clip,
                                         out.ctypes.data_as(f32p) if want_out else None,
                                         al.ctypes.data_as(f32p) if want_aligned else None,
                                         cls.ctypes.data_as(i32p), locs.ctypes.data_as(i32p),
                                         cnt.ctypes.data_as(i32p), m))
        dec = [(cls[b, :cnt[b]].copy(), locs[b, :cnt[b]].copy()) for b in range(self.B)]
        return dec, out, al

    # ---- resident-batch measurement path
    def upload_batch(self, x, T, labels, L):
        xa, xp = _f32(x)
        Ta, Tp = _i32(T)
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        self.N, self.B = int(Ta.sum()), Ta.size
        _chk(lib().clstm_b200_upload_batch(self.h, xp, Tp, self.B, lp, Lp))

    # ---- raw line images -> normalised resident batch (extras.cc normalizers on the device)
    KINDS = {"none": 0, "mean": 1, "center": 2}

    def normalize_batch(self, images, kind="center", params=None, labels=None, L=None):
        """images: list of [h][w] float arrays (row j, column i; ink = 1).  Returns the normalised widths T."""
        W = np.array([im.shape[1] for im in images], np.int32)
        H = np.array([im.shape[0] for im in images], np.int32)
        raw = np.concatenate([np.ascontiguousarray(im, np.float32).ravel() for im in images])
        ra, rp = _f32(raw)
        T = np.zeros(len(images), np.int32)
        pp = None
        if params is not None:
            pa, pp = _f32(params)
        lp = Lp = None
        if L is not None:
            la, lp = _i32(labels)
            La, Lp = _i32(L)
        _chk(lib().clstm_b200_normalize_batch(self.h, rp, W.ctypes.data_as(i32p), H.ctypes.data_as(i32p), len(images),
                                              self.KINDS[kind], pp, lp, Lp, T.ctypes.data_as(i32p)))
        self.N, self.B = int(T.sum()), T.size
        self._rawW = W
        return T

    def normalizer_state(self):
        center = np.zeros(int(self._rawW.sum()), np.float32)
        r = np.zeros(self.B, np.float32)
        _chk(lib().clstm_b200_normalizer_state(self.h, center.ctypes.data_as(f32p), r.ctypes.data_as(f32p)))
        return center, r

    def get_inputs(self):
        x = np.zeros((self.N, self.ni), np.float32)
        _chk(lib().clstm_b200_get_inputs(self.h, x.ctypes.data_as(f32p)))
        return x

    def forward_resident(self):
        out = np.zeros((self.N, self.nc), np.float32)
        _chk(lib().clstm_b200_forward_resident(self.h, out.ctypes.data_as(f32p)))
        return out

    def prefetch_batch(self, x, T, labels, L):
        """stage the next batch on the copy stream; the arrays are kept alive until the matching step_prefetched"""
        xa, xp = _f32(x)
        Ta, Tp = _i32(T)
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        self._prefetch_keep = (xa, Ta, la, La)
        self._prefetch_geom = (int(Ta.sum()), Ta.size)
        _chk(lib().clstm_b200_prefetch_batch(self.h, xp, Tp, Ta.size, lp, Lp))

    def prefetch_raw_batch(self, images, labels, L, kind="center", params=None):
        """normalise the next batch of raw line images on the copy stream while the current step runs"""
        W = np.array([im.shape[1] for im in images], np.int32)
        H = np.array([im.shape[0] for im in images], np.int32)
        raw = np.concatenate([np.ascontiguousarray(im, np.float32).ravel() for im in images])
        ra, rp = _f32(raw)
        T = np.zeros(len(images), np.int32)
        pp = None
        if params is not None:
            pa, pp = _f32(params)
        la, lp = _i32(labels)
        La, Lp = _i32(L)
        _chk(lib().clstm_b200_prefetch_raw_batch(self.h, rp, W.ctypes.data_as(i32p), H.ctypes.data_as(i32p), len(images),
                                                 self.KINDS[kind], pp, lp, Lp, T.ctypes.data_as(i32p)))
        self._prefetch_geom = (int(T.sum()), T.size)
        return T

    def step_prefetched(self, lr, momentum, clip=100.0):
        _chk(lib().clstm_b200_step_prefetched(self.h, lr, momentum, clip))
        self.N, self.B = self._prefetch_geom

    def fetch_decoded(self, max_per_line, which=0):
        """decoded classes / locations of the last step (synchronises the handle's stream)"""
        m = int(max_per_line)
        if getattr(self, "_dec_bufs", None) is None or self._dec_bufs[0].shape != (self.B, m):
            self._dec_bufs = (pinned_array((self.B, m), np.int32), pinned_array((self.B, m), np.int32),
                              pinned_array((self.B,), np.int32))
        cls, locs, cnt = self._dec_bufs
        _chk(lib().clstm_b200_fetch_decoded(self.h, which, cls.ctypes.data_as(i32p), locs.ctypes.data_as(i32p),
                                            cnt.ctypes.data_as(i32p), m))
        return [(cls[b, :cnt[b]].copy(), locs[b, :cnt[b]].copy()) for b in range(self.B)]

    def step_resident(self, lr, momentum, clip=100.0):
        _chk(lib().clstm_b200_step_resident(self.h, lr, momentum, clip))

    def synchronize(self):
        _chk(lib().clstm_b200_synchronize(self.h))

    def profile(self, enable):
        _chk(lib().clstm_b200_profile(self.h, 1 if enable else 0))

    def phase_stats(self):
        L = lib()
        n = L.clstm_b200_num_phases()
        ms = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int64)
        _chk(L.clstm_b200_phase_stats(self.h, ms.ctypes.data_as(f32p), cnt.ctypes.data_as(i64p), n))
        return {L.clstm_b200_phase_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def peer_stats(self, reset=True):
        """{launches, wait_us, data_us, nvlink_bytes} of the fused NVLink all-reduce + update kernel since the last reset
        (the first call only arms the counters)."""
        out = (C.c_double * 4)()
        _chk(lib().clstm_b200_peer_stats(self.h, out, 1 if reset else 0))
        return {"launches": int(out[0]), "wait_us": float(out[1]), "data_us": float(out[2]), "nvlink_bytes_per_launch": float(out[3])}

    def selftest_gemm(self):
        err = np.zeros(16, np.float32)
        k = lib().clstm_b200_selftest_gemm(self.h, err.ctypes.data_as(f32p), 16)
        if k < 0:
            raise Error(lib().clstm_b200_last_error().decode())
        return err[:k]

    @property
    def stream(self):
        return lib().clstm_b200_stream(self.h)

    @property
    def lstm_variant(self):
        return lib().clstm_b200_lstm_variant(self.h).decode()

    # ---- data-parallel
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        _chk(lib().clstm_b200_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, id_bytes, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        _chk(lib().clstm_b200_comm_init(self.h, C.cast(buf, C.c_void_p), rank, world))

    def p2p_handle(self):
        buf = (C.c_char * 64)()
        _chk(lib().clstm_b200_p2p_handle(self.h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def p2p_connect(self, handles, rank, world):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _chk(lib().clstm_b200_p2p_connect(self.h, C.cast(buf, C.c_void_p), rank, world))

    def allreduce_derivs(self):
        _chk(lib().clstm_b200_allreduce_derivs(self.h))
