"""Tiny training steps through every kernel family (regs / cluster-resident tensor-core / lock-step tensor-core / SIMT cluster / generic
LSTM, on-the-fly 3xTF32 and TMA-fed GEMM, CTC incl. the multi-warp lattice, decode, update, normalizers) for
compute-sanitizer runs:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clstm_b200
from clstm_b200 import synth
# (nhidden, recurrence forced through CLSTM_B200_LSTM, dense products forced through CLSTM_B200_GEMM, line lengths)
CASES = [(16, None, None, (9, 21)), (5, None, None, (9, 21)), (200, None, None, (9, 21)), (400, None, "x", (9, 21)), (200, "simt", None, (9, 21)),
         (64, "tc", None, (9, 21)), (16, None, "x", (700, 1400))]
for nh, rec, gemm, Tr in CASES:
    for k, v in (("CLSTM_B200_LSTM", rec), ("CLSTM_B200_GEMM", gemm)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    x, T, labels, L = synth.make_lines(3, Tr, 48, 11, seed=2)
    net = clstm_b200.Net(48, nh, 11)
    net.set_params(synth.trained_like(net.nparams, 0.3))
    for _ in range(2):
        dec, out, al = net.train_step(x, T, labels, L, 1e-3, 0.9, want_out=True, want_aligned=True)
    assert np.isfinite(out).all() and np.isfinite(al).all()
    net.forward(x, T); net.ctc_align(labels, L); net.backward(); net.decode(1); net.argmax(0)
    print("ok", nh, net.lstm_variant, rec, gemm, Tr)
os.environ.pop("CLSTM_B200_LSTM", None); os.environ.pop("CLSTM_B200_GEMM", None)
imgs = [synth.make_raw_line(w, h, seed=k) for k, (w, h) in enumerate([(40, 30), (17, 48), (70, 52)])]
net = clstm_b200.Net(48, 16, 11)
net.set_params(synth.trained_like(net.nparams, 0.3))
for kind in ("center", "mean"):
    T = net.normalize_batch(imgs, kind, labels=np.array([1, 2, 3], np.int32), L=[1, 1, 1])
    net.step_resident(1e-3, 0.9)
    net.synchronize()
    assert np.isfinite(net.get_inputs()).all()
    print("ok normalizer", kind, list(T))
