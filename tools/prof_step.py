"""Runs a few resident training steps of a bench workload (for ncu captures; never a bench number).
usage: python tools/prof_step.py nhidden lines T steps [Tmax]      (Tmax > T: ragged lines, the cfg3 / cfg4 shapes)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clstm_b200  # noqa: E402
from clstm_b200 import synth  # noqa: E402

nh = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
Tmax = int(sys.argv[5]) if len(sys.argv) > 5 else 0
x, Ts, labels, L = synth.make_lines(B, (T, Tmax) if Tmax > T else T, 48, 83, seed=1000)
net = clstm_b200.Net(48, nh, 83)
net.set_params(synth.reference_init(48, nh, 83, seed=0.222))
net.upload_batch(x, Ts, labels, L)
for _ in range(steps):
    net.step_resident(1e-4, 0.9, 100.0)
net.synchronize()
print("ok", net.lstm_variant)
