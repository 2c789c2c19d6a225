"""Normalises one batch of raw line images a few times (for ncu captures of normalize.cu; never a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clstm_b200
from clstm_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
imgs = [synth.make_raw_line(640, 60, seed=900 + b) for b in range(B)]
net = clstm_b200.Net(48, 16, 11)
for _ in range(3):
    T = net.normalize_batch(imgs, "center")
print("ok", int(T.sum()))
