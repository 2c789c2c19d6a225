"""Writes tests/golden/ctc_kat.json: the two known-answer tests of the reference's CTC aligner.

The numbers are transcribed from /root/reference/test-ctc.cc (test1: lines 47-74, test2: lines
76-109).  In the reference each matrix literal is written classes-by-time and then transposed
(`transpose(outputs)`, test-ctc.cc:57,63,71), so here they are stored already transposed:
outputs[t][class], targets[s][class], expected[t][class].  Tolerance: max-abs < 1e-4
(test-ctc.cc:73,108).
"""
import json, os
import numpy as np

t1_out = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1]], float).T
t1_tgt = np.eye(3).T
t1_exp = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1]], float).T
t2_out = np.array([[1, .5, 0, 0, 0, 0], [0, .5, .5, 0, 0, 0], [0, 0, .5, .5, 0, 0],
                   [0, 0, 0, .5, .5, 0], [0, 0, 0, 0, .5, 1]], float).T
t2_tgt = np.eye(5).T
t2_exp = np.array([[1., 0.12029, 0., 0., 0., 0.], [0., 0.87971, 0.40013, 0., 0., 0.],
                   [0., 0., 0.59987, 0.59987, 0., 0.], [0., 0., 0., 0.40013, 0.87971, 0.],
                   [0., 0., 0., 0., 0.12029, 1.]], float).T
kat = {"source": "tmbdev/clstm test-ctc.cc:47-109", "tolerance": 1e-4,
       "cases": [
           {"name": "test1", "outputs": t1_out.tolist(), "targets": t1_tgt.tolist(), "expected": t1_exp.tolist()},
           {"name": "test2", "outputs": t2_out.tolist(), "targets": t2_tgt.tolist(), "expected": t2_exp.tolist()}]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_kat.json"), "w") as f:
    json.dump(kat, f, indent=1)
