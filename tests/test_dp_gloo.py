"""N>1 host logic on CPU (gloo, world_size 2): sharding + derivative sum + identical update must reproduce the
single-process result on the whole minibatch (share_deltas semantics, clstm.cc:731-744)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition():
    from clstm_b200 import dp
    T = np.array([50, 10, 40, 30, 20, 60, 5])
    parts = [dp.shard_indices(T, r, 3) for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(7))
    cols = [int(T[p].sum()) for p in parts]
    assert max(cols) - min(cols) <= 35                      # longest-first deal keeps column counts close


def test_two_rank_gloo_equals_single_process(oracle, tmp_path):
    from clstm_b200 import synth
    out = str(tmp_path / "dp.npy")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533",
                           os.path.join(ROOT, "tests", "dp_worker.py"), "gloo", out, "12"], env=env, timeout=600)
    got = np.load(out)
    ni, nh, nc, B, steps = 48, 12, 20, 6, 2
    x, T, labels, L = synth.make_lines(B, (25, 45), ni, nc, seed=17)
    net = oracle.BidiOracle(ni, nh, nc, seed=0.1)
    net.set_params(synth.trained_like(synth.nparams(ni, nh, nc), 0.3, seed=3))
    for _ in range(steps):
        net.train_lines(x, T, labels, L, 1e-3, 0.9, threads=1, reps=1)
    ref = net.get_params()
    assert np.abs(got - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())
