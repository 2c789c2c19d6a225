"""2-GPU data-parallel run (NCCL all-reduce issued from inside libclstm_b200.so) against the 1-GPU result on the whole
minibatch.  Skipped on boxes with fewer than 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_gpu_equals_one_gpu(tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import clstm_b200
    from clstm_b200 import synth
    out = str(tmp_path / "dp_gpu.npy")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29544",
                           os.path.join(ROOT, "tests", "dp_worker.py"), mode, out, "100"], timeout=600)
    got = np.load(out)
    ni, nh, nc, B, steps = 48, 100, 20, 6, 2
    x, T, labels, L = synth.make_lines(B, (25, 45), ni, nc, seed=17)
    net = clstm_b200.Net(ni, nh, nc)
    net.set_params(synth.trained_like(synth.nparams(ni, nh, nc), 0.3, seed=3))
    for _ in range(steps):
        net.train_step(x, T, labels, L, 1e-3, 0.9)
    ref = net.get_params()
    assert np.abs(got - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())
