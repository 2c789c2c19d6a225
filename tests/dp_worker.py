"""Worker for the world_size-2 tests (launched by tests/test_dp_gloo.py and tests/test_gpu_multi.py).

mode "gloo": CPU. Each rank runs the reference step (oracle fwdbwd per line) on its shard, derivatives are summed with
a gloo all_reduce (share_deltas semantics), every rank applies the same update; rank 0 writes the parameters.
mode "nccl": GPU. Same with the CUDA path; the all-reduce is issued from inside libclstm_b200.so over NCCL.
mode "p2p":  GPU. The fused NVLink peer-memory kernel (all-reduce + clip + update in one launch)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clstm_b200 import synth, dp  # noqa: E402

mode, outfile = sys.argv[1], sys.argv[2]
ni, nh, nc, B, steps = 48, int(sys.argv[3]), 20, 6, 2
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
x, T, labels, L = synth.make_lines(B, (25, 45), ni, nc, seed=17)
xs, Ts, ls, Ls, idx = dp.shard_batch(x, T, labels, L, rank, world)
w0 = synth.trained_like(synth.nparams(ni, nh, nc), 0.3, seed=3)
lr, mom = 1e-3, 0.9

if mode == "gloo":
    from oracle import binding as ob
    dist.init_process_group("gloo")
    net = ob.BidiOracle(ni, nh, nc, seed=0.1)
    net.set_params(w0)
    mombuf = np.zeros_like(w0)                      # Params.d of the logical single net (derivative + momentum)
    xo = np.concatenate([[0], np.cumsum(Ts)]); lo = np.concatenate([[0], np.cumsum(Ls)])
    for _ in range(steps):
        net.clear_derivs()
        for i in range(len(Ts)):
            net.fwdbwd(xs[xo[i]:xo[i + 1]], ls[lo[i]:lo[i + 1]])
        g = torch.from_numpy(net.get_derivs().copy())
        dist.all_reduce(g, op=dist.ReduceOp.SUM)    # this step's derivatives only: momentum is not multiplied by world
        net.set_derivs(mombuf + g.numpy())
        net.sgd_update(lr, mom)
        mombuf = net.get_derivs()
    params = net.get_params()
else:
    import clstm_b200
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    net = clstm_b200.Net(ni, nh, nc, device=int(os.environ["LOCAL_RANK"]))
    net.set_params(w0)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(clstm_b200.Net.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    net.comm_init(idt.cpu().numpy().tobytes(), rank, world)
    if mode == "p2p":
        mine = torch.frombuffer(bytearray(net.p2p_handle()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        net.p2p_connect([h.cpu().numpy().tobytes() for h in allh], rank, world)
    for _ in range(steps):
        net.train_step(xs, Ts, ls, Ls, lr, mom)
    params = net.get_params()
all_p = [torch.zeros(params.size) for _ in range(world)]
if mode == "gloo":
    dist.all_gather(all_p, torch.from_numpy(params.copy()))
    assert all(torch.equal(all_p[0], p) for p in all_p), "ranks diverged"
if rank == 0:
    np.save(outfile, params)
dist.barrier()
dist.destroy_process_group()
