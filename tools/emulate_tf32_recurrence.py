"""Feasibility check for the round-2 batched tcgen05 recurrence: does a 3xTF32 recurrent product
(operands split by truncation into TF32 hi + lo, hi*hi + hi*lo + lo*hi accumulated in fp32 -- what gemm_tc.cu issues) stay
inside the 1e-4 parity bar when it is applied T = 2000 times in a row, i.e. does the split error compound through the
recurrence?  CPU only (numpy), no product code involved.

usage: python tools/emulate_tf32_recurrence.py [nhidden] [T] [lines]"""
import sys
import numpy as np


def tf32_trunc(x):
    """fp32 -> value with the 13 low mantissa bits cleared (what split_tf32 in gemm_tc.cu produces for `hi`)"""
    return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)


def tf32_read(x):
    """what the tensor core reads of an fp32 word: the top 19 bits (TF32), i.e. truncation as well"""
    return tf32_trunc(x)


def matmul_3xtf32(A, B):
    """A [m,k] * B [k,n]: products of TF32 operands are exact in fp32, accumulation in fp32 (emulated in float64 then
    rounded per k-block of 8 like the MMA's K=8 granularity)"""
    Ah = tf32_trunc(A); Al = tf32_read((A - Ah).astype(np.float32))
    Bh = tf32_trunc(B); Bl = tf32_read((B - Bh).astype(np.float32))
    acc = np.zeros((A.shape[0], B.shape[1]), np.float32)
    for k0 in range(0, A.shape[1], 8):
        sl = slice(k0, k0 + 8)
        part = (Al[:, sl].astype(np.float64) @ Bh[sl].astype(np.float64) + Ah[:, sl].astype(np.float64) @ Bl[sl].astype(np.float64)
                + Ah[:, sl].astype(np.float64) @ Bh[sl].astype(np.float64))
        acc = (acc.astype(np.float64) + part).astype(np.float32)
    return acc


def sig(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


def run(no, T, nb, mode, R, XP):
    h = np.zeros((no, nb), np.float32); c = np.zeros((no, nb), np.float32)
    hs = []
    for t in range(T):
        if mode == "fp32":
            pre = (R.astype(np.float64) @ h.astype(np.float64)).astype(np.float32)       # fp32-grade reference (double accumulate)
        elif mode == "3xtf32":
            pre = matmul_3xtf32(R, h)
        else:                                                                            # plain TF32
            pre = (tf32_trunc(R).astype(np.float64) @ tf32_trunc(h).astype(np.float64)).astype(np.float32)
        pre = pre + XP[t]
        gi, gf, go = sig(pre[0::4]), sig(pre[1::4]), sig(pre[2::4])
        ci = np.tanh(pre[3::4]).astype(np.float32)
        c = (ci * gi + gf * c).astype(np.float32)
        h = (np.tanh(c) * go).astype(np.float32)
        hs.append(h.copy())
    return np.stack(hs)


if __name__ == "__main__":
    no = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    rng = np.random.default_rng(0)
    for scale, name in ((0.01, "reference init scale (negbiased 0.01)"), (0.3 / np.sqrt(no / 100.0), "trained-like")):
        R = rng.uniform(-2 * scale, scale, (4 * no, no)).astype(np.float32) if scale == 0.01 else rng.normal(0, scale, (4 * no, no)).astype(np.float32)
        XP = rng.normal(0, 1.0, (T, 4 * no, nb)).astype(np.float32)
        ref = run(no, T, nb, "fp32", R, XP)
        for mode in ("3xtf32", "tf32"):
            got = run(no, T, nb, mode, R, XP)
            err = np.abs(got - ref).reshape(T, -1).max(1)
            print("%-40s %-7s max|dh| over all t = %.3e   at t<10: %.3e   last 100 steps: %.3e" %
                  (name, mode, err.max(), err[:10].max(), err[-100:].max()))
