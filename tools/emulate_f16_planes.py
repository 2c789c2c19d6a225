"""CPU emulation (numpy) of the fp16 hi/lo operand planes of gemm_x.cu / lstm_tcx.cu: x*s = hi + lo with hi = fp16(x*s),
lo = fp16(x*s - hi); product = hi*hi' + hi*lo' + lo*hi' (the lo*lo' term is dropped), accumulated in fp32 like the tensor core does.
Prints the error of a weight-derivative-sized reduction (K = 281 600 columns, deltas ~1e-3, activations in (-1, 1)) against the
float64 product, next to the error of a plain fp32 product of the same operands and of single-plane fp16 / TF32 operands.
usage: python tools/emulate_f16_planes.py [K]"""
import sys

import numpy as np


def split(x, s):
    xs = (x.astype(np.float32) * np.float32(s)).astype(np.float32)
    hi = np.clip(xs, -65504, 65504).astype(np.float16)
    lo = np.clip(xs - hi.astype(np.float32), -65504, 65504).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 281600
    rng = np.random.default_rng(1)
    M, N = 8, 8
    d = (rng.standard_normal((M, K)) * 1e-3 * (rng.random((M, K)) < 0.3)).astype(np.float32)      # sparse-ish deltas
    h = np.tanh(rng.standard_normal((N, K))).astype(np.float32)
    ref = d.astype(np.float64) @ h.astype(np.float64).T
    scale = np.abs(ref).max()
    sd, sh = 256.0, 16.0
    dh, dl = split(d, sd)
    hh, hl = split(h, sh)
    acc = np.zeros((M, N), np.float32)
    for k0 in range(0, K, 4096):                                   # fp32 accumulation in blocks (order differs from the device, same precision class)
        sl = slice(k0, k0 + 4096)
        acc += (dl[:, sl] @ hh[:, sl].T + dh[:, sl] @ hl[:, sl].T + dh[:, sl] @ hh[:, sl].T).astype(np.float32)
    planes = acc / np.float32(sd * sh)
    fp32 = np.zeros((M, N), np.float32)
    for k0 in range(0, K, 4096):
        sl = slice(k0, k0 + 4096)
        fp32 += (d[:, sl] @ h[:, sl].T).astype(np.float32)
    one_plane = (dh.astype(np.float64) @ hh.astype(np.float64).T) / (sd * sh)
    one_tf32 = tf32(d).astype(np.float64) @ tf32(h).astype(np.float64).T
    print("K = %d, max |C| = %.3e" % (K, scale))
    for name, v in (("fp16 hi/lo planes, 3 products (gemm_x)", planes), ("plain fp32 product", fp32), ("fp16 hi plane only", one_plane),
                    ("single TF32 (truncated)", one_tf32)):
        print("  %-42s max abs err %.3e  (%.2e of max |C|)" % (name, np.abs(v - ref).max(), np.abs(v - ref).max() / scale))


if __name__ == "__main__":
    main()
