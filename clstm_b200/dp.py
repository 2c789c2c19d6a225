"""Data-parallel host logic: shard a packed minibatch of text lines over ranks.

Lines are independent; only the parameter derivatives are exchanged (one sum all-reduce = the reference's unused
share_deltas, clstm.cc:731-744).  Sharding rule: sort lines by length (longest first) and deal them round-robin, so
the deal direction reversing every round (snake order), so every rank gets a similar number of columns."""
import numpy as np


def shard_indices(T, rank, world):
    """Line indices of `rank`: longest-first snake deal (0..w-1, w-1..0, 0..w-1, ...)."""
    order = np.argsort(-np.asarray(T), kind="stable")
    pos = np.arange(order.size)
    rnd, slot = pos // world, pos % world
    owner = np.where(rnd % 2 == 0, slot, world - 1 - slot)
    return np.sort(order[owner == rank])


def shard_batch(x, T, labels, L, rank, world):
    """Slice a packed batch (x [sumT, ni], T [B], labels [sumL], L [B]) down to the lines of `rank`."""
    T = np.asarray(T); L = np.asarray(L)
    xo = np.concatenate([[0], np.cumsum(T)]); lo = np.concatenate([[0], np.cumsum(L)])
    idx = shard_indices(T, rank, world)
    xs = [x[xo[i]:xo[i + 1]] for i in idx]
    ls = [labels[lo[i]:lo[i + 1]] for i in idx]
    ni = x.shape[1]
    return (np.concatenate(xs, 0) if xs else np.zeros((0, ni), np.float32), T[idx].astype(np.int32),
            np.concatenate(ls).astype(np.int32) if ls else np.zeros(0, np.int32), L[idx].astype(np.int32), idx)
