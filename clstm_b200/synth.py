"""Synthetic UW3-shaped text-line batches: H=48 binarised ink columns smoothed by a
3-tap box along t, transcripts of max(1, T//20) classes in [1, nclasses).  Host-side numpy only."""
import numpy as np


def make_lines(B, T, ninput=48, nclasses=83, seed=0):
    """T: int (fixed) or (lo, hi) for uniform variable lengths.  Returns x [sumT, ninput] f32, T [B] i32,
    labels [sumL] i32, L [B] i32."""
    rng = np.random.default_rng(seed)
    if isinstance(T, (tuple, list)):
        Ts = rng.integers(T[0], T[1] + 1, size=B).astype(np.int32)
    else:
        Ts = np.full(B, int(T), np.int32)
    xs, labs, Ls = [], [], []
    for b in range(B):
        t = np.arange(Ts[b])
        p = 0.2 + 0.15 * np.sin(2 * np.pi * t / (37.0 + 5 * (b % 7)) + b)      # slow sinusoid in [0.05, 0.35]
        ink = (rng.random((Ts[b], ninput)) < p[:, None]).astype(np.float32)
        pad = np.pad(ink, ((1, 1), (0, 0)), mode="edge")
        xs.append(((pad[:-2] + pad[1:-1] + pad[2:]) / 3.0).astype(np.float32))
        L = max(1, int(Ts[b]) // 20)
        labs.append(rng.integers(1, nclasses, size=L).astype(np.int32))
        Ls.append(L)
    return (np.ascontiguousarray(np.concatenate(xs, 0)), Ts, np.concatenate(labs).astype(np.int32),
            np.array(Ls, np.int32))


def trained_like(nparams, scale=0.3, seed=1):
    """'trained-like' weights (uniform +-scale) that exercise gate saturation."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, size=nparams).astype(np.float32)


def nparams(ninput, nhidden, nclasses):
    return 2 * 4 * nhidden * (1 + ninput + nhidden) + nclasses * (1 + 2 * nhidden)


def train_flops_per_column(ninput, nhidden, nclasses):
    """gate products fwd + 2x bwd, both directions, plus the softmax products."""
    return 48 * nhidden * (ninput + nhidden) + 12 * nclasses * nhidden


def reference_init(ninput, nhidden, nclasses, seed=0.1, scale=0.01):
    """The reference's deterministic weight init restated for the host side of the product:
    LCG of /root/reference/batches.cc:13-17 (env `seed`, default 0.1), mode "negbiased" with init_scale 0.01
    (clstm.cc:30-36) => uniform in [-0.02, 0.01]; draw order fwd WGI,WGF,WGO,WCI, rev WGI,WGF,WGO,WCI, W1
    (clstm.cc:588-591, clstm_prefab.cc:52-68), i outer / j inner within a matrix (batches.cc:37-38).
    Returns the flat walk_params-order vector (WCI,WGF,WGI,WGO per direction, col-major, bias column first)."""
    import math
    state = float(seed)
    s32 = np.float32(scale)
    a3 = float(np.float32(3) * s32)
    a2 = float(np.float32(2) * s32)

    def draw(rows, cols):
        nonlocal state
        m = np.empty((rows, cols), np.float64)
        for i in range(rows):
            for j in range(cols):
                state = 189843.9384938 * state + 0.328340981343
                state -= math.floor(state)
                m[i, j] = a3 * state - a2 + 0.0
        return m.astype(np.float32)

    nf = ninput + nhidden
    blocks = []
    for _ in range(2):
        wgi, wgf, wgo, wci = (draw(nhidden, nf + 1) for _ in range(4))
        blocks += [wci, wgf, wgi, wgo]            # std::map order of the parameter names
    blocks.append(draw(nclasses, 2 * nhidden + 1))
    return np.concatenate([b.T.ravel() for b in blocks]).astype(np.float32)   # col-major flatten


def make_raw_line(w, h, seed=0, nglyph=None):
    """Synthetic RAW text-line image [h][w] (row j, column i), ink = 1 on background 0, like a scanned line after the
    `raw = 1 - raw` of clstmocrtrain.cc:75: blobs ("glyphs") of varying height along a slowly drifting baseline plus
    a little noise.  Used by the normalizer parity tests and the CLI tests."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    jj = np.arange(h, dtype=np.float32)[:, None]
    ii = np.arange(w, dtype=np.float32)[None, :]
    base = 0.55 * h + 0.08 * h * np.sin(ii / max(w, 1) * 3.1 + rng.uniform(0, 3))
    xh = 0.22 * h
    n = nglyph if nglyph is not None else max(1, w // max(4, int(0.6 * h)))
    xs = np.sort(rng.uniform(0, w, n))
    for x0 in xs:
        gh = xh * rng.uniform(0.8, 2.0)
        gw = 0.25 * h * rng.uniform(0.5, 1.2)
        cy = base[0, int(min(max(x0, 0), w - 1))] - gh / 2 + rng.uniform(-0.05, 0.05) * h
        d = ((ii - x0) / gw) ** 2 + ((jj - cy) / (gh / 2)) ** 2
        img = np.maximum(img, np.clip(1.5 - d, 0, 1).astype(np.float32))
    img += rng.uniform(0, 0.02, img.shape).astype(np.float32)
    return np.clip(img, 0, 1).astype(np.float32)
