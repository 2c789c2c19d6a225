"""Generates the PNG fixtures of tests/test_host_extras.py (run once, here; outputs are committed).

Every file comes with <name>.expect.npy = the r+g+b sum per pixel ([h][w] int) that the reference's read_png
(extras.cc:305-551: libpng with STRIP_16 | STRIP_ALPHA | PACKING | EXPAND, then (r+g+b)/(3*255.0)) yields:
16-bit samples keep their high byte, grey below 8 bits is scaled to 0..255, palettes are expanded, alpha is dropped.
Writers: PIL for the common cases, a hand-rolled encoder for the cases PIL cannot write (Adam7 interlace, 2/4-bit
grey, explicit filter types, several IDAT chunks)."""
import os
import struct
import zlib

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)


def save_expect(name, rgbsum):
    np.save(os.path.join(HERE, name + ".expect.npy"), np.asarray(rgbsum, np.int32))


def chunk(t, body):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xffffffff)


def pack_rows(samples, depth):
    """samples [rows][n] ints -> list of packed byte rows (MSB first; 16 bit big endian)"""
    out = []
    for row in samples:
        if depth == 8:
            out.append(bytes(int(v) for v in row))
        elif depth == 16:
            out.append(b"".join(struct.pack(">H", int(v)) for v in row))
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in row)
            bits += "0" * (-len(bits) % 8)
            out.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    return out


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def filter_rows(rows, bpp, ftypes):
    prev = bytes(len(rows[0])) if rows else b""
    out = b""
    for y, cur in enumerate(rows):
        ft = ftypes[y % len(ftypes)]
        enc = bytearray()
        for x, v in enumerate(cur):
            a = cur[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            pred = [0, a, b, (a + b) >> 1, paeth(a, b, c)][ft]
            enc.append((v - pred) & 255)
        out += bytes([ft]) + bytes(enc)
        prev = cur
    return out


def write_raw_png(name, w, h, depth, ctype, samples, channels, plte=None, interlace=False, ftypes=(0, 1, 2, 3, 4), nidat=1):
    """samples: [h][w*channels] ints"""
    bpp = max(1, depth * channels // 8)
    body = b""
    if not interlace:
        body = filter_rows(pack_rows(samples, depth), bpp, ftypes)
    else:
        X0, Y0, DX, DY = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
        s = np.asarray(samples).reshape(h, w, channels)
        for p in range(7):
            sub = s[Y0[p]::DY[p], X0[p]::DX[p]]
            if sub.shape[0] == 0 or sub.shape[1] == 0:
                continue
            body += filter_rows(pack_rows(sub.reshape(sub.shape[0], -1), depth), bpp, ftypes)
    z = zlib.compress(body, 9)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    data += chunk(b"tEXt", b"Comment\x00fixture")          # an ancillary chunk the reader must skip
    if plte is not None:
        data += chunk(b"PLTE", bytes(int(v) for v in np.asarray(plte).ravel()))
    step = max(1, (len(z) + nidat - 1) // nidat)
    for i in range(0, len(z), step):
        data += chunk(b"IDAT", z[i:i + step])
    data += chunk(b"IEND", b"")
    open(os.path.join(HERE, name + ".png"), "wb").write(data)


w, h = 23, 9
g8 = rng.integers(0, 256, (h, w))
Image.fromarray(g8.astype(np.uint8), "L").save(os.path.join(HERE, "gray8_pil.png"))
save_expect("gray8_pil", 3 * g8)

rgb = rng.integers(0, 256, (h, w, 3))
Image.fromarray(rgb.astype(np.uint8), "RGB").save(os.path.join(HERE, "rgb8_pil.png"))
save_expect("rgb8_pil", rgb.sum(2))

rgba = rng.integers(0, 256, (h, w, 4))
Image.fromarray(rgba.astype(np.uint8), "RGBA").save(os.path.join(HERE, "rgba8_pil.png"))
save_expect("rgba8_pil", rgba[:, :, :3].sum(2))

b1 = rng.integers(0, 2, (h, w))
Image.fromarray((b1 * 255).astype(np.uint8), "L").convert("1").save(os.path.join(HERE, "gray1_pil.png"))
save_expect("gray1_pil", 3 * 255 * b1)

pal = rng.integers(0, 256, (256, 3))
idx = rng.integers(0, 256, (h, w))
im = Image.fromarray(idx.astype(np.uint8), "P")
im.putpalette([int(v) for v in pal.ravel()])
im.save(os.path.join(HERE, "pal8_pil.png"))
save_expect("pal8_pil", pal[idx].sum(1) if False else pal[idx].sum(2))

g16 = rng.integers(0, 65536, (h, w))
write_raw_png("gray16_filters", w, h, 16, 0, g16, 1)
save_expect("gray16_filters", 3 * (g16 >> 8))

for depth in (2, 4):
    g = rng.integers(0, 1 << depth, (h, w))
    write_raw_png("gray%d_filters" % depth, w, h, depth, 0, g, 1)
    save_expect("gray%d_filters" % depth, 3 * (g * 255 // ((1 << depth) - 1)))

pal4 = rng.integers(0, 256, (16, 3))
idx4 = rng.integers(0, 16, (h, w))
write_raw_png("pal4_filters", w, h, 4, 3, idx4, 1, plte=pal4)
save_expect("pal4_filters", pal4[idx4].sum(2))

ga = rng.integers(0, 256, (h, w, 2))
write_raw_png("graya8_filters", w, h, 8, 4, ga.reshape(h, -1), 2)
save_expect("graya8_filters", 3 * ga[:, :, 0])

rgb16 = rng.integers(0, 65536, (h, w, 3))
write_raw_png("rgb16_filters_3idat", w, h, 16, 2, rgb16.reshape(h, -1), 3, nidat=3)
save_expect("rgb16_filters_3idat", (rgb16 >> 8).sum(2))

wi, hi = 19, 13
rgbi = rng.integers(0, 256, (hi, wi, 3))
write_raw_png("rgb8_adam7", wi, hi, 8, 2, rgbi.reshape(hi, -1), 3, interlace=True)
save_expect("rgb8_adam7", rgbi.sum(2))
g1i = rng.integers(0, 2, (hi, wi))
write_raw_png("gray1_adam7", wi, hi, 1, 0, g1i, 1, interlace=True)
save_expect("gray1_adam7", 3 * 255 * g1i)
tiny = rng.integers(0, 256, (2, 3))
write_raw_png("gray8_adam7_tiny", 3, 2, 8, 0, tiny, 1, interlace=True)   # several empty passes
save_expect("gray8_adam7_tiny", 3 * tiny)

# cross-check every fixture against PIL's own decoder where PIL's semantics coincide with libpng's transforms
for f in sorted(os.listdir(HERE)):
    if not f.endswith(".png"):
        continue
    exp = np.load(os.path.join(HERE, f[:-4] + ".expect.npy"))
    im = Image.open(os.path.join(HERE, f))
    if im.mode in ("I;16", "I;16B", "I"):
        got = 3 * (np.asarray(im).astype(np.int64) >> 8)
    elif "16" in f:
        continue                                             # PIL converts 16-bit RGB its own way
    else:
        got = np.asarray(im.convert("RGB")).astype(np.int64).sum(2)
    assert np.array_equal(got, exp), f
    print("ok", f, im.mode, exp.shape)
