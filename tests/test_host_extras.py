"""Host pieces either side of the hot path (clstm_b200/host/clstm_extras.*):
the PNG reader against fixtures made with PIL / a hand-rolled encoder (tests/golden/png/make_png_fixtures.py),
and -- on a GPU -- the two CLI drop-ins clstmocrtrain / clstmocr end to end on synthetic line images."""
import glob
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "clstm_b200", "host")
BIN = os.path.join(ROOT, "clstm_b200", "bin")
PNGS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "png", "*.png")))


@pytest.fixture(scope="module")
def host_bin():
    import clstm_b200
    if not os.path.exists(clstm_b200.LIB_PATH):
        clstm_b200.build()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "test_host")


def test_fixture_set_is_complete():
    names = {os.path.basename(p) for p in PNGS}
    assert {"gray8_pil.png", "gray1_pil.png", "rgb8_pil.png", "rgba8_pil.png", "pal8_pil.png", "gray16_filters.png",
            "gray2_filters.png", "gray4_filters.png", "pal4_filters.png", "graya8_filters.png",
            "rgb16_filters_3idat.png", "rgb8_adam7.png", "gray1_adam7.png", "gray8_adam7_tiny.png"} <= names


@pytest.mark.parametrize("png", PNGS, ids=[os.path.basename(p) for p in PNGS])
def test_read_png_matches_libpng_semantics(host_bin, png):
    expect = np.load(png[:-4] + ".expect.npy")
    out = subprocess.check_output([host_bin, "png", png], text=True).split("\n")
    w, h = map(int, out[0].split())
    got = np.array([[int(v) for v in row.split()] for row in out[1:1 + h]])
    assert (h, w) == expect.shape
    assert np.array_equal(got, expect)


def test_read_png_rejects_garbage(host_bin, tmp_path):
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)
    r = subprocess.run([host_bin, "png", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "internal png error" in r.stderr
    good = open(PNGS[0], "rb").read()
    cut = tmp_path / "cut.png"
    cut.write_bytes(good[:len(good) // 2])
    r = subprocess.run([host_bin, "png", str(cut)], capture_output=True, text=True)
    assert r.returncode == 1 and "FATAL" in r.stderr


# ---------------------------------------------------------------------------------------------- CLIs on a GPU
GLYPH = {" ": 6, "a": 0, "b": 1, "c": 2, "d": 3, "e": 4}


def draw_line(text, rng):
    """white paper, black ink, 60 rows: character k lights a band of rows for 8 columns (test_host.cc's render_raw)"""
    T = 6 + 11 * len(text)
    ink = np.zeros((60, T + 10), np.float32)
    for k, ch in enumerate(text):
        band = GLYPH[ch]
        ink[7 + 6 * band:7 + 6 * band + 6, 5 + 4 + 11 * k:5 + 12 + 11 * k] = 1.0
    ink = np.clip(ink + rng.uniform(0, 0.04, ink.shape), 0, 1)
    return (255 * (1.0 - ink)).astype(np.uint8)


@pytest.mark.gpu
def test_cli_train_and_recognise(host_bin, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    texts = ["abc de", "ab", "cde a", "e d c", "bad cab", "a", "deed", "cab ace"]
    files = []
    for k, t in enumerate(texts):
        base = tmp_path / ("line%02d" % k)
        Image.fromarray(draw_line(t, rng), "L").save(str(base) + ".bin.png")
        (tmp_path / ("line%02d.gt.txt" % k)).write_text(t + "\n")
        files.append(str(base) + ".bin.png")
    (tmp_path / "train.txt").write_text("\n".join(files) + "\n")
    (tmp_path / "test.txt").write_text("\n".join(files[:4]) + "\n")
    env = dict(os.environ, nhidden="24", lrate="5e-3", ntrain="400", batch="8", report_every="50", test_every="100",
               save_every="200", save_name=str(tmp_path / "model"), params="0")
    r = subprocess.run([os.path.join(BIN, "clstmocrtrain"), str(tmp_path / "train.txt"), str(tmp_path / "test.txt")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "got 8 files, 4 tests" in r.stdout and "got 7 classes" in r.stdout       # blank + 6 characters
    assert "ERROR 100" in r.stdout and "saving" in r.stdout
    assert os.path.exists(tmp_path / "model-200.clstm") and os.path.exists(tmp_path / "model.clstm")
    last_err = float([ln for ln in r.stdout.split("\n") if ln.startswith("ERROR")][-1].split()[2])
    assert last_err < 0.2, r.stdout[-1500:]
    # batch=1 resumes from the checkpoint through the reference's per-line loop (start = trial + 1)
    env1 = dict(env, batch="1", load=str(tmp_path / "model-200.clstm"), ntrain="230", save_name=str(tmp_path / "resumed"),
                save_every="1000", test_every="1000")
    r1 = subprocess.run([os.path.join(BIN, "clstmocrtrain"), str(tmp_path / "train.txt")], env=env1, capture_output=True,
                        text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr[-2000:]
    assert "start 201" in r1.stdout and "ALN" in r1.stdout and os.path.exists(tmp_path / "resumed-229.clstm")
    # recognition CLI: list of files in, "<file>\t<text>" out, <base>.txt written, posteriors image on request
    env2 = dict(os.environ, load=str(tmp_path / "model-399.clstm") if os.path.exists(tmp_path / "model-399.clstm")
                else str(tmp_path / "model.clstm"), output="posteriors", params="0")
    r2 = subprocess.run([os.path.join(BIN, "clstmocr"), str(tmp_path / "test.txt")], env=env2, capture_output=True, text=True,
                        timeout=300)
    assert r2.returncode == 0, r2.stderr[-2000:]
    lines = [ln for ln in r2.stdout.split("\n") if "\t" in ln]
    assert len(lines) == 4
    ok = sum(ln.split("\t")[1] == t for ln, t in zip(lines, texts[:4]))
    assert ok >= 3, r2.stdout
    assert (tmp_path / "line00.bin.txt").read_text().strip() == lines[0].split("\t")[1]
    assert os.path.exists(tmp_path / "line00.bin.p.png")
    r3 = subprocess.run([os.path.join(BIN, "clstmocr"), str(tmp_path / "test.txt")], env=dict(os.environ, params="0"),
                        capture_output=True, text=True)
    assert r3.returncode == 1 and "must give load= parameter" in r3.stderr


# ---------------------------------------------------------------------------------------------- the reference's own fixture
OCR_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ocr")


def test_reference_png_fixture_decodes_like_pil(host_bin, tmp_path):
    """misc/textline.bin.png is an 819x88 RGBA file: the zlib-based decoder must produce what PIL produces."""
    from PIL import Image
    src = os.path.join(OCR_FIX, "textline.bin.png")
    im = Image.open(src)
    assert im.size == (819, 88)
    out = subprocess.check_output([host_bin, "png", src], text=True).split("\n")
    w, h = map(int, out[0].split())
    got = np.array([[int(v) for v in row.split()] for row in out[1:1 + h]])
    assert (w, h) == (819, 88)
    assert np.array_equal(got, np.array(im).astype(int)[:, :, :3].sum(2))       # read_png: mean of R, G, B (extras.cc:416-431)


@pytest.mark.gpu
def test_reference_ocr_fixture(tmp_path):
    """/root/reference/test-ocr.sh replayed with the drop-in CLIs: train 201 trials on the reference's text line, reload
    the trial-200 checkpoint, recognise the line and find 'performance analysis' in the output."""
    import shutil
    for f in ("textline.bin.png", "textline.gt.txt"):
        shutil.copy(os.path.join(OCR_FIX, f), tmp_path / f)
    (tmp_path / "_ocrtest.txt").write_text(str(tmp_path / "textline.bin.png") + "\n")
    env = dict(os.environ, ntrain="201", hidden="50", lrate="1e-2", save_name=str(tmp_path / "_ocrtest"), seed="0.222")
    r = subprocess.run([os.path.join(BIN, "clstmocrtrain"), str(tmp_path / "_ocrtest.txt")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(tmp_path / "_ocrtest-200.clstm"), r.stdout[-1500:]
    r2 = subprocess.run([os.path.join(BIN, "clstmocr"), str(tmp_path / "_ocrtest.txt")],
                        env=dict(os.environ, load=str(tmp_path / "_ocrtest-200.clstm")), capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "performance analysis" in r2.stdout, r2.stdout[-1000:]
