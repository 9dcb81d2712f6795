"""The reference's command line (src/bin/pbrt.rs:44-135) on the GPU path: option surface, PbrtOptions overrides in the
front end, image writers (write_image.rs).  The render itself runs in tests/hostcheck (CPU) and under -m gpu."""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_accepts_the_reference_options(P):
    from pbrt_r3_b200 import cli
    ap = cli.build_parser()
    o = ap.parse_args(["scene.pbrt", "--outfile", "a.exr", "--nthreads", "8", "--quick", "--quiet", "--pixelsamples", "4", "--cropwindow", "0,1,0,1",
                       "--logdir", "/tmp", "--logtostderr", "--minloglevel", "1", "--stats", "--profile", "--display-server", "localhost:14158"])
    assert o.pbrtfile == ["scene.pbrt"] and o.outfile == "a.exr" and o.quick and o.quiet and o.pixelsamples == 4 and o.stats
    o = ap.parse_args(["-i", "s.pbrt", "-o", "b.png", "-j", "2", "-s", "16", "--quick-full-resolution", "-k", "dir"])
    assert o.infile == "s.pbrt" and o.quick_full_resolution and o.sequential_display == "dir"
    assert cli.main([]) == -1 and cli.main(["/nonexistent.pbrt"]) == -1        # usage + exit code -1 (pbrt.rs:373-383)


def test_pbrt_options_quick_render(P):
    lib = P.capi.load()
    txt = P.scenes.cornell_box(res=64, spp=16)
    try:
        lib.pt_set_options(P.capi.pt_options(1, 0))
        d = P.HostScene.parse_string(txt).desc
        assert list(d.film.full_resolution) == [16, 16] and d.sampler.samples_per_pixel == 1      # film.rs:548-554, halton.rs:288-290
        assert P.HostScene.parse_string(txt, 8).desc.sampler.samples_per_pixel == 1              # --quick wins over --pixelsamples
        lib.pt_set_options(P.capi.pt_options(1, 1))
        d = P.HostScene.parse_string(txt).desc
        assert list(d.film.full_resolution) == [64, 64] and d.sampler.samples_per_pixel == 1
        got = P.capi.pt_options()
        lib.pt_get_options(got)
        assert (got.quick_render, got.quick_render_full_resolution) == (1, 1)
    finally:
        lib.pt_set_options(None)
    d = P.HostScene.parse_string(txt).desc
    assert list(d.film.full_resolution) == [64, 64] and d.sampler.samples_per_pixel == 16
    assert P.HostScene.parse_string(txt, 8).desc.sampler.samples_per_pixel == 8                  # pbrt.rs:234-238


def test_image_writers(P, tmp_path):
    from pbrt_r3_b200 import imageio
    rng = np.random.RandomState(0)
    rgb = rng.uniform(0, 2, (5, 7, 3)).astype(np.float32)
    rgb[0, 0] = [0.0, 0.001, np.nan]
    exr = str(tmp_path / "a.exr")
    imageio.write_image(exr, rgb)
    back = imageio.read_exr(exr)
    assert np.array_equal(back.view(np.uint32), rgb.view(np.uint32))                              # floats kept bit for bit
    # to_byte (write_image.rs:16-18): clamp(255 * gamma_correct(v), 0, 255) as u8
    b = imageio.to_bytes(np.array([[[0.0, 0.0031308, 0.5], [1.0, 2.0, np.nan]]], np.float32))
    assert b.tolist() == [[[0, 10, 187], [254, 255, 0]]]      # f32: 1.055 - 0.055 = 0.99999994 -> 254, as in the reference
    png = str(tmp_path / "a.png")
    imageio.write_image(png, rgb)
    raw = open(png, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (7, 5)
    idat = raw[raw.index(b"IDAT") + 4: raw.index(b"IEND") - 8]
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(5, 1 + 21)
    assert np.array_equal(rows[:, 1:].reshape(5, 7, 3), imageio.to_bytes(rgb))
    imageio.write_image(str(tmp_path / "a.pfm"), rgb)
    with pytest.raises(ValueError):
        imageio.write_image(str(tmp_path / "noext"), rgb)


@pytest.mark.gpu
def test_cli_renders_like_the_library(P, tmp_path):
    from pbrt_r3_b200 import imageio
    scene = tmp_path / "cornell.pbrt"
    scene.write_text(P.scenes.cornell_box(res=48, spp=4))
    out = str(tmp_path / "out.exr")
    r = subprocess.run([sys.executable, "-m", "pbrt_r3_b200", str(scene), "--outfile", out, "--pixelsamples", "2", "--stats", "--nthreads", "4"],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Camera rays traced" in r.stdout and str(50 * 50 * 2) in r.stdout
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=48, spp=4), 2)
    sc = P.Scene(hs, 0)
    rgb, _ = sc.render()
    sc.close()
    assert np.array_equal(imageio.read_exr(out), rgb)
    # Film "filename" is used when --outfile is absent; --quick = 1/4 resolution, 1 spp
    r = subprocess.run([sys.executable, "-m", "pbrt_r3_b200", "--quiet", "--quick", "-i", str(scene)], cwd=str(tmp_path), capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr)
    assert imageio.read_exr(str(tmp_path / "cornell.exr")).shape == (12, 12, 3)
    bad = tmp_path / "bad.pbrt"
    bad.write_text('Camera "perspective"\nWorldBegin\nShape "trianglemesh" "integer indices" [0 1 5] "point P" [0 0 0 1 0 0 0 1 0]\nWorldEnd\n')
    r = subprocess.run([sys.executable, "-m", "pbrt_r3_b200", str(bad)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 255 and "ERROR" in r.stderr                                           # exit(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("sampler,spp", [("halton", 8), ("sobol", 6), ("random", 4)])
def test_cli_two_ranks_match_one(P, tmp_path, sampler, spp):
    """`--gpus 2`: two processes (one per GPU over NCCL; on a 1-GPU box both on device 0 with the partial films reduced through
    gloo) render the sample shards of the PRODUCT and rank 0 resolves the sum: the image equals the single-process render.
    Sobol with 6 spp: the sampler takes 8 (rounded up to a power of two) and all 8 sample numbers must be sharded."""
    import subprocess
    import sys
    from pbrt_r3_b200 import imageio
    scene = tmp_path / "s.pbrt"
    scene.write_text(P.scenes.cornell_box(res=80, spp=spp, materials="mixed", extra_lights=True, maxdepth=6, sampler=sampler))
    outs = []
    for g in (1, 2):
        out = tmp_path / ("out_%d.exr" % g)
        env = dict(os.environ, MASTER_PORT=str(29600 + os.getpid() % 300))
        r = subprocess.run([sys.executable, "-m", "pbrt_r3_b200", str(scene), "-o", str(out), "--gpus", str(g), "--quiet"], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
        outs.append(imageio.read_exr(str(out)))
    a, b = outs
    assert a.shape == b.shape and np.isfinite(b).all()
    assert float(np.sqrt(((a - b) ** 2).mean()) / a.mean()) < 1e-5      # the film partials are summed in another order


def test_tev_and_sequential_displays(P, tmp_path):
    """--display-server / --sequential-display (displays/tev/display.rs:155-243 packet layout, 128-pixel tiles :245-330;
    displays/sequential): a mock tev server receives CreateImage + one UpdateImage per tile and reassembles the film."""
    import socket
    import struct
    import threading
    from pbrt_r3_b200 import display, imageio
    W, H = 200, 150
    rgb = np.random.RandomState(1).uniform(0, 2, (H, W, 3)).astype(np.float32)
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    got = bytearray()

    def serve():
        c, _ = srv.accept()
        while True:
            d = c.recv(1 << 16)
            if not d:
                break
            got.extend(d)
        c.close()
    th = threading.Thread(target=serve)
    th.start()

    class HS:                                   # what display.show reads of a HostScene
        filename = "out.exr"
        class desc:
            class film:
                full_resolution = (W + 10, H + 20)
                cropped_pixel_bounds = (5, 10, 5 + W, 10 + H)
    tev = display.TevDisplay("127.0.0.1:%d" % srv.getsockname()[1])
    seq = display.SequentialDisplay(str(tmp_path / "frames"))
    display.show([tev, seq], HS, rgb)
    tev.close()
    th.join(timeout=10)
    srv.close()
    # parse the stream: u32 length (of the whole packet) | u8 type | u8 grab_focus | name\0 | ...
    out = np.zeros((H + 20, W + 10, 3), np.float32)
    pos, kinds = 0, []
    while pos < len(got):
        n, kind, focus = struct.unpack_from("<IBB", got, pos)
        body = bytes(got[pos + 6:pos + n])
        name, body = body.split(b"\0", 1)
        assert name == b"out.exr"
        kinds.append(kind)
        if kind == display.CREATE_IMAGE:
            w, h, nc = struct.unpack_from("<III", body)
            assert (w, h, nc, focus) == (W + 10, H + 20, 3, 1) and body[12:] == b"R\0G\0B\0"
        else:
            assert kind == display.UPDATE_IMAGE and focus == 0
            nc, = struct.unpack_from("<I", body)
            assert body[4:10] == b"R\0G\0B\0"
            x, y, w, h = struct.unpack_from("<IIII", body, 10)
            assert struct.unpack_from("<6q", body, 26) == (0, 1, 2, 3, 3, 3) and w <= 128 and h <= 128
            out[y:y + h, x:x + w] = np.frombuffer(body, "<f4", w * h * 3, 26 + 48).reshape(h, w, 3)
        pos += n
    assert kinds[0] == display.CREATE_IMAGE and kinds.count(display.UPDATE_IMAGE) == 2 * 2     # 200 x 150 in 128-pixel tiles
    assert np.array_equal(out[10:10 + H, 5:5 + W], rgb)
    frame = imageio.read_exr(str(tmp_path / "frames" / "00000000.exr"))
    assert np.array_equal(frame[10:10 + H, 5:5 + W], rgb) and frame.shape == (H + 20, W + 10, 3)
    with pytest.raises(ValueError):
        display.TevDisplay("nohostport")


def test_exr_reader_compressions(tmp_path):
    """read_exr also reads what other encoders write (tools/run_reference.sh: the reference's `image` crate): ZIPS / ZIP /
    RLE scan-line blocks (byte predictor + even / odd interleave, OpenEXR file layout), HALF and FLOAT channels, RGBA."""
    from pbrt_r3_b200 import imageio
    rng = np.random.RandomState(3)
    h, w = 37, 11
    img = rng.rand(h, w, 4).astype(np.float32)

    def predict(raw):
        d = np.frombuffer(raw, np.uint8)
        t = np.concatenate([d[0::2], d[1::2]]).astype(np.int32)
        p = t.copy()
        p[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xff
        return p.astype(np.uint8).tobytes()

    def rle(raw):                       # literal runs only (negative counts): valid RLE
        out = b""
        for i in range(0, len(raw), 100):
            c = raw[i:i + 100]
            out += struct.pack("b", -len(c)) + c
        return out

    for comp, lines, half in ((2, 1, False), (3, 16, False), (3, 16, True), (1, 1, False), (0, 1, True)):
        dt, ptype, bpp = ("<f2", 1, 2) if half else ("<f4", 2, 4)

        def attr(n, t, data):
            return n.encode() + b"\0" + t.encode() + b"\0" + struct.pack("<i", len(data)) + data
        chans = b"".join(c + b"\0" + struct.pack("<iBBBBii", ptype, 0, 0, 0, 0, 1, 1) for c in (b"A", b"B", b"G", b"R")) + b"\0"
        box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
        hdr = struct.pack("<ii", 20000630, 2) + attr("channels", "chlist", chans) + attr("compression", "compression", bytes([comp]))
        hdr += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") + b"\0"
        blocks = []
        for y in range(0, h, lines):
            raw = b""
            for l in range(y, min(y + lines, h)):
                for ch in (3, 2, 1, 0):     # A, B, G, R
                    raw += img[l, :, ch].astype(dt).tobytes()
            enc = raw if comp == 0 else (zlib.compress(predict(raw)) if comp in (2, 3) else rle(predict(raw)))
            if comp != 0 and len(enc) >= len(raw):
                enc = raw               # the format stores a block raw when compression does not shrink it
            blocks.append(struct.pack("<ii", y, len(enc)) + enc)
        table_at = len(hdr)
        pos = table_at + 8 * len(blocks)
        offs = []
        for bl in blocks:
            offs.append(pos)
            pos += len(bl)
        f = tmp_path / ("c%d_%d.exr" % (comp, half))
        f.write_bytes(hdr + struct.pack("<%dQ" % len(offs), *offs) + b"".join(blocks))
        got = imageio.read_exr(str(f))
        assert np.array_equal(got, img[:, :, :3].astype(dt).astype(np.float32)), (comp, half)
