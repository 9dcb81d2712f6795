"""Film::write_image's file writers (reference src/core/imageio/write_image.rs:59-76): the extension selects the format,
".exr" keeps the float values (write_image_exr.rs:6-45, Rgb32F), anything else is 8-bit with the sRGB transfer curve
(write_image.rs:16-18,20-56: to_byte = clamp(255 * gamma_correct(v), 0, 255) as u8, core/base/functions.rs:13-19).
Pure host code on the already resolved RGB buffer; no third-party codecs: EXR = uncompressed scanlines with 32-bit
float channels, PNG through zlib, plus PFM for tools that read neither."""
import struct
import zlib

import numpy as np


def gamma_correct(v):
    v = np.asarray(v, np.float32)
    return np.where(v <= np.float32(0.0031308), np.float32(12.92) * v,
                    np.float32(1.055) * np.power(np.maximum(v, 0).astype(np.float32), np.float32(1.0 / 2.4)) - np.float32(0.055)).astype(np.float32)


def to_bytes(rgb):
    g = np.float32(255.0) * gamma_correct(rgb)
    g = np.where(np.isnan(g), 0, g)           # Rust `as u8` saturates and maps NaN to 0
    return np.clip(g, 0.0, 255.0).astype(np.uint8)


def write_exr(name, rgb):
    h, w, _ = rgb.shape
    img = np.ascontiguousarray(rgb, np.float32)

    def attr(n, t, data):
        return n.encode() + b"\0" + t.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chans = b"".join(c + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"   # FLOAT, alphabetical
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<ii", 20000630, 2)
    hdr += attr("channels", "chlist", chans) + attr("compression", "compression", b"\0") + attr("dataWindow", "box2i", box)
    hdr += attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    line_bytes = 3 * 4 * w
    table_at = len(hdr)
    first = table_at + 8 * h
    offsets = struct.pack("<%dQ" % h, *[first + y * (8 + line_bytes) for y in range(h)])
    with open(name, "wb") as f:
        f.write(hdr)
        f.write(offsets)
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(img[y, :, 2].tobytes() + img[y, :, 1].tobytes() + img[y, :, 0].tobytes())


def _exr_unpredict(raw):
    """undo the byte predictor + even / odd interleave of the RLE / ZIP scan-line compressions"""
    d = np.frombuffer(raw, np.uint8).astype(np.int32)
    d[1:] -= 128
    d = (np.cumsum(d) & 0xff).astype(np.uint8)
    n = len(d)
    out = np.empty(n, np.uint8)
    half = (n + 1) // 2
    out[0::2] = d[:half]
    out[1::2] = d[half:]
    return out.tobytes()


def _exr_unrle(raw, size):
    out = bytearray()
    i = 0
    while i < len(raw) and len(out) < size:
        c = raw[i] - 256 if raw[i] > 127 else raw[i]
        i += 1
        if c < 0:
            out += raw[i:i - c]
            i += -c
        else:
            out += bytes([raw[i]]) * (c + 1)
            i += 1
    return bytes(out)


def read_exr(name):
    """Scan-line OpenEXR reader: R, G, B (any extra channels ignored) as HALF or FLOAT, compression NONE / RLE / ZIPS / ZIP:
    the files write_exr produces and the ones the reference's `image` crate encoder writes (tools/run_reference.sh)."""
    b = open(name, "rb").read()
    assert struct.unpack("<i", b[:4])[0] == 20000630, "not an OpenEXR file"
    version = struct.unpack("<i", b[4:8])[0]
    assert not (version & 0x200), "tiled OpenEXR files are not supported"
    p = 8
    attrs = {}
    while b[p] != 0:
        e = b.index(b"\0", p); n = b[p:e].decode(); p = e + 1
        e = b.index(b"\0", p); p = e + 1
        sz = struct.unpack("<i", b[p:p + 4])[0]; p += 4
        attrs[n] = b[p:p + sz]; p += sz
    p += 1
    chans = []
    c, q = attrs["channels"], 0
    while c[q] != 0:
        e = c.index(b"\0", q)
        nm = c[q:e].decode()
        ptype = struct.unpack("<i", c[e + 1:e + 5])[0]
        chans.append((nm, ptype))
        q = e + 17
    comp = attrs["compression"][0]
    assert comp in (0, 1, 2, 3), "unsupported OpenEXR compression %d" % comp
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines_per_block = 16 if comp == 3 else 1
    nblocks = (h + lines_per_block - 1) // lines_per_block
    offs = struct.unpack("<%dQ" % nblocks, b[p:p + 8 * nblocks])
    bpp = {0: 4, 1: 2, 2: 4}
    dt = {0: "<u4", 1: "<f2", 2: "<f4"}
    line_bytes = sum(bpp[t] for _, t in chans) * w
    out = np.zeros((h, w, 3), np.float32)
    for off in offs:
        y, size = struct.unpack("<ii", b[off:off + 8])
        nl = min(lines_per_block, y1 - y + 1)
        raw = b[off + 8:off + 8 + size]
        want = nl * line_bytes
        if comp != 0 and size < want:
            raw = _exr_unpredict(zlib.decompress(raw) if comp in (2, 3) else _exr_unrle(raw, want))
        q = 0
        for l in range(nl):
            for nm, t in chans:            # channels are stored in alphabetical order, one after the other per scan line
                row = np.frombuffer(raw[q:q + bpp[t] * w], dt[t]).astype(np.float32)
                q += bpp[t] * w
                if nm in ("R", "G", "B"):
                    out[y - y0 + l, :, "RGB".index(nm)] = row
    return out


def write_png(name, rgb):
    img = to_bytes(rgb)
    h, w, _ = img.shape
    raw = b"".join(b"\0" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(name, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_pfm(name, rgb):
    h, w, _ = rgb.shape
    with open(name, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(rgb[::-1], "<f4").tobytes())


def write_ppm(name, rgb):
    img = to_bytes(rgb)
    with open(name, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def write_image(name, rgb):
    """write_image (write_image.rs:59-76): rgb = H x W x 3 float32 of the cropped pixel bounds."""
    ext = name.rsplit(".", 1)[-1].lower() if "." in name.rsplit("/", 1)[-1] else ""
    if ext == "":
        raise ValueError("write_image: error")          # the reference fails on names without an extension
    if ext == "exr":
        write_exr(name, rgb)
    elif ext == "pfm":
        write_pfm(name, rgb)
    elif ext == "ppm":
        write_ppm(name, rgb)
    elif ext == "png":
        write_png(name, rgb)
    else:
        raise ValueError("write_image: unsupported image format .%s (exr, png, pfm, ppm)" % ext)
