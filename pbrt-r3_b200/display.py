"""Film displays of the reference's command line (`--display-server host:port`, `-k / --sequential-display dir`):
`TevDisplay` streams the image to a running tev viewer over its IPC protocol (reference src/displays/tev/display.rs:
IPCChannel :22-150, IPCGen::create_image :155-181 / update_image :183-243, 128-pixel tiles :245-330; tev_display.rs),
`SequentialDisplay` writes numbered EXR frames (src/displays/sequential/sequential_display.rs).  Film::render_start /
update_display / render_end (core/film/film.rs:278-352, 424-438) drive them.

The reference updates the displays after every 16x16 render tile it merges; the wavefront has no per-tile completion
order (a wave covers the whole film), so the GPU path sends ONE update per render: the finished film, cut into the same
128 x 128 tiles."""
import os
import socket
import struct

import numpy as np

TILE_SIZE = 128          # display.rs:245
CREATE_IMAGE, UPDATE_IMAGE = 4, 6   # DisplayDirective (display.rs:11-18)


def create_image_packet(name, width, height, channel_names):
    """IPCGen::create_image: u32 length | u8 4 | u8 grab_focus = 1 | name\\0 | u32 width, height, n_channels | channel names\\0"""
    b = struct.pack("<iBB", 0, CREATE_IMAGE, 1) + name.encode() + b"\0"
    b += struct.pack("<III", width, height, len(channel_names))
    for c in channel_names:
        b += c.encode() + b"\0"
    return struct.pack("<I", len(b)) + b[4:]


def update_image_packet(name, x, y, width, height, image, channel_names):
    """IPCGen::update_image: u32 length | u8 6 | u8 grab_focus = 0 | name\\0 | u32 n_channels | channel names\\0 | u32 x, y, w, h |
    i64 channel offsets 0..n | i64 channel strides n.. | f32 pixels, interleaved"""
    n = len(channel_names)
    img = np.ascontiguousarray(image, "<f4").reshape(height, width, n)
    b = struct.pack("<iBB", 0, UPDATE_IMAGE, 0) + name.encode() + b"\0" + struct.pack("<I", n)
    for c in channel_names:
        b += c.encode() + b"\0"
    b += struct.pack("<IIII", x, y, width, height)
    b += b"".join(struct.pack("<q", i) for i in range(n)) + b"".join(struct.pack("<q", n) for _ in range(n))
    b += img.tobytes()
    return struct.pack("<I", len(b)) + b[4:]


def gen_tiles(x, y, width, height, image):
    """DisplayItem::gen_tiles: the update region cut into tiles of at most 128 x 128, row-major"""
    n = image.shape[-1]
    img = image.reshape(height, width, n)
    for y0 in range(0, height, TILE_SIZE):
        for x0 in range(0, width, TILE_SIZE):
            y1, x1 = min(y0 + TILE_SIZE, height), min(x0 + TILE_SIZE, width)
            yield x + x0, y + y0, x1 - x0, y1 - y0, img[y0:y1, x0:x1]


class TevDisplay:
    def __init__(self, hostname):
        parts = hostname.split(":")
        if len(parts) < 2:
            raise ValueError('Expected "host:port" for display server address. Given "%s".' % hostname)
        self.addr = (":".join(parts[:-1]), int(parts[-1]))
        self.sock = socket.create_connection(self.addr, timeout=10)   # IPv4 first, like the reference
        self.title, self.channel_names = None, None

    def start(self, title, resolution, channel_names):
        self.title, self.channel_names = title, list(channel_names)
        self.sock.sendall(create_image_packet(title, resolution[0], resolution[1], self.channel_names))

    def update(self, x, y, width, height, buffer):
        for tx, ty, tw, th, tile in gen_tiles(x, y, width, height, np.asarray(buffer, np.float32).reshape(height, width, -1)):
            self.sock.sendall(update_image_packet(self.title, tx, ty, tw, th, tile, self.channel_names))

    def end(self):
        pass

    def close(self):
        try:
            self.sock.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        self.sock.close()


class SequentialDisplay:
    def __init__(self, output_dir):
        self.output_dir, self.count, self.buffer = output_dir, 0, None
        os.makedirs(output_dir, exist_ok=True)

    def start(self, title, resolution, channel_names):
        self.buffer = np.zeros((resolution[1], resolution[0], len(channel_names)), np.float32)
        self.count = 0

    def update(self, x, y, width, height, buffer):
        from . import imageio
        self.buffer[y:y + height, x:x + width] = np.asarray(buffer, np.float32).reshape(height, width, -1)
        imageio.write_image(os.path.join(self.output_dir, "%08d.exr" % self.count), self.buffer)
        self.count += 1

    def end(self):
        pass

    def close(self):
        pass


def show(displays, host_scene, rgb):
    """Film::render_start + one update_display over the cropped pixel bounds + render_end"""
    if not displays:
        return
    f = host_scene.desc.film
    cb = f.cropped_pixel_bounds
    res = (int(f.full_resolution[0]), int(f.full_resolution[1]))
    for d in displays:
        d.start(host_scene.filename, res, ["R", "G", "B"])
        d.update(int(cb[0]), int(cb[1]), int(cb[2] - cb[0]), int(cb[3] - cb[1]), rgb)
        d.end()
