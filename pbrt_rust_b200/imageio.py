"""Minimal PNG reader for image textures: what `image::open(..).into_rgb8()` hands to
texture/imagemap.rs:75-89 (8-bit RGB, alpha dropped, grey replicated, palettes expanded).
Non-interlaced 8-bit PNGs only — enough for texture fixtures; anything else raises."""
import struct
import zlib

import numpy as np


def read_png_rgb8(path):
    raw = open(path, "rb").read()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, plte, hdr = 8, [], None, None
    while pos < len(raw):
        ln, ty = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + ln]
        if ty == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif ty == b"PLTE":
            plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif ty == b"IDAT":
            idat.append(body)
        elif ty == b"IEND":
            break
        pos += 12 + ln
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or interlace != 0:
        raise ValueError(f"{path}: only non-interlaced 8-bit PNGs are supported")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    data = zlib.decompress(b"".join(idat))
    stride = w * ch
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        f = data[y * (stride + 1)]
        line = np.frombuffer(data, np.uint8, stride, y * (stride + 1) + 1).astype(np.int32)
        cur = np.zeros(stride, np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:  # filters with a left-neighbour dependency: sequential per byte
            for i in range(stride):
                a = cur[i - ch] if i >= ch else 0
                b = prev[i]
                c = prev[i - ch] if i >= ch else 0
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) >> 1
                elif f == 4:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad filter type {f}")
                cur[i] = (line[i] + p) & 255
        out[y] = cur
        prev = cur
    px = out.reshape(h, w, ch)
    if ctype == 0:
        rgb = np.repeat(px, 3, axis=2)
    elif ctype == 2:
        rgb = px
    elif ctype == 3:
        rgb = plte[px[..., 0]]
    elif ctype == 4:
        rgb = np.repeat(px[..., :1], 3, axis=2)
    else:
        rgb = px[..., :3]
    return np.ascontiguousarray(rgb, np.uint8)


def read_image(path):
    """read_image (imagemap.rs:75-89): (h, w, 3) float32 texels = byte / 255."""
    return read_png_rgb8(path).astype(np.float32) / np.float32(255)
