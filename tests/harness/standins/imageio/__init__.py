"""Harness stand-in for ``imageio`` (not installed offline): 8-bit RGB/RGBA PNG read/write through zlib, and a
``mimwrite`` that stores the frames as .npy next to the requested video name.  TEST HARNESS ONLY."""
import struct
import zlib

import numpy as np

written = []  # (path, shape) of everything written through this module, for the harness' assertions


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def imwrite(path, img, **_kw):
    a = np.asarray(img)
    if a.dtype != np.uint8:
        a = (255 * np.clip(a, 0, 1)).astype(np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    color = {1: 0, 3: 2, 4: 6}[c]
    raw = b"".join(b"\x00" + a[r].tobytes() for r in range(h))
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0)) \
        + _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b"")
    with open(path, "wb") as fh:
        fh.write(png)
    written.append((str(path), a.shape))


def imread(path, **_kw):
    with open(path, "rb") as fh:
        data = fh.read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "stand-in imageio reads PNG only"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if tag == b"IHDR":
            w, h, depth, color, _c, _f, interlace = struct.unpack(">IIBBBBB", body)
            assert depth == 8 and interlace == 0 and color in (0, 2, 6)
            ch = {0: 1, 2: 3, 6: 4}[color]
        elif tag == b"IDAT":
            idat += body
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.uint8)
    prev = np.zeros(w * ch, np.int32)
    for r in range(h):
        f, line = int(raw[r, 0]), raw[r, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:  # sub / average / paeth: sequential along the row
            cur = np.zeros_like(line)
            for i in range(w * ch):
                a = cur[i - ch] if i >= ch else 0
                b = prev[i]
                c = prev[i - ch] if i >= ch else 0
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b) // 2
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        out[r] = cur
        prev = cur
    return out.reshape(h, w, ch) if ch > 1 else out.reshape(h, w)


def mimwrite(path, frames, **_kw):
    a = np.asarray(frames)
    np.save(str(path) + ".npy", a)
    written.append((str(path), a.shape))
