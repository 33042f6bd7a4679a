"""Minimal OpenEXR writer/reader for the host side (the reference writes <scene>.exr through
Bitmap::save, bitmap.cpp:82-107: three FLOAT channels R,G,B).  Scan-line, uncompressed, float32."""
import struct

import numpy as np


def write_exr(path, rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w, _ = rgb.shape

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(c + b"\0" + struct.pack("<iBxxxii", 2, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    head = struct.pack("<II", 20000630, 2)
    head += attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0")
    head += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box)
    head += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    head += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0))
    head += attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    line = 8 + 12 * w
    table_at = len(head)
    offsets = struct.pack(f"<{h}Q", *[table_at + 8 * h + line * y for y in range(h)])
    body = bytearray()
    for y in range(h):
        body += struct.pack("<ii", y, 12 * w)
        body += rgb[y, :, 2].tobytes() + rgb[y, :, 1].tobytes() + rgb[y, :, 0].tobytes()
    with open(path, "wb") as f:
        f.write(head + offsets + bytes(body))


def read_exr(path):
    """Reads what write_exr (and the oracle's bitmap shim) writes: uncompressed FLOAT B,G,R."""
    b = open(path, "rb").read()
    if struct.unpack_from("<I", b, 0)[0] != 20000630:
        raise ValueError("not an OpenEXR file")
    p, chans, comp, dw = 8, [], 0, None
    while b[p]:
        e = b.index(b"\0", p); name = b[p:e].decode(); p = e + 1
        e = b.index(b"\0", p); p = e + 1
        (size,) = struct.unpack_from("<i", b, p); p += 4
        if name == "channels":
            q = p
            while b[q]:
                e = b.index(b"\0", q); cname = b[q:e].decode(); q = e + 1
                (typ,) = struct.unpack_from("<i", b, q); q += 16
                chans.append((cname, typ))
        elif name == "compression":
            comp = b[p]
        elif name == "dataWindow":
            dw = struct.unpack_from("<iiii", b, p)
        p += size
    p += 1
    if comp != 0 or any(t != 2 for _, t in chans):
        raise ValueError("read_exr only handles uncompressed FLOAT files")
    w, h = dw[2] - dw[0] + 1, dw[3] - dw[1] + 1
    out = np.zeros((h, w, 3), np.float32)
    for y in range(h):
        (off,) = struct.unpack_from("<Q", b, p + 8 * y)
        row = np.frombuffer(b, np.float32, len(chans) * w, off + 8).reshape(len(chans), w)
        for i, (cname, _) in enumerate(chans):
            k = {"R": 0, "G": 1, "B": 2}.get(cname[-1].upper())
            if k is not None:
                out[y, :, k] = row[i]
    return out
