"""Headless image output (replaces the reference's Silk.NET/ImGui window, Film.fs:38-92)."""
import numpy as np


def texture_to_rows(texture_wh):
    """Color[w,h] (width, height, 4) -> row-major (height, width, 3) f32."""
    return np.ascontiguousarray(np.transpose(texture_wh[:, :, :3], (1, 0, 2))).astype(np.float32)


def write_pfm(path, rows_rgb):
    """rows_rgb: (height, width, 3) float32, row 0 = top.  PFM stores bottom-up, little-endian."""
    img = np.asarray(rows_rgb, dtype="<f4")
    h, w, _ = img.shape
    with open(path, "wb") as fh:
        fh.write(f"PF\n{w} {h}\n-1.0\n".encode("ascii"))
        fh.write(np.ascontiguousarray(img[::-1]).tobytes())


def write_png(path, rgba8_rows):
    """rgba8_rows: (height, width, 4) uint8 as produced by Film.PostProcess().  Plain zlib PNG, no imaging library."""
    import struct
    import zlib
    img = np.ascontiguousarray(np.asarray(rgba8_rows, np.uint8))
    h, w, c = img.shape
    assert c == 4
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
                 chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
