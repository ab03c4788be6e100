"""The reference's raster cache: the directory `<map>_raster_cache/` that TopDownMap writes after
rasterising an svg map (saveRasterizedMaps, top_down_map.cpp:197-211) and reads instead of a map file
(loadRasterizedMaps, :213-224), so that caches written by either implementation load in the other.

One file per flattened class, `class<i>.png`: the binary class map (0 inside the class, 1 elsewhere) times 255 as an
8-bit single-channel PNG, flipped vertically "to look like the input map".  Loading flips back and scales by 1/255
(cv::Mat::convertTo with a float factor: 255 * float(1/255) == 1.0f exactly), and computeDists then rounds whatever
the file held to 0 / 1 (convertTo CV_8U, :306).

The PNG codec is the third-party part (OpenCV's bundled libpng).  Its FORMAT is pinned by the PNG specification, so
this module carries a small dependency-free codec (zlib from the standard library): 8-bit grayscale, non-interlaced,
the five scan-line filters on the read side, filter 0 on the write side.  tests/test_host_math.py checks both
directions against cv2.imwrite / cv2.imread where cv2 is installed.

Arrays use this repo's layout for col-major rows x cols images: numpy shape (cols, rows), C-contiguous."""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png_gray(path: str, img: np.ndarray, level: int = 6) -> None:
    """img: (height, width) uint8, row 0 = top line"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    raw = np.zeros((h, w + 1), dtype=np.uint8)      # filter byte 0 (None) in front of every scan line
    raw[:, 1:] = img
    with open(path, "wb") as f:
        f.write(_SIG)
        f.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)))
        f.write(_chunk(b"IDAT", zlib.compress(raw.tobytes(), level)))
        f.write(_chunk(b"IEND", b""))


def read_png_gray(path: str) -> np.ndarray:
    """-> (height, width) uint8.  8-bit grayscale, non-interlaced files only (what cv::imwrite produces for CV_8UC1)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != _SIG:
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if len(body) != n or struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] != (zlib.crc32(tag + body) & 0xFFFFFFFF):
            raise ValueError(f"{path}: damaged chunk {tag!r}")
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
        pos += 12 + n
    if hdr is None:
        raise ValueError(f"{path}: no IHDR")
    w, h, depth, colour, _, _, interlace = hdr
    if (depth, colour, interlace) != (8, 0, 0):
        raise ValueError(f"{path}: only 8-bit grayscale non-interlaced PNGs are raster-cache files (depth {depth}, colour type {colour})")
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8)
    if raw.size != h * (w + 1):
        raise ValueError(f"{path}: {raw.size} bytes of image data for {w} x {h}")
    raw = raw.reshape(h, w + 1)
    out = np.zeros((h, w), dtype=np.uint8)
    prev = np.zeros(w, dtype=np.uint8)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:]
        if ft == 0:
            cur = line.copy()
        elif ft == 2:                               # Up
            cur = line + prev
        elif ft == 1:                               # Sub: a running sum modulo 256 (bytes per pixel = 1)
            cur = np.cumsum(line, dtype=np.uint64).astype(np.uint8)
        elif ft in (3, 4):                          # Average / Paeth depend on the pixel just reconstructed
            cur = np.zeros(w, dtype=np.uint8)
            a = c = 0
            for x in range(w):
                b = int(prev[x])
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
                a = (int(line[x]) + pred) & 255
                cur[x] = a
                c = b
        else:
            raise ValueError(f"{path}: filter type {ft}")
        out[y] = cur
        prev = cur
    return out


def layer_to_image(layer: np.ndarray) -> np.ndarray:
    """(cols, rows) binary float layer -> the (rows, cols) uint8 image saveRasterizedMaps writes: x255 with
    saturate_cast<uchar> (round half to even, clamped), then flipped vertically (:205-208)"""
    img = np.clip(np.rint(np.asarray(layer, dtype=np.float32).T * np.float32(255)), 0, 255).astype(np.uint8)
    return img[::-1]


def image_to_layer(img: np.ndarray) -> np.ndarray:
    """the inverse path of loadRasterizedMaps (:217-222): flip back, scale by float(1/255) -> (cols, rows) float32"""
    return np.ascontiguousarray((img[::-1].astype(np.float32) * np.float32(1.0 / 255)).T)


def save_rasterized_maps(cache_dir: str, layers: np.ndarray) -> None:
    """layers (C, cols, rows) binary class maps as getRasterMap leaves them (BEFORE computeDists)"""
    os.makedirs(cache_dir, mode=0o700, exist_ok=True)
    for i, layer in enumerate(layers):
        write_png_gray(os.path.join(cache_dir, f"class{i}.png"), layer_to_image(layer))


def load_rasterized_maps(cache_dir: str, num_classes: int) -> np.ndarray:
    """-> (C, cols, rows) float32 class maps, ready for tdr_map_set_binary_layers"""
    return np.stack([image_to_layer(read_png_gray(os.path.join(cache_dir, f"class{i}.png"))) for i in range(num_classes)])
