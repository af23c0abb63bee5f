"""Readers for the image files the `lajolla` command writes (PFM; EXR through the front end's own reader is exercised
by tests/test_frontend.py), for tests and tools."""
import numpy as np


def read_pfm(path):
    """(h, w, 3) or (h, w) float32.  Rows are stored top-down, as the reference's imwrite leaves them (image.cpp:139-149)."""
    with open(path, "rb") as f:
        magic = f.readline().strip()
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        ch = 3 if magic == b"PF" else 1
        data = np.frombuffer(f.read(w * h * ch * 4), dtype="<f4" if scale < 0 else ">f4")
    return data.reshape((h, w, ch) if ch == 3 else (h, w)).astype(np.float32)
