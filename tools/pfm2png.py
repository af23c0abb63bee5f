"""Tone-map a PFM (as written by lajolla's imwrite) to an sRGB PNG for eyeballing."""
import sys
import numpy as np
from PIL import Image


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(w * h * 12), dtype="<f4" if scale < 0 else ">f4")
    # lajolla writes rows top-to-bottom (it does not flip like the PFM convention)
    return data.reshape(h, w, 3).astype(np.float32)


def tonemap(img, exposure=1.0):
    x = np.clip(img * exposure, 0, 1)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1 / 2.4) - 0.055)


if __name__ == "__main__":
    img = read_pfm(sys.argv[1])
    exp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    Image.fromarray((tonemap(img, exp) * 255 + 0.5).astype(np.uint8)).save(sys.argv[2])
