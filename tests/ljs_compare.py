"""Field-by-field comparison of two flat scene descriptions (lajolla_public_b200.ljs.SceneDesc): the host front end's
parse of a scene XML against the oracle's dump of the reference's own parse_scene() of the same file."""
import dataclasses

import numpy as np


def _cmp(path, a, b, out, rtol, atol):
    if dataclasses.is_dataclass(a):
        for f in dataclasses.fields(a):
            _cmp(f"{path}.{f.name}", getattr(a, f.name), getattr(b, f.name), out, rtol, atol)
    elif isinstance(a, (list, tuple)) and a and (dataclasses.is_dataclass(a[0]) or isinstance(a[0], np.ndarray)):
        if len(a) != len(b):
            out.append(f"{path}: {len(a)} entries vs {len(b)}")
            return
        for i, (x, y) in enumerate(zip(a, b)):
            _cmp(f"{path}[{i}]", x, y, out, rtol, atol)
    elif a is None or b is None:
        if (a is None) != (b is None):
            out.append(f"{path}: {'missing' if a is None else 'present'} vs {'missing' if b is None else 'present'}")
    else:
        x, y = np.asarray(a), np.asarray(b)
        if x.shape != y.shape:
            out.append(f"{path}: shape {x.shape} vs {y.shape}")
        elif x.dtype.kind in "iub" or y.dtype.kind in "iub":
            if not np.array_equal(x, y):
                out.append(f"{path}: {x.tolist() if x.size < 8 else '...'} vs {y.tolist() if y.size < 8 else '...'} ({int((x != y).sum())} of {x.size} differ)")
        else:
            x64, y64 = x.astype(np.float64), y.astype(np.float64)
            bad = ~(np.isclose(x64, y64, rtol=rtol, atol=atol) | (np.isnan(x64) & np.isnan(y64)))
            if bad.any():
                k = np.argmax(np.abs(x64 - y64) * bad)
                out.append(f"{path}: {int(bad.sum())} of {x.size} values differ, worst {x64.flat[k]!r} vs {y64.flat[k]!r}")


def compare(mine, ref, rtol=2e-6, atol=1e-7):
    """List of human-readable differences (empty = equal up to double-vs-float rounding of derived quantities)."""
    out = []
    _cmp("scene", mine, ref, out, rtol, atol)
    return out


def exact_fraction(mine, ref):
    """(number of float values bit-identical, total) over every float array of the two descriptions."""
    same = total = 0

    def walk(a, b):
        nonlocal same, total
        if dataclasses.is_dataclass(a):
            for f in dataclasses.fields(a):
                walk(getattr(a, f.name), getattr(b, f.name))
        elif isinstance(a, (list, tuple)) and a and (dataclasses.is_dataclass(a[0]) or isinstance(a[0], np.ndarray)):
            for x, y in zip(a, b):
                walk(x, y)
        elif a is not None and b is not None:
            x, y = np.asarray(a), np.asarray(b)
            if x.shape == y.shape and x.dtype.kind == "f":
                same += int((x.astype(np.float32).view(np.uint32) == y.astype(np.float32).view(np.uint32)).sum())
                total += x.size

    walk(mine, ref)
    return same, total
