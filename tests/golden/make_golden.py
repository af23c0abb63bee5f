#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the reference's own object code (needs oracle/_ref built from
/root/reference: `python -c "import __graft_entry__ as g; g.build()"`).  Run from the repo root:
    python tests/golden/make_golden.py
Each scene file holds the oracle's answers to the query batches of tests/golden_lib.py:run_checks (ray casts,
PathVertex assembly, BSDF eval/pdf/sample incl. the fixed tuple of the reference's tests/materials.cpp:57-62,
light sampling, camera rays, media).  pcg32.npz holds raw PCG32 outputs (bit-exact integer KAT, pcg.h:22-47);
cbox_render.npz the reference's own render() of cbox at 16 spp reduced to 16x16-pixel tile means.
The device code used as `scene` while recording is the host simulation (tests/hostsim); it only has to pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import golden_lib  # noqa: E402
import hostsim_lib  # noqa: E402
import lajolla_public_b200 as lj  # noqa: E402
import oracle_lib  # noqa: E402

SCENES = ["cbox", "veach_mi", "matpreview", "disney_bsdf", "volpath_test6", "hetvol"]


def main():
    for name in SCENES:
        with hostsim_lib.simulated():
            sc = lj.parse_scene(oracle_lib.scene_ljs(name))
        rec = golden_lib.RecordingRef(oracle_lib.RefScene(oracle_lib.scene_xml(name), threads=2))
        print(name, golden_lib.run_checks(sc, rec, name))
        rec.save(name)
    u, f = oracle_lib.pcg32(0, 8, 32)
    u2, f2 = oracle_lib.pcg32(2 ** 40 + 17, 4, 16, seed=42)
    np.savez_compressed(os.path.join(golden_lib.GOLDEN_DIR, "pcg32.npz"), u_default=u, f_default=f, u_seed42=u2, f_seed42=f2)
    ref = oracle_lib.RefScene(oracle_lib.scene_xml("cbox"))
    img, _ = ref.render(spp=16)
    tiles = img.reshape(32, 16, 32, 16, 3).mean(axis=(1, 3))
    np.savez_compressed(os.path.join(golden_lib.GOLDEN_DIR, "cbox_render.npz"), tiles=tiles.astype(np.float32), spp=np.int32(16))
    print("tiles mean", tiles.mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
