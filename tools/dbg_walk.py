#!/usr/bin/env python
"""Debug: which walks differ between the staged walk kernels and the serial walk."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import lajolla_public_b200 as lj
from lajolla_public_b200 import abi
import oracle_lib, parity_checks as pc
name = sys.argv[1] if len(sys.argv) > 1 else "volpath_test6"
sc = lj.parse_scene(oracle_lib.scene_ljs(name)); ref = oracle_lib.RefScene(oracle_lib.scene_xml(name))
q = pc.make_walk_queries(sc, ref, 1 << 16)
base = sc.nee_walks(q)
for k, nm in ((abi.LJ_TRACE_WALK_WHOLE, "whole"), (abi.LJ_TRACE_WALK_STEP, "step"), (abi.LJ_TRACE_WALK_STAGED, "staged")):
    got = sc.nee_walks(q, kernel=k)
    bad = (got.view(np.uint32) != base.view(np.uint32)).any(axis=1)
    print(name, nm, "differ:", int(bad.sum()))
    if bad.any():
        idx = np.nonzero(bad)[0][:8]
        for i in idx:
            print("  walk", i, "medium", q["medium_id"][i], "budget", q["budget"][i], "got", got[i], "base", base[i], "ulps", (got[i].view(np.int32) - base[i].view(np.int32)))
        print("  media of differing walks:", np.unique(q["medium_id"][bad], return_counts=True), " budgets:", np.unique(q["budget"][bad], return_counts=True))
        print("  zero contributions among differing:", int((np.abs(base[bad]).max(axis=1) == 0).sum()))
