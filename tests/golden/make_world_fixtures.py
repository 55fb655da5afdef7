#!/usr/bin/env python
"""Convert the reference's path fixtures (/root/reference/test/path/*.world, YAML with comma-separated float lists)
into small .npz files committed next to this script.  Run in the build container only (the GPU box has no
/root/reference).  Keys keep the .world names: s_m,posE_m,posN_m,psi_rad,k_1pm,grade_rad,edgeL_m,edgeR_m,UxDes_mps,AxDes_mps2,isOpen.
"""
import os
import sys

import numpy as np
import yaml

SRC = "/root/reference/test/path"
DST = os.path.dirname(os.path.abspath(__file__))
NAMES = ["skidpadoval", "vail", "westpaddock", "curvy", "EastPaddock", "flidpadoval", "newskidpadoval", "paddockoval"]

for name in NAMES:
    with open(os.path.join(SRC, name + ".world")) as f:
        d = yaml.safe_load(f)
    out = {}
    for k, v in d.items():
        if isinstance(v, str):
            out[k] = np.array([float(x) for x in v.split(",")], dtype=np.float64)
        else:
            out[k] = np.array(v)
    np.savez_compressed(os.path.join(DST, f"world_{name}.npz"), **out)
    print(name, {k: (a.shape, float(np.min(a)), float(np.max(a))) for k, a in out.items() if a.ndim})

# Serialised path messages (test/path/*.msg): the raw bytes of the two small / representative ones travel as uint8 arrays so that the reader is
# tested on the reference's own bytes on any box; for every fixture with a .world twin the decoded arrays are compared with the twin here.
sys.path.insert(0, os.path.dirname(os.path.dirname(DST)))
import importlib  # noqa: E402
world = importlib.import_module("pigeon.jl_b200.world")
raw = {}
for name in NAMES + ["variable_speed"]:
    m = world.read_msg(os.path.join(SRC, name + ".msg"))
    if name != "variable_speed":
        w = np.load(os.path.join(DST, f"world_{name}.npz"))
        worst = max(float(np.max(np.abs(m[k] - w[k]))) for k in world.WORLD_KEYS)
        assert worst < 1e-9 and m["isOpen"] == int(w["isOpen"]), (name, worst)
        print(name, ".msg == .world to", worst, "frame_id", m["frame_id"])
    if name in ("variable_speed", "curvy"):
        raw[name] = np.fromfile(os.path.join(SRC, name + ".msg"), dtype=np.uint8)
np.savez_compressed(os.path.join(DST, "msg_raw.npz"), **raw)
