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
