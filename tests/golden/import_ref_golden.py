#!/usr/bin/env python
"""Turns the container written by julia/make_golden.jl (run on a machine with Julia 1.0.x + the reference's env) into the golden fixtures
tests/golden/ref_<group>.npz that tests/test_ref_golden.py (CPU: oracle vs reference) and tests/test_gpu_ref_golden.py (-m gpu: CUDA vs reference) consume.

    python tests/golden/import_ref_golden.py tests/golden/ref_golden.bin [out_dir]

Container "PGNGOLD1": records [u32 name length][name][u8 dtype 1=f64 2=i32 3=f32][u32 ndims][u64 dims...][column-major data].
Julia arrays are column-major: a (d1, ..., dk) Julia array becomes the numpy array of shape (dk, ..., d1), so per-node vectors come out
node-major ([N][k]) as the C ABI has them; stacks of MATRICES (A, B, B0, Bf, H: Julia (rows, cols, T)) are additionally transposed in their
last two axes so that entry [t][i][j] is the matrix entry (i, j)."""
import os
import struct
import sys

import numpy as np

DTYPES = {1: np.float64, 2: np.int32, 3: np.float32}
MATRIX_STACKS = ("A", "B", "B0", "Bf", "H", "M_HJI")


def read_container(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != b"PGNGOLD1":
            raise ValueError(f"{path}: not a PGNGOLD1 container")
        while True:
            head = f.read(4)
            if not head:
                break
            (nlen,) = struct.unpack("<I", head)
            name = f.read(nlen).decode()
            (code,) = struct.unpack("<B", f.read(1))
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack("<%dQ" % nd, f.read(8 * nd))
            dt = np.dtype(DTYPES[code]).newbyteorder("<")
            cnt = int(np.prod(dims)) if nd else 1
            a = np.frombuffer(f.read(cnt * dt.itemsize), dtype=dt).reshape(tuple(reversed(dims))).astype(DTYPES[code])
            if name.rsplit("/", 1)[-1] in MATRIX_STACKS and a.ndim >= 2:
                a = np.swapaxes(a, -1, -2)
            out[name] = np.ascontiguousarray(a)
    return out


def group_of(name):
    top = name.split("/")
    if top[0] in ("sim", "sim25"):
        return f"{top[0]}_{top[1]}"
    return top[0]


def main():
    src = sys.argv[1]
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.dirname(os.path.abspath(src))
    groups = {}
    for name, a in read_container(src).items():
        groups.setdefault(group_of(name), {})[name] = a
    for g, d in groups.items():
        path = os.path.join(out_dir, f"ref_{g}.npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {len(d)} arrays")


if __name__ == "__main__":
    main()
