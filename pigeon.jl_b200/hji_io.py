"""On-disk forms of the HJI cache.

The reference keeps `grid_knots`, `V_raw` and `∇V_raw` in a JLD2 file (`HJICache(fname)` / `save`, src/HJI_computation.jl:39-64; the file is
downloaded by deps/build.jl).  `load_hji_cache` / `save_hji_cache` read and write
  * that JLD2 container (an HDF5 dialect) through the minimal parser / writer of pigeon.jl_b200/jld2.py — any path ending in `.jld2`, or any
    file that starts with JLD2's text header;
  * the flat little-endian dump below, which `julia/export_hji_cache.jl` produces where Julia + JLD2.jl are installed (a fallback that does not
    depend on the container format at all):

    bytes 0..7     magic  b"PGNHJI1\\0"
    7 x int32      grid dimensions n1..n7
    sum(n) x f32   knots of dimension 1, 2, ..., 7 (grid_knots)
    prod(n) x f32  V_raw, dimension 1 fastest (Julia Array{Float32,7} memory order)
    7*prod(n) x f32  ∇V_raw, the 7 components fastest, then dimension 1 (Julia Array{Float32,8} of size (7, n1, ..., n7))

which is also the argument layout of pgn_set_hji_cache."""
import numpy as np

from .mpc import HJICache

MAGIC = b"PGNHJI1\x00"


def save_hji_cache(path, cache):
    if str(path).endswith(".jld2"):
        from .jld2 import save_hji_jld2
        return save_hji_jld2(path, cache.grid_knots, cache.V, cache.gradV)
    dims = np.array([len(k) for k in cache.grid_knots], dtype="<i4")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(dims.tobytes())
        f.write(np.concatenate(cache.grid_knots).astype("<f4").tobytes())
        f.write(np.ascontiguousarray(cache.V.ravel(order="F"), dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(cache.gradV.ravel(order="F"), dtype="<f4").tobytes())


def load_hji_cache(path):
    with open(path, "rb") as f:
        head = f.read(16)
    if str(path).endswith(".jld2") or head.startswith(b"Julia data file") or head.startswith(b"\x89HDF"):
        from .jld2 import load_hji_jld2
        return HJICache(*load_hji_jld2(path))
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a PGNHJI1 file")
        dims = np.frombuffer(f.read(28), dtype="<i4")
        if len(dims) != 7 or np.any(dims < 2):
            raise ValueError(f"{path}: bad grid dimensions {dims}")
        nk, nn = int(dims.sum()), int(np.prod(dims.astype(np.int64)))
        knots = np.frombuffer(f.read(4 * nk), dtype="<f4")
        V = np.frombuffer(f.read(4 * nn), dtype="<f4")
        g = np.frombuffer(f.read(4 * 7 * nn), dtype="<f4")
        if len(knots) != nk or len(V) != nn or len(g) != 7 * nn or f.read(1):
            raise ValueError(f"{path}: truncated or oversized file")
    off = np.concatenate([[0], np.cumsum(dims)])
    grid_knots = [knots[off[d]:off[d + 1]].copy() for d in range(7)]
    shape = tuple(int(d) for d in dims)
    return HJICache(grid_knots, V.reshape(shape, order="F"), g.reshape((7,) + shape, order="F"))
