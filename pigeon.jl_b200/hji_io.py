"""On-disk form of the HJI cache for hosts without Julia.

The reference keeps `grid_knots`, `V_raw` and `∇V_raw` in a JLD2 file (`HJICache(fname)` / `save`, src/HJI_computation.jl:39-64; the file is
downloaded by deps/build.jl).  JLD2 is an HDF5 dialect with Julia type encodings; this image has neither Julia nor an HDF5 library, and no
copy of the real file to validate a parser against, so the JLD2 container itself is not read here.  Instead `julia/export_hji_cache.jl` (run
once where Julia + JLD2.jl are installed) dumps exactly those three objects, in Julia's memory order, into the flat little-endian file below,
which this module reads and writes:

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
    dims = np.array([len(k) for k in cache.grid_knots], dtype="<i4")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(dims.tobytes())
        f.write(np.concatenate(cache.grid_knots).astype("<f4").tobytes())
        f.write(np.ascontiguousarray(cache.V.ravel(order="F"), dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(cache.gradV.ravel(order="F"), dtype="<f4").tobytes())


def load_hji_cache(path):
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
