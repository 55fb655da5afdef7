""".world path files (reference test/path/*.world: one `key: v1, v2, ...` line per field of the `path` message) -> TrajectoryTube.

The reference reads these through ROS (`path` messages, src/ros_integration.jl:13-19); the file form is the YAML dump of that message with the
float arrays written as comma-separated lists.  Keys: s_m, posE_m, posN_m, psi_rad, k_1pm, grade_rad, edgeL_m, edgeR_m, UxDes_mps, AxDes_mps2,
isOpen.  No YAML dependency: the format is line-oriented."""
import numpy as np

from .mpc import TrajectoryTube

WORLD_KEYS = ["s_m", "posE_m", "posN_m", "psi_rad", "k_1pm", "grade_rad", "edgeL_m", "edgeR_m", "UxDes_mps", "AxDes_mps2"]


def read_world(path):
    """Returns {key: float64 array} (+ 'isOpen': int).  Raises ValueError on ragged or missing fields."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            key, sep, val = line.partition(":")
            if not sep:
                raise ValueError(f"{path}: malformed line {line[:40]!r}")
            key, val = key.strip(), val.strip().strip("[]")
            if key == "isOpen":
                out[key] = int(float(val))
            else:
                out[key] = np.array([float(x) for x in val.split(",") if x.strip()], dtype=np.float64)
    missing = [k for k in WORLD_KEYS if k not in out]
    if missing:
        raise ValueError(f"{path}: missing fields {missing}")
    n = len(out["s_m"])
    if n < 2 or any(len(out[k]) != n for k in WORLD_KEYS):
        raise ValueError(f"{path}: fields must all have the same length >= 2")
    return out


def write_world(path, fields, is_open=1):
    with open(path, "w") as f:
        for k in WORLD_KEYS:
            f.write(f"{k}: " + ", ".join(repr(float(x)) for x in np.asarray(fields[k], dtype=np.float64)) + "\n")
        f.write(f"isOpen: {int(is_open)}\n")


def trajectory_from_world(path):
    """TrajectoryTube(p::path) (src/ros_integration.jl:13-16) of a .world file."""
    return TrajectoryTube.from_path(read_world(path))
