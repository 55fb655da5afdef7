""".world / .msg path files (reference test/path/*.world: one `key: v1, v2, ...` line per field of the `path` message) -> TrajectoryTube.

The reference reads these through ROS (`path` messages, src/ros_integration.jl:13-19); the file form is the YAML dump of that message with the
float arrays written as comma-separated lists.  Keys: s_m, posE_m, posN_m, psi_rad, k_1pm, grade_rad, edgeL_m, edgeR_m, UxDes_mps, AxDes_mps2,
isOpen.  No YAML dependency: the format is line-oriented."""
import struct

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


# ---- serialised `path` messages (reference test/path/*.msg, written by test/path/world2pathmsg.py through the ROS 1 serialiser) -------------
# Wire layout, little-endian, as found in the nine fixtures (the message definition itself lives in an un-vendored ROS package):
#   std_msgs/Header: uint32 seq, uint32 stamp.secs, uint32 stamp.nsecs, string frame_id (uint32 length + bytes);
#   8 bytes, zero in every fixture (kept opaque as 'reserved');
#   ten float64[] (uint32 count + values) in the order of MSG_KEYS;
#   isOpen as one little-endian 8-byte integer.
MSG_KEYS = WORLD_KEYS          # s, E, N, psi, k, grade, edge L, edge R, Ux des, Ax des: same order as the .world lines


def read_msg(path):
    """Returns {key: float64 array} with the .world key names, + 'isOpen', 'frame_id', 'seq', 'stamp' (secs, nsecs), 'reserved' (bytes).
    Raises ValueError on truncated, ragged or trailing data."""
    with open(path, "rb") as f:
        b = f.read()
    off = 0

    def take(n):
        nonlocal off
        if off + n > len(b):
            raise ValueError(f"{path}: truncated at byte {off} (need {n}, have {len(b) - off})")
        off += n
        return b[off - n:off]

    seq, secs, nsecs, flen = struct.unpack("<IIII", take(16))
    out = {"seq": seq, "stamp": (secs, nsecs), "frame_id": take(flen).decode("utf-8", "replace"), "reserved": bytes(take(8))}
    for k in MSG_KEYS:
        n, = struct.unpack("<I", take(4))
        out[k] = np.frombuffer(take(8 * n), dtype="<f8").astype(np.float64)
    out["isOpen"] = int(struct.unpack("<q", take(8))[0])
    if off != len(b):
        raise ValueError(f"{path}: {len(b) - off} trailing bytes")
    n = len(out["s_m"])
    if n < 2 or any(len(out[k]) != n for k in MSG_KEYS):
        raise ValueError(f"{path}: fields must all have the same length >= 2")
    return out


def write_msg(path, fields, is_open=1, frame_id="", seq=0, stamp=(0, 0), reserved=b"\0" * 8):
    fid = frame_id.encode("utf-8")
    with open(path, "wb") as f:
        f.write(struct.pack("<IIII", seq, stamp[0], stamp[1], len(fid)) + fid + bytes(reserved))
        for k in MSG_KEYS:
            a = np.ascontiguousarray(fields[k], dtype="<f8")
            f.write(struct.pack("<I", a.size) + a.tobytes())
        f.write(struct.pack("<q", int(is_open)))


def trajectory_from_msg(path):
    """TrajectoryTube(p::path) (src/ros_integration.jl:13-16) of a serialised path message."""
    return TrajectoryTube.from_path(read_msg(path))
