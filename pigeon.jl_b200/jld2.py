"""Minimal JLD2 (HDF5 dialect) reader / writer for the reference's HJI cache file.

The reference keeps the cache in `BicycleCAvoid.jld2` with exactly three objects (src/HJI_computation.jl:39-64, deps/build.jl:1-4):

    grid_knots :: NTuple{7,Vector{Float32}}      V_raw :: Array{Float32,7}      ∇V_raw :: Array{Float32,N} (7 components fastest)

JLD2 0.1 writes an HDF5 file with a 512-byte text header, a version-2 superblock at byte 512 whose BASE ADDRESS is 512 (every address in the
file is relative to it), version-2 object headers ("OHDR", Jenkins lookup3 checksums), link messages in the root group, contiguous (or, for
tiny objects, compact) dataset layouts, and Julia types encoded as committed datatypes under the group `_types`; a tuple of arrays is a scalar
dataset whose compound datatype has one 8-byte object reference per element, each pointing at the array's own dataset.

`read_jld2(path)` walks exactly that subset of the HDF5 file format specification (superblock v0/v2/v3, object headers v1/v2 with
continuation blocks, link / dataspace / datatype / layout messages, shared (committed) datatypes, fixed-point / floating-point / compound /
reference classes, contiguous and compact layouts) and returns {name: numpy array | tuple of arrays}.  `write_jld2(path, objects)` emits the
same structure.  No HDF5 library and no Julia exist in this image: the pair is validated against each other and against the format
specification's structure (signatures, checksums, address arithmetic) in tests/test_host_cpu.py; a byte-level check against a file written by
JLD2.jl itself needs the machine julia/make_golden.jl runs on (the recipe also copies such a file next to the golden vectors)."""
import struct

import numpy as np

HDR_LEN = 512
SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


# ---- Jenkins lookup3 (hashlittle), the checksum of version-2 HDF5 metadata --------------------------------------------------------------------
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
    a = b = c = (0xDEADBEEF + len(data) + initval) & 0xFFFFFFFF
    n, off = len(data), 0
    M = 0xFFFFFFFF
    while n > 12:
        a = (a + int.from_bytes(data[off:off + 4], "little")) & M
        b = (b + int.from_bytes(data[off + 4:off + 8], "little")) & M
        c = (c + int.from_bytes(data[off + 8:off + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        off += 12; n -= 12
    if n == 0:
        return c
    tail = data[off:off + n] + b"\x00" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


# ---- reader -------------------------------------------------------------------------------------------------------------------------------
class _File:
    def __init__(self, buf):
        self.b = buf
        pos = None
        for cand in (0, 512, 1024, 2048):
            if buf[cand:cand + 8] == SIG:
                pos = cand
                break
        if pos is None:
            raise ValueError("no HDF5 superblock signature at 0, 512, 1024 or 2048")
        self.sb = pos
        ver = buf[pos + 8]
        if ver in (2, 3):
            so, sl = buf[pos + 9], buf[pos + 10]
            if so != 8 or sl != 8:
                raise ValueError("only 8-byte offsets and lengths are supported")
            self.base, _ext, self.eof, self.root = struct.unpack_from("<QQQQ", buf, pos + 12)
            if lookup3(buf[pos:pos + 44]) != struct.unpack_from("<I", buf, pos + 44)[0]:
                raise ValueError("superblock checksum mismatch")
            self.root_is_ohdr = True
        elif ver in (0, 1):
            so, sl = buf[pos + 13], buf[pos + 14]
            if so != 8 or sl != 8:
                raise ValueError("only 8-byte offsets and lengths are supported")
            o = pos + 24 + (4 if ver == 1 else 0)
            self.base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", buf, o)
            # root group symbol table entry: link name offset (8), object header address (8), ...
            self.root = struct.unpack_from("<Q", buf, o + 32 + 8)[0]
            self.root_is_ohdr = True
        else:
            raise ValueError(f"unsupported superblock version {ver}")

    def at(self, addr):
        return self.base + addr


def _messages(f, addr):
    """Yields (type, flags, bytes) of every message of the object header at file address `addr` (versions 1 and 2, continuation blocks followed)."""
    b, p = f.b, f.at(addr)
    out = []
    if b[p:p + 4] == b"OHDR":
        if b[p + 4] != 2:
            raise ValueError("unsupported object header version")
        flags = b[p + 5]
        q = p + 6
        if flags & 0x20:
            q += 16
        if flags & 0x10:
            q += 4
        szlen = 1 << (flags & 3)
        chunk = int.from_bytes(b[q:q + szlen], "little")
        q += szlen
        if lookup3(b[p:q + chunk]) != struct.unpack_from("<I", b, q + chunk)[0]:
            raise ValueError(f"object header checksum mismatch at {addr}")
        blocks = [(q, q + chunk)]
        track = bool(flags & 0x04)
        while blocks:
            s, e = blocks.pop(0)
            while s + 4 <= e:
                mt, ms, mf = b[s], struct.unpack_from("<H", b, s + 1)[0], b[s + 3]
                s += 4 + (2 if track else 0)
                body = b[s:s + ms]
                s += ms
                if mt == 0x10:      # continuation: offset, length of an "OCHK" block
                    co, cl = struct.unpack_from("<QQ", body, 0)
                    cp = f.at(co)
                    if b[cp:cp + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    if lookup3(b[cp:cp + cl - 4]) != struct.unpack_from("<I", b, cp + cl - 4)[0]:
                        raise ValueError("continuation block checksum mismatch")
                    blocks.append((cp + 4, cp + cl - 4))
                elif mt != 0:
                    out.append((mt, mf, bytes(body)))
        return out
    # version 1 header
    if b[p] != 1:
        raise ValueError(f"no object header at {addr}")
    nmsg, _refs, hsize = struct.unpack_from("<HII", b, p + 2)
    blocks = [(p + 16, p + 16 + hsize)]
    while blocks and nmsg > 0:
        s, e = blocks.pop(0)
        while s + 8 <= e and nmsg > 0:
            mt, ms, mf = struct.unpack_from("<HHB", b, s)
            body = b[s + 8:s + 8 + ms]
            s += 8 + ms
            nmsg -= 1
            if mt == 0x10:
                co, cl = struct.unpack_from("<QQ", body, 0)
                blocks.append((f.at(co), f.at(co) + cl))
            elif mt != 0:
                out.append((mt, mf, bytes(body)))
    return out


def _parse_datatype(f, body, flags=0):
    """-> ("float" | "int" | "ref" | "compound" | "opaque", numpy dtype or member list, size)"""
    if flags & 0x02:        # shared message: version, type, address of the committed datatype's object header
        ver = body[0]
        addr = struct.unpack_from("<Q", body, 2 if ver >= 2 else 8)[0]
        for mt, mf, mb in _messages(f, addr):
            if mt == 0x03:
                return _parse_datatype(f, mb, mf & ~0x02)
        raise ValueError("committed datatype without a datatype message")
    cls, ver = body[0] & 0x0F, body[0] >> 4
    bits = body[1] | (body[2] << 8) | (body[3] << 16)
    size = struct.unpack_from("<I", body, 4)[0]
    if cls == 0:
        return "int", np.dtype(("<" if not bits & 1 else ">") + ("i" if bits & 8 else "u") + str(size)), size
    if cls == 1:
        return "float", np.dtype(("<" if not bits & 1 else ">") + "f" + str(size)), size
    if cls == 7:
        return "ref", np.dtype("<u8"), size
    if cls == 6:
        n = bits & 0xFFFF
        members, q = [], 8
        for _ in range(n):
            e = body.index(b"\x00", q)
            name = body[q:e].decode()
            q = e + 1
            if ver < 3:
                q = (q + 7) & ~7 if (q - 8) % 8 else q          # names padded to 8 bytes in versions 1 and 2
                off = struct.unpack_from("<I", body, q)[0]
                q += 4
                if ver == 1:
                    q += 28
            else:
                nb = max(1, (size.bit_length() + 7) // 8)
                off = int.from_bytes(body[q:q + nb], "little")
                q += nb
            sub = _parse_datatype(f, body[q:])
            q += _datatype_len(body[q:])
            members.append((name, off, sub))
        return "compound", members, size
    return "opaque", None, size


def _datatype_len(body):
    cls, ver = body[0] & 0x0F, body[0] >> 4
    if cls == 0:
        return 12
    if cls == 1:
        return 20
    if cls == 7:
        return 8
    if cls == 6:
        n = body[1] | (body[2] << 8)
        size = struct.unpack_from("<I", body, 4)[0]
        q = 8
        for _ in range(n):
            e = body.index(b"\x00", q)
            q = e + 1
            if ver < 3:
                q = (q + 7) & ~7 if (q - 8) % 8 else q
                q += 4 + (28 if ver == 1 else 0)
            else:
                q += max(1, (size.bit_length() + 7) // 8)
            q += _datatype_len(body[q:])
        return q
    raise ValueError(f"unsupported datatype class {cls}")


def _read_dataset(f, addr, depth=0):
    dims, dt, layout = None, None, None
    for mt, mf, body in _messages(f, addr):
        if mt == 0x01:
            ver, rank, fl = body[0], body[1], body[2]
            q = 8 if ver == 1 else 4
            dims = struct.unpack_from("<%dQ" % rank, body, q) if rank else ()
        elif mt == 0x03:
            dt = _parse_datatype(f, body, mf)
        elif mt == 0x08:
            ver, cls = body[0], body[1]
            if ver not in (3, 4):
                raise ValueError("unsupported data layout version")
            if cls == 0:
                n = struct.unpack_from("<H", body, 2)[0]
                layout = ("compact", body[4:4 + n])
            elif cls == 1:
                a, n = struct.unpack_from("<QQ", body, 2)
                layout = ("contiguous", a, n)
            else:
                raise ValueError("chunked datasets are not supported")
    if dims is None or dt is None or layout is None:
        raise ValueError(f"object at {addr} is not a dataset")
    raw = layout[1] if layout[0] == "compact" else (b"" if layout[1] == UNDEF else f.b[f.at(layout[1]):f.at(layout[1]) + layout[2]])
    kind, np_dt, size = dt
    count = int(np.prod(dims)) if dims else 1
    if kind in ("float", "int"):
        a = np.frombuffer(raw, dtype=np_dt, count=count)
        # HDF5 dataspaces list the slowest dimension first; JLD2 writes Julia's dimensions reversed, so this is the Julia array in Fortran order
        return a.reshape(tuple(reversed(dims)), order="F").copy() if dims else a[0]
    if kind == "compound" and all(m[2][0] == "ref" for m in np_dt):
        if depth > 2:
            raise ValueError("reference chain too deep")
        return tuple(_read_dataset(f, struct.unpack_from("<Q", raw, off)[0], depth + 1) for _, off, _ in np_dt)
    if kind == "ref":
        return tuple(_read_dataset(f, int(r), depth + 1) for r in np.frombuffer(raw, dtype="<u8", count=count))
    raise ValueError(f"unsupported datatype ({kind}) for the dataset at {addr}")


def read_jld2(path, names=None):
    """{name: array | tuple of arrays} of the datasets linked from the root group (sub-groups such as `_types` are skipped)."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    out = {}
    for mt, mf, body in _messages(f, f.root):
        if mt != 0x06:
            continue
        fl = body[1]
        q = 2
        ltype = 0
        if fl & 0x08:
            ltype = body[q]; q += 1
        if fl & 0x04:
            q += 8
        if fl & 0x10:
            q += 1
        ln = 1 << (fl & 3)
        nlen = int.from_bytes(body[q:q + ln], "little"); q += ln
        name = body[q:q + nlen].decode("utf-8"); q += nlen
        if ltype != 0 or (names is not None and name not in names):
            continue
        addr = struct.unpack_from("<Q", body, q)[0]
        try:
            out[name] = _read_dataset(f, addr)
        except ValueError:
            if names is not None:
                raise
    return out


# ---- writer -------------------------------------------------------------------------------------------------------------------------------
def _msg(mtype, body, flags=0):
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _ohdr(messages):
    body = b"".join(messages)
    head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body))          # flags: 4-byte chunk size, no times, no creation order
    blk = head + body
    return blk + struct.pack("<I", lookup3(blk))


def _dt_float32():
    # class 1 version 1; bits: little-endian, IEEE padding, mantissa normalisation 2 (implied), sign bit 31; properties: offset 0, precision 32,
    # exponent at 23 (8 bits), mantissa at 0 (23 bits), bias 127
    return bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_ref():
    return bytes([0x17, 0x00, 0x00, 0x00]) + struct.pack("<I", 8)


def _dt_tuple_of_refs(n):
    body = bytes([0x36, n & 0xFF, (n >> 8) & 0xFF, 0x00]) + struct.pack("<I", 8 * n)      # class 6 (compound) version 3
    for i in range(n):
        body += str(i + 1).encode() + b"\x00" + bytes([8 * i]) + _dt_ref()                 # member "1".."n", 1-byte offset (size < 256)
    return body


def _dataspace(dims):
    if not dims:
        return bytes([2, 0, 0, 0])                                                         # version 2, rank 0, scalar
    return bytes([2, len(dims), 0, 1]) + b"".join(struct.pack("<Q", d) for d in dims)


def _link(name, addr):
    nb = name.encode("utf-8")
    return bytes([1, 0x10 | 0x00, 1]) + bytes([len(nb)]) + nb + struct.pack("<Q", addr)      # version 1, flags: charset present (UTF-8), 1-byte name length


class _Writer:
    def __init__(self):
        self.buf = bytearray()        # everything after the 512-byte header; addresses are offsets into this buffer (base address 512)

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf += b"\x00"
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr_f_order_bytes, julia_dims, dtype_msg):
        data = self.alloc(arr_f_order_bytes)
        msgs = [_msg(0x01, _dataspace(tuple(reversed(julia_dims)))), dtype_msg,
                _msg(0x08, bytes([3, 1]) + struct.pack("<QQ", data, len(arr_f_order_bytes)))]
        return self.alloc(_ohdr(msgs))


def write_jld2(path, objects):
    """objects: {name: float32 numpy array (Julia index order) | tuple of float32 vectors}.  Writes the JLD2 structure described above."""
    w = _Writer()
    w.alloc(b"\x00" * 48)                                  # the superblock's place (filled in last)
    links = []
    f32 = _msg(0x03, _dt_float32())
    types_links = []
    for name, obj in objects.items():
        if isinstance(obj, (tuple, list)):
            refs = [w.dataset(np.asarray(v, dtype="<f4").tobytes(), (len(v),), f32) for v in obj]
            # the tuple's compound datatype is committed under _types (JLD2 shares it by address) and referenced with a shared datatype message
            tdt = w.alloc(_ohdr([_msg(0x03, _dt_tuple_of_refs(len(refs)))]))
            types_links.append(_msg(0x06, _link("%08d" % (len(types_links) + 1), tdt)))
            shared = _msg(0x03, bytes([3, 2]) + struct.pack("<Q", tdt), flags=0x02)          # shared message v3, type 2 = committed, address
            data = b"".join(struct.pack("<Q", r) for r in refs)
            da = w.alloc(data)
            links.append(_msg(0x06, _link(name, w.alloc(_ohdr([_msg(0x01, _dataspace(())), shared, _msg(0x08, bytes([3, 1]) + struct.pack("<QQ", da, len(data)))])))))
        else:
            a = np.asarray(obj, dtype="<f4")
            links.append(_msg(0x06, _link(name, w.dataset(np.asfortranarray(a).tobytes(order="F"), a.shape, f32))))
    if types_links:
        links.append(_msg(0x06, _link("_types", w.alloc(_ohdr(types_links)))))
    root = w.alloc(_ohdr(links))
    eof = len(w.buf)
    sb = SIG + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", HDR_LEN, UNDEF, eof, root)
    sb += struct.pack("<I", lookup3(sb))
    w.buf[0:48] = sb
    header = b"Julia data file (HDF5), version 0.2.0 (written by pigeon.jl_b200/jld2.py)"
    with open(path, "wb") as fh:
        fh.write(header + b"\x00" * (HDR_LEN - len(header)))
        fh.write(bytes(w.buf))


# ---- the HJI cache ------------------------------------------------------------------------------------------------------------------------
def load_hji_jld2(path):
    """HJICache(fname) of the reference (src/HJI_computation.jl:47-57): grid_knots, V_raw, ∇V_raw -> (knots, V (n1..n7), gradV (7, n1..n7))."""
    d = read_jld2(path, names=("grid_knots", "V_raw", "∇V_raw"))
    missing = [k for k in ("grid_knots", "V_raw", "∇V_raw") if k not in d]
    if missing:
        raise ValueError(f"{path}: objects {missing} not found")
    knots = [np.asarray(k, dtype=np.float32) for k in d["grid_knots"]]
    dims = tuple(len(k) for k in knots)
    V = np.asarray(d["V_raw"], dtype=np.float32)
    g = np.asarray(d["∇V_raw"], dtype=np.float32)
    if len(knots) != 7 or V.shape != dims or g.size != 7 * V.size:
        raise ValueError(f"{path}: inconsistent shapes: knots {dims}, V_raw {V.shape}, ∇V_raw {g.shape}")
    # save() writes Array(reinterpret(Float32, ∇V.coefs)): size (7*n1, n2, ..., n7) on Julia 1.0, (7, n1, ..., n7) on later versions — same memory
    return knots, V, g.reshape((7,) + dims, order="F")


def save_hji_jld2(path, knots, V, gradV):
    """save(fname, cache) (src/HJI_computation.jl:59-64): ∇V_raw with the size Julia 1.0's reinterpret gives, (7*n1, n2, ..., n7)."""
    V = np.asarray(V, dtype=np.float32)
    g = np.asarray(gradV, dtype=np.float32).reshape((7 * V.shape[0],) + V.shape[1:], order="F")
    write_jld2(path, {"grid_knots": tuple(np.asarray(k, dtype=np.float32) for k in knots), "V_raw": V, "∇V_raw": g})
