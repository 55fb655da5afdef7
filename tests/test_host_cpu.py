"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/pigeon_b200.h declares (no compute calls
without a GPU), the static QP analysis is correct (emulated factor/solve against a dense solve), the Python mirror of the
reference API, the synthetic workload generators and the world_size-2 sharding/gather plumbing (gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def p():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("pgn_build", os.path.join(ROOT, "pigeon.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    import pigeon.jl_b200 as pkg
    return pkg


def test_library_exports_every_declared_symbol(p):
    hdr = open(os.path.join(ROOT, "include", "pigeon_b200.h")).read()
    declared = re.findall(r"PGN_API\s+(?:const char\*|int)\s+(pgn_\w+)\s*\(", hdr)
    assert len(declared) >= 35
    assert sorted(declared) == sorted(p.SYMBOLS)
    lib = p.load()
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", p.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pgn_\w+)", nm))
    assert set(declared) <= exported
    assert not any(s.startswith("orc_") for s in re.findall(r" T (\w+)", nm))      # the oracle is not linked into the product


def test_no_cpu_fallback_without_gpu(p):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(p.PigeonError, match="no CUDA device|CUDA"):
        p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), p.straight_trajectory(30.0, 5.0), 4)


def test_product_does_not_reference_the_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "pigeon.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in src and "liboracle" not in src and '#include "../../oracle' not in src, f


@pytest.mark.parametrize("args", ["0 10 20 0", "0 10 20 1", "0 5 10 0", "1 10 20 0", "1 10 20 1", "0 1 0 0", "1 3 2 0", "0 10 20 0 w8", "1 10 20 0 w8"])
def test_static_qp_tables_factor_and_solve(args, tmp_path):
    exe = str(tmp_path / "structure_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "structure_check.cpp"),
                           os.path.join(ROOT, "pigeon.jl_b200", "csrc", "pgn_structure.cpp")])
    env = dict(os.environ)
    if args.endswith(" w8"):      # the programs of the 256-thread ADMM builds (8 warps)
        args = args[:-3]; env["PGN_CHECK_NWARPS"] = "8"
    r = subprocess.run([exe] + args.split(), capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(x.split("=") for x in r.stdout.splitlines()[0].split())
    kind, Ns, Nl = (int(a) for a in args.split()[:3])
    m = o.Mpc(kind, N_short=Ns, N_long=Nl)
    assert (int(kv["n"]), int(kv["m"]), int(kv["nnzA"])) == (m.n, m.m, m.nnzA)      # same canonical QP as the oracle
    assert int(kv["rz_bad"]) == 0 and int(kv["rz_prog"]) == 1      # the Ruiz norm program covers every adjacency list of these QPs


def test_nested_dissection_cuts_solve_depth(tmp_path):
    exe = str(tmp_path / "structure_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "structure_check.cpp"),
                           os.path.join(ROOT, "pigeon.jl_b200", "csrc", "pgn_structure.cpp")])
    nd = dict(x.split("=") for x in subprocess.run([exe, "0", "10", "20", "0"], capture_output=True, text=True).stdout.splitlines()[0].split())
    md = dict(x.split("=") for x in subprocess.run([exe, "0", "10", "20", "1"], capture_output=True, text=True).stdout.splitlines()[0].split())
    assert int(nd["nlev"]) * 3 < int(md["nlev"])
    assert int(nd["nnzL"]) < 1.6 * int(md["nnzL"])


def test_python_mirror_parameters_match_oracle(p):
    x1 = p.X1()
    vp = o.x1()
    assert [x1[k] for k in o.VP_NAMES] == list(vp)
    assert list(p.CoupledControlParams().values()) == list(o.control_params_default(o.MPC_COUPLED))
    assert list(p.DecoupledControlParams().values()) == list(o.control_params_default(o.MPC_DECOUPLED))
    assert p.CoupledControlParams(Q_e=3.0)["Q_e"] == 3.0
    with pytest.raises(TypeError):
        p.CoupledControlParams(nope=1)


def test_trajectory_tube_from_path_matches_oracle(p):
    w = np.load(os.path.join(ROOT, "tests", "golden", "world_vail.npz"))
    tr = p.TrajectoryTube.from_path(w)
    assert np.array_equal(tr.t, o.invcumtrapz(w["UxDes_mps"], w["s_m"]))
    assert len(tr) == 1000 and np.all(tr.phi == 0)
    st = p.straight_trajectory(30.0, 5.0)
    assert list(st.t) == [0.0, 6.0] and list(st.N) == [0.0, 30.0] and list(st.edge_L) == [4.0, 4.0]
    with pytest.raises(ValueError):
        p.TrajectoryTube([0, 1], [0, 1, 2], [1, 1], [0, 0], [0, 0], [0, 1], [0, 0], [0, 0])


def test_synthetic_workload_is_deterministic_and_feasible(p):
    a = p.synthetic.synthetic_trajectories(n_traj=4, n_nodes=200)
    b = p.synthetic.synthetic_trajectories(n_traj=4, n_nodes=200)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert np.all(np.diff(a["t"], axis=1) > 0) and np.all(np.abs(a["kappa"]) <= 0.07 + 1e-12)
    assert np.all(a["V"] >= 4 - 1e-9) and np.all(a["V"] <= 12 + 1e-9)
    assert np.all(a["V"] ** 2 * np.abs(a["kappa"]) <= 0.5 * 0.92 * 9.80665 + 1e-9)
    tid, state, control, t0 = p.synthetic.synthetic_batch(a, 16)
    # the sampled states are near their trajectory: oracle path_coordinates gives |e| ~ 0.3 m
    for i in range(16):
        tr = o.Trajectory(**{k: a[k][tid[i]] for k in o.TRAJ_FIELDS})
        s, e, _ = tr.path_coordinates(state[i, 0], state[i, 1])
        assert abs(e) < 1.5
    knots, V, g = p.synthetic.analytic_hji_grid((3, 3, 3, 3, 3, 3, 3))
    assert V.shape == (3,) * 7 and g.shape == (7,) + (3,) * 7 and V.dtype == np.float32


def test_shard_range_partitions_exactly(p):
    from pigeon.jl_b200 import sharding
    for total in (1, 7, 8, 1024, 65536, 1000003):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, total, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import pigeon.jl_b200  # noqa: F401
    from pigeon.jl_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(total * 3, dtype=np.float64).reshape(total, 3)
    iters = (np.arange(total) % 7).astype(np.int32)
    lo, hi = sharding.shard_range(total, world, rank)
    g = sharding.gather_batch(full[lo:hi], total, dist).numpy()
    gi = sharding.gather_batch(iters[lo:hi], total, dist).numpy()
    q.put((rank, bool(np.array_equal(g, full)), bool(np.array_equal(gi, iters))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gather_world_size_2_gloo(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)


def test_world_reader_round_trip_and_fixtures(p, tmp_path):
    """.world path files (reference test/path/*.world) -> TrajectoryTube: the line-oriented reader round-trips a written file bit for bit,
    agrees with the committed fixtures (converted by tests/golden/make_world_fixtures.py with a YAML parser) where the reference tree
    is present, and rejects ragged files."""
    golden = os.path.join(ROOT, "tests", "golden")
    names = sorted(f[6:-4] for f in os.listdir(golden) if f.startswith("world_") and f.endswith(".npz"))
    assert len(names) == 8
    for name in names:
        w = np.load(os.path.join(golden, f"world_{name}.npz"))
        f = str(tmp_path / f"{name}.world")
        p.write_world(f, w, is_open=int(w["isOpen"]))
        r = p.read_world(f)
        for k in p.world.WORLD_KEYS:
            assert np.array_equal(r[k], w[k]), (name, k)
        assert r["isOpen"] == int(w["isOpen"])
        t = p.trajectory_from_world(f)
        assert len(t) == len(w["s_m"]) and t.t[0] == 0.0 and np.all(np.diff(t.t) > 0)
        ref = os.path.join("/root/reference/test/path", name + ".world")
        if os.path.exists(ref):                                   # build container only
            rr = p.read_world(ref)
            for k in p.world.WORLD_KEYS:
                assert np.array_equal(rr[k], w[k]), (name, k)
    bad = str(tmp_path / "bad.world")
    w = dict(np.load(os.path.join(golden, "world_curvy.npz")))
    w["psi_rad"] = w["psi_rad"][:-1]
    p.write_world(bad, w)
    with pytest.raises(ValueError):
        p.read_world(bad)


def test_msg_reader_on_reference_bytes(p, tmp_path):
    """Serialised `path` messages (reference test/path/*.msg): the reader decodes the reference's own bytes (two fixtures travel raw in
    tests/golden/msg_raw.npz), agrees bit for bit with the .world twin, re-serialises to the identical bytes, loads the fixture that has no
    .world twin (variable_speed.msg, 28 nodes, varying speed) into a TrajectoryTube, and rejects truncated files."""
    golden = os.path.join(ROOT, "tests", "golden")
    raw = np.load(os.path.join(golden, "msg_raw.npz"))
    assert sorted(raw.files) == ["curvy", "variable_speed"]
    for name in raw.files:
        f = str(tmp_path / f"{name}.msg")
        raw[name].tofile(f)
        m = p.read_msg(f)
        g = str(tmp_path / f"{name}_again.msg")
        p.write_msg(g, m, is_open=m["isOpen"], frame_id=m["frame_id"], seq=m["seq"], stamp=m["stamp"], reserved=m["reserved"])
        assert open(g, "rb").read() == raw[name].tobytes()
    m = p.read_msg(str(tmp_path / "curvy.msg"))
    w = np.load(os.path.join(golden, "world_curvy.npz"))
    assert m["frame_id"] == "curvy" and m["isOpen"] == int(w["isOpen"])
    for k in p.world.WORLD_KEYS:
        assert np.array_equal(m[k], w[k]), k
    v = p.read_msg(str(tmp_path / "variable_speed.msg"))
    assert len(v["s_m"]) == 28 and v["frame_id"] == "world" and np.ptp(v["UxDes_mps"]) > 0.1
    t = p.trajectory_from_msg(str(tmp_path / "variable_speed.msg"))
    assert len(t) == 28 and t.t[0] == 0.0 and np.all(np.diff(t.t) > 0)
    assert np.allclose(np.diff(t.t), 2 * np.diff(v["s_m"]) / (v["UxDes_mps"][:-1] + v["UxDes_mps"][1:]), rtol=1e-13)
    bad = str(tmp_path / "bad.msg")
    raw["variable_speed"][:-9].tofile(bad)
    with pytest.raises(ValueError):
        p.read_msg(bad)
    np.concatenate([raw["variable_speed"], np.zeros(3, np.uint8)]).tofile(bad)
    with pytest.raises(ValueError):
        p.read_msg(bad)
    ref = "/root/reference/test/path"
    if os.path.isdir(ref):                                        # build container only: every .msg equals its .world twin
        for f in sorted(os.listdir(ref)):
            if f.endswith(".msg") and os.path.exists(os.path.join(ref, f[:-4] + ".world")):
                mm, ww = p.read_msg(os.path.join(ref, f)), p.read_world(os.path.join(ref, f[:-4] + ".world"))
                assert all(np.array_equal(mm[k], ww[k]) for k in p.world.WORLD_KEYS) and mm["isOpen"] == ww["isOpen"], f


def test_hji_cache_file_round_trip(p, tmp_path):
    """The flat PGNHJI1 file (julia/export_hji_cache.jl writes it from the reference's JLD2 objects) round-trips an HJICache bit for bit, in the
    memory order pgn_set_hji_cache expects, and rejects truncated files."""
    knots, V, gV = p.synthetic.analytic_hji_grid((5, 4, 5, 4, 3, 4, 3))
    c = p.HJICache(knots, V, gV)
    f = str(tmp_path / "cache.pgnhji")
    p.save_hji_cache(f, c)
    assert os.path.getsize(f) == 8 + 28 + 4 * (sum(len(k) for k in knots) + V.size + gV.size)
    r = p.load_hji_cache(f)
    assert all(np.array_equal(a, b) for a, b in zip(r.grid_knots, c.grid_knots))
    assert np.array_equal(r.V, c.V) and np.array_equal(r.gradV, c.gradV)
    raw = open(f, "rb").read()
    # Julia memory order: V[i1, ...] with dimension 1 fastest; gradient components fastest
    v0 = np.frombuffer(raw[36 + 4 * sum(len(k) for k in knots):][:8], dtype="<f4")
    assert v0[0] == V[0, 0, 0, 0, 0, 0, 0] and v0[1] == V[1, 0, 0, 0, 0, 0, 0]
    g0 = np.frombuffer(raw[36 + 4 * (sum(len(k) for k in knots) + V.size):][:32], dtype="<f4")
    assert np.array_equal(g0[:7], gV[:, 0, 0, 0, 0, 0, 0, 0]) and g0[7] == gV[0, 1, 0, 0, 0, 0, 0, 0]
    open(f, "wb").write(raw[:-5])
    with pytest.raises(ValueError):
        p.load_hji_cache(f)


def test_synthetic_batch_vectorised_equals_the_loop(p):
    """The vectorised workload generator (a million vehicles in seconds) draws the same random numbers in the same order as the per-vehicle
    loop it replaced: every seed gives bit-identical batches, so round-1 and round-2 numbers are measured on the same vehicles."""
    tr = p.synthetic.synthetic_trajectories(n_traj=8, n_nodes=200)
    for B, seed in ((257, p.synthetic.SEED + 17), (64, 5), (3, 1)):
        a, b = p.synthetic.synthetic_batch(tr, B, seed=seed), p.synthetic._synthetic_batch_loop(tr, B, seed=seed)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_jld2_hji_cache_round_trip_and_structure(p, tmp_path):
    """BicycleCAvoid.jld2 (src/HJI_computation.jl:39-64): the minimal JLD2 / HDF5 writer and reader agree with each other bit for bit, and the
    file has the structure JLD2 0.1 writes: 512-byte text header, version-2 superblock at 512 with base address 512 and a valid lookup3
    checksum, "OHDR" object headers, root links `grid_knots`, `V_raw`, `∇V_raw`, `_types`; ∇V_raw with the size Julia 1.0's reinterpret gives."""
    import struct
    from pigeon.jl_b200 import jld2
    knots, V, gV = p.synthetic.analytic_hji_grid((5, 4, 5, 4, 3, 4, 3))
    c = p.HJICache(knots, V, gV)
    f = str(tmp_path / "BicycleCAvoid.jld2")
    p.save_hji_cache(f, c)
    r = p.load_hji_cache(f)
    assert all(np.array_equal(a, b) for a, b in zip(r.grid_knots, c.grid_knots))
    assert np.array_equal(r.V, c.V) and np.array_equal(r.gradV, c.gradV)
    raw = open(f, "rb").read()
    assert raw.startswith(b"Julia data file (HDF5)") and raw[512:520] == b"\x89HDF\r\n\x1a\n" and raw[520] == 2
    base, ext, eof, root = struct.unpack_from("<QQQQ", raw, 524)
    assert base == 512 and eof == len(raw) - 512 and raw[512 + root:512 + root + 4] == b"OHDR"
    assert jld2.lookup3(raw[512:556]) == struct.unpack_from("<I", raw, 556)[0]
    # lookup3 known answers (Bob Jenkins' lookup3.c self-test): hashlittle("", 0) = 0xdeadbeef, hashlittle("Four score and seven years ago", 0) = 0x17770551
    assert jld2.lookup3(b"") == 0xDEADBEEF and jld2.lookup3(b"Four score and seven years ago") == 0x17770551 and jld2.lookup3(b"Four score and seven years ago", 1) == 0xCD628161
    d = jld2.read_jld2(f)
    assert set(d) == {"grid_knots", "V_raw", "∇V_raw"} and isinstance(d["grid_knots"], tuple) and len(d["grid_knots"]) == 7
    assert d["∇V_raw"].shape == (7 * 5, 4, 5, 4, 3, 4, 3) and d["V_raw"].shape == (5, 4, 5, 4, 3, 4, 3)
    # Julia memory order inside the file: V_raw with dimension 1 fastest
    pos = raw.find(np.asfortranarray(V).tobytes(order="F")[:64])
    assert pos > 512
    # a flipped byte in an object header is caught by its checksum; a truncated file is rejected
    bad = bytearray(raw); bad[512 + root + 12] ^= 0x40
    open(f, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        p.load_hji_cache(f)
    open(f, "wb").write(raw[:600])
    with pytest.raises((ValueError, struct.error, IndexError)):
        p.load_hji_cache(f)
    # the later-Julia shape of ∇V_raw, (7, n1, ..., n7), is accepted as well (same memory)
    jld2.write_jld2(f, {"grid_knots": tuple(knots), "V_raw": V, "∇V_raw": gV})
    r2 = p.load_hji_cache(f)
    assert np.array_equal(r2.gradV, c.gradV)


@pytest.mark.parametrize("dim", [1, 2, 7, 12, 36, 59])
def test_two_pivot_sweep_formulas_give_minus_inverse(dim):
    """The dense-tail sweep of k_admm (pgn_admm_kernel.inc, factor()) restated on a full symmetric matrix: block pivot {p, p+1} with
    B^-1 = [e1 e2; e2 e3]; ordinary rows use w = U_i B^-1, the two pivot rows count as zero and use w = -(row of B^-1); every element becomes
    S_ik - w1 c1_k - w2 c2_k and the two pivot columns receive w1, w2; an odd dimension ends with one scalar pivot.  On a quasi-definite
    matrix (positive and negative pivots, as the Schur complement of the KKT top separators) the result must be -S^-1."""
    rng = np.random.default_rng(dim)
    npos = (dim + 1) // 2
    G = rng.standard_normal((dim, dim))
    S = np.zeros((dim, dim))
    S[:npos, :npos] = G[:npos, :npos] @ G[:npos, :npos].T + npos * np.eye(npos)
    S[npos:, npos:] = -(G[npos:, npos:] @ G[npos:, npos:].T + (dim - npos + 1) * np.eye(dim - npos))
    S[npos:, :npos] = 0.3 * G[npos:, :npos]; S[:npos, npos:] = S[npos:, :npos].T
    perm = rng.permutation(dim)                      # interleave the signs as the elimination order does
    S = S[np.ix_(perm, perm)]
    ref = -np.linalg.inv(S)
    A = S.copy()
    p = 0
    while p + 1 < dim:
        c1, c2 = A[:, p].copy(), A[:, p + 1].copy()
        a, b, c = c1[p], c1[p + 1], c2[p + 1]
        dinv = 1.0 / (a * c - b * b)
        e1, e2, e3 = c * dinv, -(b * dinv), a * dinv
        new = np.empty_like(A)
        for i in range(dim):
            w1, w2 = c1[i] * e1 + c2[i] * e2, c1[i] * e2 + c2[i] * e3
            base = A[i].copy()
            if i == p: w1, w2, base = -e1, -e2, np.zeros(dim)
            if i == p + 1: w1, w2, base = -e2, -e3, np.zeros(dim)
            row = base - w1 * c1 - w2 * c2
            row[p], row[p + 1] = w1, w2
            new[i] = row
        A = new
        p += 2
    if p < dim:
        cc = A[:, p].copy()
        dinv = 1.0 / cc[p]
        new = A - np.outer(cc * dinv, cc)
        new[:, p] = cc * dinv; new[p, :] = cc * dinv; new[p, p] = -dinv
        A = new
    assert np.allclose(A, A.T, rtol=0, atol=1e-9 * np.abs(ref).max())
    assert np.allclose(A, ref, rtol=1e-9, atol=1e-10 * np.abs(ref).max())
