"""CPU tests of the host-side mirror of the reference interface: snapshot CSV / stats file layout
(F/SnapshotRW.jl), the sph_manager CLI surface (F/sph_manager.jl:13-98), IC generators, and the multi-GPU
host logic (partition arithmetic + unique-id distribution) under torch.distributed / gloo with world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import ROOT


def test_julia_float_text():
    from astrophysical_sph_b200.snapshot_rw import julia_float_str as f

    cases = {1.0: "1.0", 5e12: "5.0e12", 6.6743e-8: "6.6743e-8", 1e-4: "0.0001", 1e-5: "1.0e-5", 1e5: "100000.0",
             1e6: "1.0e6", 999999.9: "999999.9", 1.9891e33: "1.9891e33", 0.576: "0.576", -2.5e-7: "-2.5e-7",
             1.5e7: "1.5e7", 0.0: "0.0", 123456.789: "123456.789"}
    for x, s in cases.items():
        assert f(x) == s
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-30, 30, 200), [2.0**53, 1e-300, 1e300]]):
        assert float(f(x)) == x                      # round trip
        assert ("e" in f(x)) or ("." in f(x))        # the reader types by text shape (F/SnapshotRW.jl:147)


def test_snapshot_round_trip_and_layout(tmp_path):
    from astrophysical_sph_b200 import snapshot_rw as S

    rng = np.random.default_rng(1)
    N = 257
    pos = np.asfortranarray(rng.standard_normal((N, 3)) * 1e17)
    vel = np.asfortranarray(rng.standard_normal((N, 3)) * 1e5)
    K = rng.random(N) * 1e14
    consts = {"iterID": 7, "N": N, "Kh": 50, "Kgr": 20, "t": 1.5e11, "tEnd": 5e12, "M": 1.9891e33, "R": 1.07e17,
              "alpha": 1.0, "beta": 2.0, "G": 6.6743e-8, "theta": 0.576, "m": 1.9891e33 / N, "gamma": 5 / 3}
    rlin = np.linspace(0, 1.6e17, 11)
    rho = rng.random(11) * 1e-18
    path = S.write_snapshot("7", "gaussian_sphere", pos, vel, K=K, constants=consts, rlin=rlin, rho_radial=rho,
                            root=str(tmp_path))
    assert path.endswith(os.path.join("snapshots", "gaussian_sphere", "bin", "7snap.csv"))
    lines = open(path).read().splitlines()
    assert lines[0] == "type,x,y,z,vx,vy,vz,K,rlin,rho_radial,constants"          # F/SnapshotRW.jl:37-49
    assert len(lines) == 1 + N + 3
    assert lines[1].startswith("particle,") and lines[1].count(",") == 10 and lines[1].endswith(",,,")
    assert lines[N + 1].startswith("rlin,,,,,,,,") and lines[N + 2].startswith("rho_radial,,,,,,,,,")
    assert lines[N + 3].startswith("constants,,,,,,,,,,") and "iterID=7;" in lines[N + 3] and "alpha=1.0" in lines[N + 3]
    back = S.read_snapshot(path)
    assert np.array_equal(back["pos"], pos) and np.array_equal(back["vel"], vel) and np.array_equal(back["K"], K)
    assert back["pos"].flags["F_CONTIGUOUS"]
    assert np.array_equal(back["rlin"], rlin) and np.array_equal(back["rho_radial"], rho)
    assert back["constants"] == consts
    assert isinstance(back["constants"]["N"], int) and isinstance(back["constants"]["alpha"], float)
    # isothermal snapshot: empty K column -> None
    p2 = S.write_snapshot("1", "boss_bodenheimer", pos, vel, constants={"iterID": 1, "cs": 1.0e4}, root=str(tmp_path))
    assert S.read_snapshot(p2)["K"] is None


def test_stats_mmap_layout(tmp_path):
    from astrophysical_sph_b200 import snapshot_rw as S

    fn = str(tmp_path / "snapshots" / "x" / "stats")
    arr, _ = S.open_or_create_stats_mmap(fn)
    assert os.path.getsize(fn) == 100000 * 10 * 8                                  # F/SnapshotRW.jl:171-176
    row = np.arange(10, dtype=float) + 0.5
    S.update_stats_row(arr, 3, row)
    arr.flush()
    raw = np.fromfile(fn, dtype=np.float64)
    for f in range(10):                                                           # column-major: offset 8((f)*100000 + (r-1))
        assert raw[f * 100000 + 2] == row[f]
    assert np.array_equal(S.get_stats_up_to(arr, 3)[2], row)
    with pytest.raises(AssertionError):
        S.update_stats_row(arr, 0, row)
    arr2, _ = S.open_or_create_stats_mmap(fn)                                      # reopen r+
    assert arr2[2, 4] == row[4]


def test_cli_surface_matches_reference():
    from astrophysical_sph_b200 import sph_manager as M

    a = M.parse_command_line(["--run", "--EOS", "isothermal", "--ic_type", "boss_bodenheimer"])
    assert (a.snapID, a.snapInterval, a.keepSnaps, a.showPlots, a.kwargs, a.generate) == (1, 10, True, True, "", False)
    a = M.parse_command_line(["--generate", "--EOS", "polytropic", "--ic_type", "gaussian_sphere", "--kwargs",
                              "N=5000,R=5.38552341e16,flag=true,name=abc", "--keepSnaps", "false", "--snapID", "4"])
    assert a.generate and not a.run and a.keepSnaps is False and a.snapID == 4
    kw = M.parse_kwargs(a.kwargs)
    assert kw == {"N": 5000, "R": 5.38552341e16, "flag": True, "name": "abc"} and isinstance(kw["N"], int)
    with pytest.raises(SystemExit):
        M.parse_command_line(["--run", "--ic_type", "x"])                          # --EOS is required (:22-25)
    with pytest.raises(SystemExit):
        M.parse_command_line(["--run", "--EOS", "isothermal", "--ic-type", "x"])  # the README's hyphenated spelling is not a flag


def test_generate_writes_reference_layout(tmp_path):
    from astrophysical_sph_b200 import snapshot_rw as S
    from astrophysical_sph_b200 import sph_manager as M

    M.main(["--generate", "--EOS", "isothermal", "--ic_type", "boss_bodenheimer", "--kwargs", "N=500,T=10",
            "--root", str(tmp_path)])
    snap = S.read_snapshot(S.snapshot_path(1, "boss_bodenheimer", str(tmp_path)))
    c = snap["constants"]
    assert set(c) == {"iterID", "N", "Kh", "Kgr", "t", "tEnd", "M", "R", "alpha", "beta", "G", "theta", "m", "cs", "U"}
    assert c["N"] == 500 and snap["pos"].shape == (500, 3) and snap["K"] is None
    M.main(["--generate", "--EOS", "polytropic", "--ic_type", "gaussian_sphere", "--kwargs", "N=400", "--root", str(tmp_path)])
    snap = S.read_snapshot(S.snapshot_path(1, "gaussian_sphere", str(tmp_path)))
    assert "gamma" in snap["constants"] and "cs" not in snap["constants"] and snap["K"].shape == (400,)
    with pytest.raises(ValueError):
        M.main(["--generate", "--EOS", "isothermal", "--ic_type", "nope", "--root", str(tmp_path)])


def test_target_ranges_tile_the_sorted_order():
    from astrophysical_sph_b200.parallel import chunk_size, padded_size, target_range, upload_slice

    for N in (64, 1000, 1679, 1680, 1681, 100_003, 1_000_000, 16_000_000):
        for P in (1, 2, 3, 4, 5, 6, 7, 8, 9, 16):
            c = chunk_size(N, P)
            assert c % 128 == 0 and c * P >= N and padded_size(N, P) == c * P
            assert padded_size(N, P) <= (N + 127) // 128 * 128 + 128 * 16      # NS_alloc of sph_create
            r = [target_range(N, P, k) for k in range(P)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[i][1] == r[i + 1][0] for i in range(P - 1))
            assert all(a % 128 == 0 for a, _ in r if a < N)                    # list tiles never straddle ranks
            assert max(b - a for a, b in r) <= c
            u = [upload_slice(N, P, k) for k in range(P)]
            assert u[0][0] == 0 and u[-1][1] == N and all(u[i][1] == u[i + 1][0] for i in range(P - 1))
    with pytest.raises(ValueError):
        target_range(1000, 17, 0)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from astrophysical_sph_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
uid = parallel.share_unique_id(dist, lambda: bytes(range(128)))
assert uid == bytes(range(128))
N = 100003
t0, t1 = parallel.target_range(N, dist.get_world_size(), dist.get_rank())
box = [None, None]
dist.all_gather_object(box, (t0, t1))
assert box[0][0] == 0 and box[0][1] == box[1][0] and box[1][1] == N
dist.barrier()
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
"""


def test_two_rank_host_logic_over_gloo(tmp_path):
    """world_size-2 run of the multi-GPU host plumbing on CPU: rank 0's 128-byte communicator id reaches rank 1 and
    the two target ranges tile [0, N)."""
    import socket

    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
