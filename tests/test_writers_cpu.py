"""Output formats: the HDF5 result file (PyCFD_ML_accelerated.py:517-544 /
bfs_ml_accelerated.py:722-752, written there with h5py) and the centerline text file ("bfs code given by
sir.py":359-384), produced here without h5py.  Checked against two files the reference itself committed:
tests/golden/ref_bfs_coarse_Re400_10x10.h5 and ref_bfs_Re400_centerline.dat (copied by tests/golden/make_golden.py).
No GPU: the writer methods are called on a stand-in object that carries the attributes they read."""
import os
import types

import numpy as np

from srcfd import h5lite, solver as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_H5 = os.path.join(GOLD, "ref_bfs_coarse_Re400_10x10.h5")
REF_DAT = os.path.join(GOLD, "ref_bfs_Re400_centerline.dat")


def _walk(path):
    """{object path: (sorted message types of its v1 object header, layout class or None)} + superblock version."""
    with open(path, "rb") as f:
        r = h5lite._Reader(f.read())
    out = {}

    def visit(addr, name):
        msgs = r.messages(addr)
        types_ = sorted({m[0] for m in msgs if m[0] not in (0x0000, 0x0010, 0x0012)})    # nil / continuation / mtime are padding
        layout = dtype_raw = None
        stab = None
        for mtype, body, msize in msgs:
            if mtype == 0x0008:
                layout = (r.b[body], r.b[body + 1])                      # (version, class)
            elif mtype == 0x0003:
                dtype_raw = bytes(r.b[body:body + 8])                    # class+version, bit fields, size
            elif mtype == 0x0011:
                stab = (r.u(body, 8), r.u(body + 8, 8))
        out[name] = (types_, layout, dtype_raw)
        if stab:
            for child, caddr in r.iter_group(*stab):
                visit(caddr, f"{name}/{child}")

    visit(r.u(r.root_entry + 8, 8), "")
    return r.b[8], out


def _stand_in(ref_group, case_type):
    """What _save_results_hdf5 / _save_centerline_data read from a CFDSolver, filled from a reference result group."""
    nx, ny = int(ref_group.attrs["nx"]), int(ref_group.attrs["ny"])
    Var = np.zeros((3, nx + 2, ny + 2))
    for k, n in enumerate("uvp"):
        Var[k, 1:-1, 1:-1] = np.asarray(ref_group[n].data).reshape(ny, nx).T
    return types.SimpleNamespace(
        Var=Var, case_name=str(ref_group.attrs["case_name"]), case_type=case_type,
        mesh=types.SimpleNamespace(nx=nx, ny=ny, lx=float(ref_group.attrs.get("lx", 1.0)), ly=float(ref_group.attrs.get("ly", 1.0))),
        fluid=types.SimpleNamespace(Re=int(ref_group.attrs["reynolds_number"])), step_height=float(ref_group.attrs.get("step_height", 1.0)))


def test_h5_round_trip(tmp_path):
    root = h5lite.Group()
    root.attrs = {"title": "round trip", "version": 3}
    g = h5lite.Group()
    g.attrs = {"case_name": "lid driven cavity", "reynolds_number": 100.0, "nx": 7, "flag": np.int32(5)}
    rng = np.random.default_rng(0)
    g["u"] = rng.standard_normal(35)
    g["m"] = rng.standard_normal((5, 7)).astype(np.float32)
    g["i"] = np.arange(12, dtype=np.int64).reshape(3, 4)
    sub = h5lite.Group(); sub["w"] = np.array([1.5, -2.5]); g["sub"] = sub
    root["Re100.0_mesh7x5"] = g                                # hazard H10: float Re in the group name
    root["empty"] = h5lite.Group()
    path = str(tmp_path / "rt.h5")
    h5lite.write_h5(path, root)
    back = h5lite.read_h5(path)
    assert set(back.keys()) == {"Re100.0_mesh7x5", "empty"} and back.attrs["title"] == "round trip" and back.attrs["version"] == 3
    b = back["Re100.0_mesh7x5"]
    assert b.attrs["case_name"] == "lid driven cavity" and b.attrs["reynolds_number"] == 100.0 and b.attrs["nx"] == 7 and b.attrs["flag"] == 5
    for k in ("u", "m", "i"):
        assert np.array_equal(b[k].data, g[k]) and b[k].data.dtype == np.asarray(g[k]).dtype and b[k].data.shape == np.asarray(g[k]).shape
    assert np.array_equal(b["sub"]["w"].data, [1.5, -2.5])
    # appending a second group keeps the first (the reference opens its file in 'a' mode and replaces one group)
    back["second"] = h5lite.Group(); back["second"]["z"] = np.zeros(3)
    h5lite.write_h5(path, back)
    again = h5lite.read_h5(path)
    assert set(again.keys()) == {"Re100.0_mesh7x5", "empty", "second"} and np.array_equal(again["Re100.0_mesh7x5"]["u"].data, g["u"])


def test_result_file_has_the_reference_layout(tmp_path):
    """_save_results_hdf5 on the fields of a committed reference result: same group name, datasets, dtypes, shapes,
    attribute names / kinds / values, and -- object by object -- the same kinds of HDF5 header messages, dataset layout
    class and float64 datatype encoding as the file h5py wrote."""
    ref = h5lite.read_h5(REF_H5)
    (gname, rg), = ref.items()
    assert gname == "Re400_mesh10x10"
    fake = _stand_in(rg, "BFS")
    path = str(tmp_path / "out.h5")
    S.CFDSolver._save_results_hdf5(fake, path, S.CFDSolver._group_name(fake))
    mine = h5lite.read_h5(path)
    assert list(mine.keys()) == [gname]
    mg = mine[gname]
    assert sorted(mg.keys()) == sorted(rg.keys()) == ["p", "u", "v", "x", "y"]
    for k in rg:
        assert mg[k].data.dtype == rg[k].data.dtype == np.float64 and mg[k].data.shape == rg[k].data.shape == (100,)
        assert np.array_equal(mg[k].data, rg[k].data), k       # x, y: the same linspace/meshgrid; u, v, p: what went in
    assert list(mg.attrs.keys()) == list(rg.attrs.keys())      # same names in the same order
    for k, v in rg.attrs.items():
        assert mg.attrs[k] == v and np.asarray(mg.attrs[k]).dtype.kind == np.asarray(v).dtype.kind, k
    sb_r, walk_r = _walk(REF_H5)
    sb_m, walk_m = _walk(path)
    assert sb_m == sb_r == 0                                   # superblock version 0
    assert walk_m.keys() == walk_r.keys()
    for name in walk_r:
        tr, lr, dr = walk_r[name]
        tm, lm, dm = walk_m[name]
        assert tm == tr, (name, tm, tr)                        # dataspace / datatype / fill / layout / attribute / symbol table
        assert lm == lr and dm == dr, (name, lm, lr, dm, dr)   # contiguous layout v3, IEEE float64 little-endian
    # the LDC variant (PyCFD_ML_accelerated.py:523-544) has no domain/step attributes
    fake2 = _stand_in(rg, None); fake2.case_name = "lid driven cavity"
    S.CFDSolver._save_results_hdf5(fake2, path, "Re400_mesh10x10")
    assert list(h5lite.read_h5(path)[gname].attrs.keys()) == ["case_name", "reynolds_number", "nx", "ny", "total_points"]


def test_centerline_file_matches_the_reference_text(tmp_path):
    """_save_centerline_data: the reference's own bfs_Re400_centerline.dat is reproduced byte for byte from fields that
    round to its six decimals (header lines, tab layout, .6f columns)."""
    with open(REF_DAT) as f:
        want = f.read()
    rows = np.loadtxt(REF_DAT)
    nx = ny = 10
    Var = np.zeros((3, nx + 2, ny + 2))
    Var[0, nx // 2, 1:-1] = rows[:, 1]
    Var[1, 1:-1, ny // 2] = rows[:, 3]
    fake = types.SimpleNamespace(Var=Var, mesh=types.SimpleNamespace(nx=nx, ny=ny, lx=10.0, ly=3.0), fluid=types.SimpleNamespace(Re=400))
    path = str(tmp_path / "c.dat")
    S.CFDSolver._save_centerline_data(fake, path)
    with open(path) as f:
        got = f.read()
    assert got == want
    # ragged case (nx != ny): the shorter column is padded with two tabs, as in the reference loop
    fake.mesh = types.SimpleNamespace(nx=4, ny=6, lx=1.0, ly=1.0)
    fake.Var = np.arange(3 * 6 * 8, dtype=float).reshape(3, 6, 8)
    S.CFDSolver._save_centerline_data(fake, path)
    lines = open(path).read().splitlines()
    assert len(lines) == 4 + 6 and lines[4 + 5].endswith("\t") and lines[4 + 5].count("\t") == 2 and lines[4].count("\t") == 3
