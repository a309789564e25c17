"""Host conveniences around the hot path: file meshes through `meshio.read`, the
inputs-only pickle cache (interpolator.pyx:93-165,180-191,244-252), `load_face_data` (:456-499),
`get_data` / `get_dict` (:511-547). meshio is not in this image, so a stand-in module whose `read`
unpickles a mesh is installed for the duration of the test."""
import os
import pickle
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def fake_meshio(monkeypatch):
    mod = types.ModuleType("meshio")
    mod.calls = []

    def read(filename):
        mod.calls.append(filename)
        with open(filename, "rb") as f:
            return pickle.load(f)
    mod.read = read
    monkeypatch.setitem(sys.modules, "meshio", mod)
    return mod


def test_file_mesh_is_cached_and_the_cache_reproduces_the_weights(tmp_path, fake_meshio):
    import ninpol_b200
    from ninpol_b200 import meshgen
    mesh = meshgen.make_case("mixed", 8, a=2, b=4)
    path = str(tmp_path / "box.msh")
    with open(path, "wb") as f:
        pickle.dump(mesh, f)

    I = ninpol_b200.Interpolator()
    I.CACHE_PATH = str(tmp_path)
    assert I.is_cached(path) is None and I.is_cached("") is None
    I.load_mesh(filename=path)
    assert fake_meshio.calls == [path]
    cached = I.is_cached(path)
    assert cached is not None and os.path.dirname(cached) == str(tmp_path)
    assert os.path.basename(cached) == "box" + hex(os.path.getsize(path)) + ".pkl"
    first = {m: I.interpolate("u", m) for m in ("idw", "ls", "gls")}

    J = ninpol_b200.Interpolator()
    J.CACHE_PATH = str(tmp_path)
    J.load_mesh(filename=path)                       # second load: from the cache, meshio.read not called
    assert fake_meshio.calls == [path]
    assert J.variable_to_index == I.variable_to_index
    for m, (W, nv) in first.items():
        W2, nv2 = J.interpolate("u", m)
        assert np.array_equal(W.indptr, W2.indptr) and np.array_equal(W.indices, W2.indices)
        assert np.array_equal(W.data, W2.data, equal_nan=True) and np.array_equal(nv, nv2, equal_nan=True)

    K = ninpol_b200.Interpolator()
    K.load_mesh(mesh_obj=mesh)                       # and the object path gives the same thing
    W3, _ = K.interpolate("u", "gls")
    assert np.array_equal(first["gls"][0].data, W3.data, equal_nan=True)


def test_missing_meshio_is_a_clear_error(tmp_path, monkeypatch):
    import ninpol_b200
    monkeypatch.setitem(sys.modules, "meshio", None)     # import meshio -> ImportError
    p = tmp_path / "x.vtk"
    p.write_bytes(b"0")
    I = ninpol_b200.Interpolator()
    I.CACHE_PATH = str(tmp_path)
    with pytest.raises(ImportError, match="meshio"):
        I.load_mesh(filename=str(p))


def test_get_data_get_dict_and_face_data():
    import ninpol_b200
    from ninpol_b200 import meshgen
    mesh = meshgen.make_case("tet", 5)
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    n_e, n_p, n_f = I.grid.n_elems, I.grid.n_points, I.grid.n_faces
    u = np.asarray(mesh.cell_data["u"][0])
    idx = np.array([0, 3, n_e - 1])
    assert np.array_equal(I.get_data("cells", idx, "u"), u[idx])
    flag = np.asarray(mesh.point_data["neumann_flag_u"], dtype=float)
    assert np.array_equal(I.get_data("points", np.arange(n_p), "neumann_flag_u"), flag)
    with pytest.raises(ValueError, match="not found"):
        I.get_data("cells", idx, "nope")
    d = I.get_dict()
    assert set(d) >= {"point_ordering", "variable_to_index", "cells_data", "points_data",
                      "cells_data_dimensions", "points_data_dimensions"}
    vi = d["variable_to_index"]["cells"]
    assert np.array_equal(d["cells_data"][vi["u"]][:n_e], u)
    perm = np.asarray(mesh.cell_data["permeability"][0]).reshape(n_e, 9)
    assert np.array_equal(d["cells_data"][vi["permeability"]][:9 * n_e], perm.reshape(-1))
    assert d["cells_data_dimensions"][vi["permeability"]] == 9
    assert np.array_equal(d["cells_data"][vi["diff_mag"]][:n_e], (1 - 3.0 / (perm[:, 0] + perm[:, 4] + perm[:, 8])) ** 2)

    # face data in grid order, then through a user connectivity that lists the faces in reverse
    vals = np.arange(n_f, dtype=float).reshape(n_f, 1)
    I.load_face_data({"flux": vals})
    assert I.variable_to_index["faces"]["flux"] == 0
    assert np.array_equal(np.asarray(I.faces_data)[0], vals[:, 0])
    inpofa = np.asarray(I.grid.inpofa)
    I.load_face_data({"flux": vals[::-1].copy()}, face_connectivity=inpofa[::-1].copy())
    # user row i describes grid face n_f-1-i; data[face_to_grid] (interpolator.pyx:499) undoes the reversal
    got = np.asarray(I.faces_data)[0]
    assert np.array_equal(got, vals[:, 0])


def test_cache_is_private_and_not_served_to_another_file_of_equal_name_and_size(tmp_path, fake_meshio):
    """The reference keys its cache on basename + file size under a world-writable directory; here the cache lives in a
    per-user 0700 directory and records the source's path, size and mtime, so a different file that merely has the same
    name and size is rebuilt instead of silently loading the other mesh."""
    import stat
    import ninpol_b200
    from ninpol_b200 import meshgen
    I = ninpol_b200.Interpolator()
    st = os.stat(I.CACHE_PATH)
    assert stat.S_IMODE(st.st_mode) & 0o077 == 0 and st.st_uid == os.getuid()
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir()
    b.mkdir()
    mesh_a = meshgen.make_case("tet", 4)
    mesh_b = meshgen.make_case("tet", 4, seed=7)            # same sizes, other coordinates and fields
    pa, pb = str(a / "box.msh"), str(b / "box.msh")
    for path, mesh in ((pa, mesh_a), (pb, mesh_b)):
        with open(path, "wb") as f:
            pickle.dump(mesh, f)
    assert os.path.getsize(pa) == os.path.getsize(pb)
    I.CACHE_PATH = str(tmp_path)
    I.load_mesh(filename=pa)
    Wa, _ = I.interpolate("u", "idw")
    J = ninpol_b200.Interpolator()
    J.CACHE_PATH = str(tmp_path)
    J.load_mesh(filename=pb)                                 # same cache name: must NOT be served mesh a
    assert fake_meshio.calls == [pa, pb]
    Wb, _ = J.interpolate("u", "idw")
    assert not np.array_equal(Wa.data, Wb.data)
    assert np.array_equal(np.asarray(J.grid.point_coords), np.asarray(mesh_b.points))
