"""Grid — host-side view of the device-resident mesh structures.

Mirrors the readonly attributes of the reference's `Grid` (ninpol/_interpolator/grid.pxd:128-187):
same names, same dtypes (int64 / float64), same shapes and -1 padding.  Arrays are copied from the
device the first time they are read and cached; nothing is computed on the host.
"""
import warnings

import numpy as np

from . import element_tables as et

_I = np.int64
_F = np.float64


class Grid:
    _SCALARS = ("dim", "n_elems", "n_points", "n_faces", "n_edges", "MX_ELEMENTS_PER_POINT", "MX_POINTS_PER_POINT",
                "MX_ELEMENTS_PER_FACE", "MX_FACES_PER_POINT")

    def __init__(self, ctx, tables, logging=False, build_edges=False):
        self._ctx = ctx
        self._cache = {}
        self.logging = logging
        self.build_edges = build_edges
        self.npoel, self.nfael, self.lnofa, self.lpofa, self.nedel, self.lpoed = tables
        self.are_elements_loaded = True
        self.are_coords_loaded = True
        self.are_structures_built = True
        self.are_centroids_calculated = True
        self.are_normals_calculated = True

    # shape of every exported array, from device-side scalars
    def _shape(self, name):
        s = self._ctx.scalar
        ne, np_, nf = s("n_elems"), s("n_points"), s("n_faces")
        table = {
            "point_coords": ((np_, 3), _F), "centroids": ((ne, 3), _F), "faces_centers": ((nf, 3), _F),
            "normal_faces": ((nf, 3), _F), "faces_areas": ((nf,), _F),
            "inpoel": ((ne, et.MAX_POINTS_PER_ELEMENT), _I), "element_types": ((ne,), _I),
            "esuel": ((ne, et.MAX_FACES_PER_ELEMENT), _I), "infael": ((ne, et.MAX_FACES_PER_ELEMENT), _I),
            "inpofa": ((nf, et.MAX_POINTS_PER_FACE), _I),
            "esup_ptr": ((np_ + 1,), _I), "fsup_ptr": ((np_ + 1,), _I), "psup_ptr": ((np_ + 1,), _I),
            "esuf_ptr": ((nf + 1,), _I), "boundary_faces": ((nf,), _I), "boundary_points": ((np_,), _I),
            "inedel": ((ne, et.MAX_EDGES_PER_ELEMENT), _I),
        }
        if name == "inpoed":
            return ((s("n_edges"), et.MAX_POINTS_PER_EDGE), _I)
        if name in table:
            return table[name]
        if name in ("esup", "fsup", "esuf", "psup"):
            return ((s("len_" + name),), _I)
        raise AttributeError(name)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in Grid._SCALARS:
            return self._ctx.scalar(name)
        if name in ("inpoed", "inedel") and not self.build_edges:
            return np.zeros((0, 0), dtype=_I)   # grid.pyx:132-133
        cache = self.__dict__["_cache"]
        if name not in cache:
            shape, dtype = self._shape(name)
            cache[name] = self._ctx.array(name, shape, dtype)
        return cache[name]

    def get_data(self):
        """Same dictionary as the reference's Grid.get_data (grid.pyx:583-658): raw arrays plus esup /
        psup / esuf / fsup as 2-D arrays padded with -1."""
        if not self.are_structures_built:
            raise ValueError("The structures have not been built.")
        data = {k: getattr(self, k) for k in Grid._SCALARS if k != "dim"}
        for k in ("point_coords", "centroids", "normal_faces", "faces_centers", "faces_areas", "boundary_faces",
                  "boundary_points", "inpoel", "element_types", "inpofa", "infael", "inpoed", "inedel"):
            data[k] = np.array(getattr(self, k))

        def pad(ptr, flat, width):
            n = len(ptr) - 1
            out = -np.ones((n, width), dtype=_I)
            cnt = np.diff(ptr)
            rows = np.repeat(np.arange(n), cnt)
            cols = np.arange(len(flat)) - np.repeat(ptr[:-1], cnt)
            out[rows, cols] = flat
            return out

        data["esup"] = pad(self.esup_ptr, self.esup, self.MX_ELEMENTS_PER_POINT)
        data["psup"] = pad(self.psup_ptr, self.psup, self.MX_POINTS_PER_POINT)
        data["esuf"] = pad(self.esuf_ptr, self.esuf, self.MX_ELEMENTS_PER_FACE)
        data["fsup"] = pad(self.fsup_ptr, self.fsup, self.MX_FACES_PER_POINT)
        return data
