"""Minimal stand-in for the `meshio` package — TEST INFRASTRUCTURE ONLY.

The compiled reference under oracle/_ref imports `meshio` at module import time
(reference ninpol/_interpolator/interpolator.pyx:8) but only ever calls `meshio.read` when it is
given a file name (interpolator.pyx:188).  meshio is not installed in this image and there is no
network, so this module provides just the two containers the reference touches when it is handed a
mesh object (`load_mesh(mesh_obj=...)`): `.points`, `.cells[i].type/.data`, `.cell_data`,
`.cell_data_dict`, `.point_data`, `.cells_dict`.
"""
import numpy as np


class CellBlock:
    def __init__(self, cell_type, data):
        self.type = cell_type
        self.data = np.asarray(data)

    def __len__(self):
        return len(self.data)

    def __repr__(self):
        return f"<shim CellBlock {self.type} x{len(self.data)}>"


class Mesh:
    def __init__(self, points, cells, point_data=None, cell_data=None):
        self.points = np.asarray(points)
        blocks = []
        if isinstance(cells, dict):
            cells = list(cells.items())
        for c in cells:
            if isinstance(c, CellBlock) or (hasattr(c, "type") and hasattr(c, "data")):
                blocks.append(CellBlock(c.type, c.data))
            else:
                blocks.append(CellBlock(c[0], c[1]))
        self.cells = blocks
        self.point_data = {} if point_data is None else point_data
        self.cell_data = {} if cell_data is None else cell_data

    @property
    def cells_dict(self):
        out = {}
        for blk in self.cells:
            out.setdefault(blk.type, []).append(blk.data)
        return {k: np.concatenate(v) for k, v in out.items()}

    @property
    def cell_data_dict(self):
        out = {}
        for name, per_block in self.cell_data.items():
            by_type = {}
            for values, blk in zip(per_block, self.cells):
                by_type.setdefault(blk.type, []).append(np.asarray(values))
            out[name] = {t: np.concatenate(v) for t, v in by_type.items()}
        return out


def read(filename, file_format=None):
    raise RuntimeError("meshio shim: reading mesh files is not available (meshio is not installed)")
