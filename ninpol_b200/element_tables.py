"""Element-type tables: local point numbering of edges and faces per element type, meshio ordering,
faces counter-clockwise seen from outside.  Same content and dict layout as the reference's
`ninpol/utils/point_ordering.yaml:6-53` (exposed there as `Interpolator.point_ordering`)."""
import numpy as np

NUM_ELEMENT_TYPES = 8        # ninpol_defines.pxd:5
MAX_POINTS_PER_ELEMENT = 8   # :2
MAX_FACES_PER_ELEMENT = 6    # :3
MAX_POINTS_PER_FACE = 4      # :4
MAX_EDGES_PER_ELEMENT = 12   # :6
MAX_POINTS_PER_EDGE = 2      # :8


def _el(tid, npts, edges, faces):
    return {"element_type": tid, "number_of_points": npts, "edges": edges, "faces": faces}


POINT_ORDERING = {"elements": {
    "vertex": _el(0, 1, [], []),
    "line": _el(1, 2, [[0, 1]], []),
    "triangle": _el(2, 3, [[0, 1], [1, 2], [2, 0]], []),
    "quad": _el(3, 4, [[0, 1], [1, 2], [2, 3], [3, 0]], []),
    "tetra": _el(4, 4, [[0, 1], [1, 2], [2, 0], [0, 3], [1, 3], [2, 3]],
                 [[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]]),
    "hexahedron": _el(5, 8, [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]],
                      [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]]),
    "wedge": _el(6, 6, [[0, 1], [1, 2], [2, 0], [3, 4], [4, 5], [5, 3], [0, 3], [1, 4], [2, 5]],
                 [[0, 2, 1], [3, 4, 5], [0, 1, 4, 3], [1, 2, 5, 4], [0, 3, 5, 2]]),
    "pyramid": _el(7, 5, [[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 4], [2, 4], [3, 4]],
                   [[0, 3, 2, 1], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]]),
}}

TYPES_PER_DIMENSION = {0: ["vertex"], 1: ["line"], 2: ["triangle", "quad"],
                       3: ["tetra", "hexahedron", "wedge", "pyramid"]}


def tables_for_dim(dim, point_ordering=POINT_ORDERING):
    """npoel, nfael, lnofa, lpofa, nedel, lpoed as Interpolator.process_mesh fills them
    (interpolator.pyx:274-330): int64, -1 everywhere except the types of the mesh dimension; in 2-D
    the "faces" are the edges (interpolator.pyx:296-298)."""
    T = NUM_ELEMENT_TYPES
    npoel = -np.ones(T, dtype=np.int64)
    nfael = -np.ones(T, dtype=np.int64)
    lnofa = -np.ones((T, MAX_FACES_PER_ELEMENT), dtype=np.int64)
    lpofa = -np.ones((T, MAX_FACES_PER_ELEMENT, MAX_POINTS_PER_FACE), dtype=np.int64)
    nedel = -np.ones(T, dtype=np.int64)
    lpoed = -np.ones((T, MAX_EDGES_PER_ELEMENT, MAX_POINTS_PER_EDGE), dtype=np.int64)
    faces_key = "edges" if dim == 2 else "faces"
    for name, el in point_ordering["elements"].items():
        t = el["element_type"]
        npoel[t] = el["number_of_points"]
        if name not in TYPES_PER_DIMENSION[dim]:
            continue
        fl = el.get(faces_key, [])
        nfael[t] = len(fl)
        for i, face in enumerate(fl):
            lnofa[t, i] = len(face)
            lpofa[t, i, :len(face)] = face
        nedel[t] = len(el.get("edges", []))
        for i, e in enumerate(el.get("edges", [])):
            lpoed[t, i, :] = e
    return npoel, nfael, lnofa, lpofa, nedel, lpoed
