"""Field output (Domain::output domain.tpp:373-524 writes Silo): the VTK writer of the C ABI on host arrays -- no GPU.
Cell connectivity must be the reference's low-order cells (every GLL cell one quad / hexahedron, same vertex order)."""
import ctypes as C
import numpy as np
import pytest
from oracle import meshgen


def _parse(path):
    tok = open(path).read().split("\n")
    i = next(k for k, l in enumerate(tok) if l.startswith("POINTS"))
    npts = int(tok[i].split()[1])
    pts = np.array([[float(v) for v in l.split()] for l in tok[i + 1:i + 1 + npts]])
    j = next(k for k, l in enumerate(tok) if l.startswith("CELLS"))
    nc = int(tok[j].split()[1])
    cells = [[int(v) for v in l.split()] for l in tok[j + 1:j + 1 + nc]]
    k = next(q for q, l in enumerate(tok) if l.startswith("CELL_TYPES"))
    types = [int(l) for l in tok[k + 1:k + 1 + nc]]
    fields = {}
    for q, l in enumerate(tok):
        if l.startswith("SCALARS"):
            fields[l.split()[1]] = np.array([float(v) for v in tok[q + 2:q + 2 + npts]])
    return pts, cells, types, fields


@pytest.mark.parametrize("dim,nel,N", [(2, 3, 4), (3, 2, 3)])
def test_vtk_writer(prfdd, tmp_path, dim, nel, N):
    L = prfdd.lib()
    rec = meshgen.generate(str(tmp_path), dim, nel, N, nranks=1, eps=0.05, write=False)[0]
    n = N + 1
    E = rec["E"]
    x, y, z = (np.ascontiguousarray(np.asarray(rec[k], dtype=np.float64).ravel()) for k in "xyz")
    u = np.sin(x) * np.cos(y) + z
    w = x * y
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    names = (C.c_char_p * 2)(b"u", b"w")
    fields = (C.c_void_p * 2)(u.ctypes.data, w.ctypes.data)
    path = str(tmp_path / "out.vtk")
    assert L.prfdd_write_vtk(path.encode(), C.c_int(dim), C.c_int(n), C.c_int(E), vp(x), vp(y), vp(z), C.c_int(2), names, fields) == 0
    pts, cells, types, f = _parse(path)
    assert pts.shape == (E * n ** dim, 3)
    assert np.array_equal(pts[:, 0], x) and np.array_equal(pts[:, 1], y) and np.array_equal(pts[:, 2], z if dim == 3 else 0 * x)
    assert len(cells) == E * N ** dim and set(types) == {9 if dim == 2 else 12}
    assert np.array_equal(f["u"], u) and np.array_equal(f["w"], w)          # %.17g round-trips doubles
    # connectivity: the reference's low-order cells (domain.tpp:416-445)
    c0 = cells[0]
    assert c0[0] == (4 if dim == 2 else 8)
    assert c0[1:5] == [0, 1, 1 + n, n]
    if dim == 3:
        assert c0[5:9] == [n * n, n * n + 1, n * n + 1 + n, n * n + n]
    last = cells[-1]
    assert max(last[1:]) == E * n ** dim - 1                                    # last cell ends at the last point of the last element
    # every cell of a mildly deformed mesh has positive volume in the reference's vertex order
    for c in cells[:: max(1, len(cells) // 50)]:
        p = pts[c[1:]]
        if dim == 2:
            a = 0.5 * sum(p[i, 0] * p[(i + 1) % 4, 1] - p[(i + 1) % 4, 0] * p[i, 1] for i in range(4))
            assert a > 0
        else:
            vol = np.linalg.det(np.stack([p[1] - p[0], p[3] - p[0], p[4] - p[0]]))
            assert vol > 0
    assert L.prfdd_write_vtk(b"/nonexistent_dir/x.vtk", C.c_int(dim), C.c_int(n), C.c_int(E), vp(x), vp(y), vp(z), C.c_int(0), None, None) == -2
    assert L.prfdd_write_vtk(path.encode(), C.c_int(4), C.c_int(n), C.c_int(E), vp(x), vp(y), vp(z), C.c_int(0), None, None) == -8


def test_vtk_writer_mixed_degrees(prfdd, tmp_path):
    """prfdd_write_vtk_mixed: elements of different degrees in one file (the region mesh of Subdomain::output, subdomain.tpp:4648-4791)"""
    L = prfdd.lib()
    dim = 3
    ns = np.array([4, 2, 3], np.int32)                        # points per side of the three elements
    pts = []
    for e, n in enumerate(ns):
        g = np.linspace(0.0, 1.0, n)
        X, Y, Z = np.meshgrid(g + 1.1 * e, g, g, indexing="ij")
        pts.append(np.stack([X.transpose(2, 1, 0).ravel(), Y.transpose(2, 1, 0).ravel(), Z.transpose(2, 1, 0).ravel()], 1))   # i fastest
    P = np.concatenate(pts)
    x, y, z = (np.ascontiguousarray(P[:, k]) for k in range(3))
    fld = x + 2 * y + 3 * z
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    names = (C.c_char_p * 1)(b"lin")
    fields = (C.c_void_p * 1)(fld.ctypes.data)
    path = str(tmp_path / "mixed.vtk")
    assert L.prfdd_write_vtk_mixed(path.encode(), C.c_int(dim), C.c_int(3), vp(ns), vp(x), vp(y), vp(z), C.c_int(1), names, fields) == 0
    got, cells, types, f = _parse(path)
    assert got.shape == (int((ns.astype(np.int64) ** 3).sum()), 3) and np.array_equal(got, P)
    assert len(cells) == int(((ns - 1).astype(np.int64) ** 3).sum()) and set(types) == {12}
    assert np.array_equal(f["lin"], fld)
    # first cell of the second element starts at that element's offset and uses ITS points per side
    c = cells[27]
    o, n = 64, 2
    assert c[1:] == [o, o + 1, o + 1 + n, o + n, o + n * n, o + n * n + 1, o + n * n + 1 + n, o + n * n + n]
    for c in cells:
        p = got[c[1:]]
        assert np.linalg.det(np.stack([p[1] - p[0], p[3] - p[0], p[4] - p[0]])) > 0
    assert L.prfdd_write_vtk_mixed(path.encode(), C.c_int(dim), C.c_int(3), vp(np.array([4, 1, 3], np.int32)), vp(x), vp(y), vp(z), C.c_int(0), None, None) == -8
