"""The product's C++ box-mesh writer and the oracle's independent numpy generator produce the same
files in the reference's on-disk format (domain.tpp:43-224): integer files byte-identical, float files
to 1e-13 of the field's scale.  Also checks the conventions the reference relies on."""
import os
import numpy as np
import pytest
from oracle import meshgen

CASES = [(2, 4, 3, 1, 0.0), (2, (4, 6), 5, 2, 0.1), (3, (2, 4, 2), 3, 1, 0.05), (3, 4, 3, 4, 0.05), (3, 4, 2, 8, 0.07), (2, 16, 7, 1, 0.0)]


@pytest.mark.parametrize("dim,nel,N,nr,eps", CASES)
def test_generators_agree(prfdd, tmp_path, dim, nel, N, nr, eps):
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    meshgen.generate(d1, dim, nel, N, nranks=nr, eps=eps)
    prfdd.mesh_generate_box(d2, dim, nel, N, nr, eps)
    sub = "lx1_%d" % (N + 1)
    files = sorted(os.listdir(os.path.join(d1, sub)))
    assert files == sorted(os.listdir(os.path.join(d2, sub))) and len(files) == 13 * nr
    for fn in files:
        a = open(os.path.join(d1, sub, fn), "rb").read()
        b = open(os.path.join(d2, sub, fn), "rb").read()
        if fn.startswith(("glo_num", "node_degree", "size", "p_mask")):
            assert a == b, fn
        else:
            x, y = np.frombuffer(a), np.frombuffer(b)
            assert x.shape == y.shape
            assert np.abs(x - y).max() <= 1e-13 * max(1.0, np.abs(x).max()), fn


def test_conventions(tmp_path):
    d = str(tmp_path)
    recs7 = meshgen.generate(d, 3, 2, 3, nranks=2, eps=0.03)
    recs1 = meshgen.generate(d, 3, 2, 1, nranks=2, eps=0.03)
    for r7, r1 in zip(recs7, recs1):
        n = 4
        corners = [0, n - 1, n * (n - 1), n * n - 1]
        corners = corners + [c + n * n * (n - 1) for c in corners]
        # vertex ids identical at every ladder degree (subdomain.tpp:930-966)
        assert np.array_equal(r7["glo_num"][:, corners], r1["glo_num"])
        assert r7["glo_num"].min() >= 1
    allg = np.concatenate([r["glo_num"].ravel() for r in recs7])
    alld = np.concatenate([r["node_degree"].ravel() for r in recs7])
    ids, counts = np.unique(allg, return_counts=True)
    assert ids.size == 7 ** 3 and ids[0] == 1 and ids[-1] == 7 ** 3          # dense 1-based numbering
    lookup = dict(zip(ids.tolist(), counts.tolist()))
    assert all(lookup[g] == d_ for g, d_ in zip(allg.tolist(), alld.tolist()))   # node_degree = global multiplicity
