"""Oracle vs the reference's own code, live (only where oracle/_ref has been
built, i.e. in the container that has /root/reference): more randomised jobs
than the committed golden fixtures hold."""
import itertools

import numpy as np
import pytest

import cases

ref = pytest.importorskip("ref")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")

DT = [np.float64, np.float32, np.int32, np.int64]


@pytest.mark.parametrize("np_,unique,method,seed", [(1, 0, 1, 0), (1, 1, 1, 1), (2, 0, 1, 2), (2, 0, 3, 3),
                                                    (4, 0, 1, 4), (4, 1, 1, 5), (2, 1, 3, 6),
                                                    (8, 0, 1, 7), (8, 0, 3, 8), (8, 1, 1, 9)])
def test_oracle_matches_reference(O, np_, unique, method, seed):
    ids = [cases.random_ids(260 + 13 * r, 70, 1000 * seed + r, p_flag=0.0 if unique else 0.25) for r in range(np_)]
    combos = [(1, 8, 0, 0, 0)] + [(m, vn, d, o, t) for (m, vn), (d, o), t in
                                  itertools.product(((0, 1), (1, 2), (2, 2)), ((0, 0), (2, 3), (3, 2), (1, 1)), (0, 1))]
    # gs_bpr (src/gs_defs.h:37-42; xxt.upc:623 issues it on sint), every domain the reference instantiates it for
    combos += [(0, 1, 2, 4, 0), (1, 2, 3, 4, 1), (2, 2, 0, 4, 0), (0, 1, 1, 4, 1)]
    data = [[cases.values(ids[r].size * (1 if m == 0 else vn), DT[d], cases.OPS_BPR[o], 31 * k + r) for r in range(np_)]
            for k, (m, vn, d, o, t) in enumerate(combos)]
    mine = [[a.copy() for a in row] for row in data]
    maps, _ = ref.run(ids, combos, data, unique=unique, method=method, want_maps=True)
    orc = O.Oracle(ids, unique=unique, method=O.METHOD_ALLREDUCE if method == 3 else O.METHOD_PAIRWISE)
    for r in range(np_):
        for w in range(3):
            assert np.array_equal(orc.map(r, w), maps[r][w]), (r, w)
    for k, (m, vn, d, o, t) in enumerate(combos):
        us = [[row[i * ids[r].size:(i + 1) * ids[r].size] for i in range(vn)] if m == 2 else row
              for r, row in enumerate(mine[k])]
        orc.exec(us, m, vn, d, o, t)
        for r in range(np_):
            if DT[d] in (np.float64, np.float32) and o in (0, 1) and method == 3:
                np.testing.assert_array_max_ulp(mine[k][r], data[k][r], maxulp=4)
            else:
                assert np.array_equal(mine[k][r], data[k][r]), (k, r)
