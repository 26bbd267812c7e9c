"""Reverse Cuthill-McKee order and Matrix::permute against the reference.

  * live against find_new_order_RCM of the UNMODIFIED reference (oracle/_ref), on graphs chosen to hit
    its decisions: degree ties, several components, isolated vertices, unsymmetric patterns, duplicate
    entries, diagonal entries, a star and a path;
  * against committed orders produced by it (tests/golden/rcm_orders.json, made by this file's
    `python tests/test_reorder.py --make-golden`);
  * "<path>__RCM" loads the same permuted entries as the reference's load_matrix; "__GP4" permutes nothing.
Host code only: no GPU needed.
"""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import spmv_cache_trace_b200 as sp
from oracle.oracle import Ref
from spmv_cache_trace_b200 import matrix_market

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "rcm_orders.json")
MTX = os.path.join(HERE, "golden", "poisson2D.mtx")


def graph(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "random_sym":  # symmetric pattern, many degree ties
        n, m = 60, 150
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
        i, j = np.concatenate([a, b, np.arange(n)]), np.concatenate([b, a, np.arange(n)])
    elif kind == "unsym_dups":  # directed edges, duplicates kept, isolated vertices
        n, m = 80, 120
        i, j = rng.integers(0, n // 2, m), rng.integers(0, n, m)
        i, j = np.concatenate([i, i[:20]]), np.concatenate([j, j[:20]])
    elif kind == "components":  # three separate grids + isolated nodes
        n = 3 * 16 + 5
        i, j = [], []
        for c in range(3):
            for r in range(4):
                for q in range(4):
                    v = c * 16 + r * 4 + q
                    for dr, dq in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                        if 0 <= r + dr < 4 and 0 <= q + dq < 4:
                            i.append(v); j.append(c * 16 + (r + dr) * 4 + q + dq)
        i, j = np.array(i), np.array(j)
        p = rng.permutation(n)
        i, j = p[i], p[j]
    elif kind == "star_path":
        n = 40
        i = np.concatenate([np.zeros(19, int), np.arange(1, 20), np.arange(20, 39), np.arange(21, 40)])
        j = np.concatenate([np.arange(1, 20), np.zeros(19, int), np.arange(21, 40), np.arange(20, 39)])
    else:
        raise ValueError(kind)
    k = rng.permutation(len(i))
    return n, (i[k] + 1).astype(np.int32), (j[k] + 1).astype(np.int32), rng.uniform(-1, 1, len(i))


KINDS = [(k, s) for k in ("random_sym", "unsym_dups", "components", "star_path") for s in (1, 2)]


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libspmvref.so not built")
@pytest.mark.parametrize("kind,seed", KINDS)
def test_rcm_order_matches_the_reference(kind, seed, capfd):
    n, i, j, a = graph(kind, seed)
    want = Ref().from_entries(n, n, i, j, a).order_rcm()
    capfd.readouterr()  # the reference prints progress to stdout
    mm = matrix_market.from_entries(n, n, i, j, a)
    got = matrix_market.find_new_order_RCM(mm)
    assert np.array_equal(got, want)
    assert sorted(got) == list(range(n))  # a permutation


def test_rcm_golden_orders():
    gold = json.load(open(GOLD))
    for g in gold["cases"]:
        n, i, j, a = graph(g["kind"], g["seed"])
        got = matrix_market.find_new_order_RCM(matrix_market.from_entries(n, n, i, j, a))
        assert got.tolist() == g["new_order"], (g["kind"], g["seed"])
    # the reference's golden fixture: RCM narrows the band of poisson2D
    mm = matrix_market.load_matrix(MTX)
    order = matrix_market.find_new_order_RCM(mm)
    assert order.tolist() == gold["poisson2D_new_order"]


def test_permute_and_path_suffixes():
    mm = matrix_market.load_matrix(MTX)
    i0, j0, a0 = mm.row_indices(), mm.column_indices(), mm.values_real()
    order = matrix_market.find_new_order_RCM(mm)
    pm = matrix_market.load_matrix(MTX + "__RCM")
    assert np.array_equal(pm.row_indices(), order[i0 - 1] + 1) and np.array_equal(pm.column_indices(), order[j0 - 1] + 1)
    assert np.array_equal(pm.values_real(), a0)
    mm.permute(order)
    assert np.array_equal(mm.row_indices(), pm.row_indices()) and np.array_equal(mm.column_indices(), pm.column_indices())
    bw0 = int(np.abs(i0 - j0).max())
    bw1 = int(np.abs(pm.row_indices() - pm.column_indices()).max())
    assert bw1 <= bw0
    # graph partitioning without METIS is the identity (matrix-market-reorder.cpp:172-180)
    gp = matrix_market.load_matrix(MTX + "__GP4")
    assert np.array_equal(gp.row_indices(), i0) and np.array_equal(gp.column_indices(), j0)
    assert np.array_equal(matrix_market.find_new_order_GP(gp, 4), np.arange(gp.rows))
    both = matrix_market.load_matrix(MTX + "__GP4__RCM")
    assert np.array_equal(both.row_indices(), pm.row_indices())
    if Ref.available():
        rm = Ref().load(MTX + "__RCM")
        ri, rj, ra = rm.entries()
        assert np.array_equal(ri, pm.row_indices()) and np.array_equal(rj, pm.column_indices()) and np.array_equal(ra, a0)
    # errors: rectangular matrices and wrong-size permutations
    rect = matrix_market.from_entries(3, 4, [1, 2], [1, 4], [1.0, 2.0])
    with pytest.raises(sp.matrix_error):
        matrix_market.find_new_order_RCM(rect)
    with pytest.raises(sp.matrix_error):
        mm.permute(np.arange(5))
    with pytest.raises(sp.matrix_error):
        mm.permute(np.full(mm.rows, mm.rows))  # out of range
    with pytest.raises(sp.matrix_error):
        matrix_market.load_matrix("/no/such/file.mtx__RCM")


def test_rcm_cost_is_linear_on_many_components():
    """The reference rescans every node per component (quadratic); 200 000 isolated vertices plus a chain
    must take well under a second here."""
    import time
    n = 200000
    i = np.arange(1, 1000, dtype=np.int32)
    mm = matrix_market.from_entries(n, n, np.concatenate([i, i + 1]), np.concatenate([i + 1, i]), np.ones(2 * len(i)))
    t0 = time.time()
    order = matrix_market.find_new_order_RCM(mm)
    assert time.time() - t0 < 5.0
    assert len(set(order.tolist())) == n


if __name__ == "__main__" and "--make-golden" in sys.argv:
    cases = []
    for kind, seed in KINDS:
        n, i, j, a = graph(kind, seed)
        cases.append({"kind": kind, "seed": seed, "new_order": Ref().from_entries(n, n, i, j, a).order_rcm().tolist()})
    doc = {"generator": "python tests/test_reorder.py --make-golden (find_new_order_RCM of the reference via oracle/_ref)",
           "cases": cases, "poisson2D_new_order": Ref().load(MTX).order_rcm().tolist()}
    json.dump(doc, open(GOLD, "w"))
    print("wrote", GOLD)
