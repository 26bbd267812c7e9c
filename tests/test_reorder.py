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


# ---- graph-partitioning order (find_new_order_GP, matrix-market-reorder.cpp:183-278) ------------------------------

def _edgecut(i, j, part):
    """undirected simple edges whose ends lie in different parts"""
    a, b = np.minimum(i, j) - 1, np.maximum(i, j) - 1
    keep = a != b
    e = np.unique(np.stack([a[keep], b[keep]], 1), axis=0)
    return int((part[e[:, 0]] != part[e[:, 1]]).sum())


def _grid7(n, shuffle_seed=None):
    """3-D 7-point pattern on an n^3 grid (1-based entries), optionally with the vertices renumbered at random"""
    from oracle.generators_ref import stencil_entries
    i, j, a = stencil_entries(1, n, n, n)
    i, j = i.astype(np.int64), j.astype(np.int64)
    if shuffle_seed is not None:
        p = np.random.default_rng(shuffle_seed).permutation(n ** 3)
        i, j = p[i - 1] + 1, p[j - 1] + 1
    return n ** 3, i.astype(np.int32), j.astype(np.int32), a


def test_order_from_parts_matches_the_reference_loops(oracle):
    """The grouping step after the METIS call (:246-266), restated in oracle/spmv_oracle.c loop for loop."""
    rng = np.random.default_rng(3)
    for n, k in ((1, 1), (17, 4), (1000, 16), (4096, 7), (50, 64)):
        part = rng.integers(0, k, n).astype(np.int32)
        if n > 20:
            part[rng.integers(0, n, n // 3)] = k - 1  # uneven parts, possibly empty ones
        want = oracle.order_from_parts(part, k)
        got = matrix_market.order_from_parts(part, k)
        assert np.array_equal(got, want)
        assert sorted(got.tolist()) == list(range(n))
        inv = np.argsort(got)  # inv[new] = old: parts ascend, old indices ascend inside a part
        assert np.all(np.diff(part[inv]) >= 0)
        for q in range(k):
            assert np.all(np.diff(inv[part[inv] == q]) > 0)
    with pytest.raises(sp.matrix_error):
        matrix_market.order_from_parts(np.array([0, 5], np.int32), 4)


@pytest.mark.parametrize("n,k,shuffled", [(12, 8, False), (12, 8, True), (10, 3, True), (16, 16, True)])
def test_kway_partition_properties(n, k, shuffled):
    """The METIS stand-in: a valid partition (every part number in range, none empty), inside the reference's balance
    bound (ubvec = 1.05, :200), deterministic, independent of the vertex numbering up to quality, and with a cut far
    below what the numbering it was given offers."""
    N, i, j, a = _grid7(n, 11 if shuffled else None)
    mm = matrix_market.from_entries(N, N, i, j, a)
    part, cut = matrix_market.partition_kway(mm, k)
    assert part.min() == 0 and part.max() == k - 1 and len(np.unique(part)) == k
    sizes = np.bincount(part, minlength=k)
    assert sizes.max() <= max(-(-N // k), int(1.05 * N / k))
    assert cut == _edgecut(i, j, part)
    part2, cut2 = matrix_market.partition_kway(mm, k)
    assert np.array_equal(part, part2) and cut == cut2
    contiguous = (np.arange(N, dtype=np.int64) * k // N).astype(np.int32)  # the reference row partition of this numbering
    cut_contig = _edgecut(i, j, contiguous)
    edges = _edgecut(i, j, np.arange(N, dtype=np.int32))  # every edge
    # recursive level-structure bisection of a grid: no worse than 1.5 x what k axis-aligned slabs cut ((k - 1) n^2 edges)
    assert cut <= 1.5 * (k - 1) * n * n < edges, (cut, edges)
    if shuffled:
        assert cut < cut_contig / 3, (cut, cut_contig)  # a random numbering cuts (k - 1)/k of all edges
    # errors
    with pytest.raises(sp.matrix_error):
        matrix_market.partition_kway(mm, 0)
    with pytest.raises(sp.matrix_error):
        matrix_market.partition_kway(matrix_market.from_entries(3, 4, [1], [4], [1.0]), 2)


def test_kway_partition_components_isolated_vertices_and_directed_entries():
    # two chains, a star, isolated vertices; entries given in ONE direction only (the graph is made undirected)
    i = np.concatenate([np.arange(1, 30), np.arange(41, 70), np.full(15, 80)]).astype(np.int32)
    j = np.concatenate([np.arange(2, 31), np.arange(42, 71), np.arange(81, 96)]).astype(np.int32)
    n = 120
    mm = matrix_market.from_entries(n, n, np.concatenate([i, [5, 5]]), np.concatenate([j, [5, 6]]), np.ones(len(i) + 2))  # + diagonal, + duplicate
    for k in (1, 2, 5, 120, 200):
        part, cut = matrix_market.partition_kway(mm, k)
        assert part.min() >= 0 and part.max() < k
        assert np.bincount(part, minlength=k).max() <= max(-(-n // k), int(1.05 * n / k))
        assert cut == _edgecut(np.concatenate([i, [5]]), np.concatenate([j, [6]]), part)
        if k == 1:
            assert cut == 0
    part, cut = matrix_market.partition_kway(mm, 2)
    assert cut <= 2  # two parts of 60: at most the chains are cut once each
    empty = matrix_market.from_entries(0, 0, [], [], [])
    part, cut = matrix_market.partition_kway(empty, 4)
    assert part.size == 0 and cut == 0


def test_gp_order_as_a_layout_lever_for_the_row_partition(tmp_path):
    """What the reordering is FOR on this path: a shuffled grid makes every rank of the row-partitioned mode reference all
    of x; after "__RCM" + the graph-partitioning order (the partitioner standing in for METIS) the ranks' rows reference
    their own slice and a thin halo again -- seen in the exchange plan of the same 8-way row partition."""
    n, P = 14, 8
    N, i, j, a = _grid7(n, 5)
    path = str(tmp_path / "shuffled.mtx")
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (N, N, len(i)))
        for r, c, v in zip(i, j, a):
            f.write("%d %d %.17g\n" % (r, c, v))
    plain = matrix_market.load_matrix(path + "__GP%d" % P)  # default: identity, like the reference without METIS
    assert np.array_equal(plain.row_indices(), i)
    sp.set_global_option("mm.gp_partitioner", 1)
    try:
        gp = matrix_market.load_matrix(path + "__GP%d" % P)
        both = matrix_market.load_matrix(path + "__GP%d__RCM" % P)  # RCM first, then GP (matrix-market.cpp:817-825)
        assert np.array_equal(matrix_market.find_new_order_GP(matrix_market.load_matrix(path), P),
                              matrix_market.find_new_order_GP(matrix_market.load_matrix(path), P, partitioner=True))
    finally:
        sp.set_global_option("mm.gp_partitioner", 0)
    order = matrix_market.find_new_order_GP(matrix_market.load_matrix(path), P, partitioner=True)
    assert np.array_equal(gp.row_indices(), order[i - 1] + 1) and np.array_equal(gp.column_indices(), order[j - 1] + 1)
    assert sorted(zip(both.row_indices().tolist(), both.column_indices().tolist())) != sorted(zip(i.tolist(), j.tolist()))

    def remote_columns(mm):
        """columns outside the rank's own slice that its rows reference, summed over the ranks of the reference partition"""
        starts = sp.partition.rows_ref(mm.rows, P)
        r, c = mm.row_indices().astype(np.int64) - 1, mm.column_indices().astype(np.int64) - 1
        owner_r = np.searchsorted(starts, r, side="right") - 1
        owner_c = np.searchsorted(starts, c, side="right") - 1
        far = owner_r != owner_c
        return len(np.unique(owner_r[far] * mm.rows + c[far]))

    shuffled, grouped, banded = remote_columns(plain), remote_columns(gp), remote_columns(both)
    assert grouped < shuffled / 3 and banded < shuffled / 3, (shuffled, grouped, banded)


GP_GOLD = os.path.join(HERE, "golden", "gp_partitions.json")
GP_CASES = [("grid", 9, 4, None), ("grid", 9, 4, 3), ("grid", 12, 8, 11), ("random_sym", 1, 5, None), ("components", 2, 3, None)]


def _gp_case(kind, a, k, shuffle):
    if kind == "grid":
        N, i, j, v = _grid7(a, shuffle)
    else:
        N, i, j, v = graph(kind, a)
    return matrix_market.from_entries(N, N, i, j, v), k


def test_kway_partition_golden():
    """The partitioner is deterministic: committed partitions (made by `python tests/test_reorder.py --make-golden`)
    pin it against accidental changes; the orders derived from them go through the exact grouping step."""
    gold = json.load(open(GP_GOLD))
    assert len(gold["cases"]) == len(GP_CASES)
    for g, case in zip(gold["cases"], GP_CASES):
        mm, k = _gp_case(*case)
        part, cut = matrix_market.partition_kway(mm, k)
        assert part.tolist() == g["part"] and cut == g["edgecut"], case
        assert matrix_market.find_new_order_GP(mm, k, partitioner=True).tolist() == g["new_order"], case


if __name__ == "__main__" and "--make-golden" in sys.argv:
    gp = []
    for case in GP_CASES:
        mm, k = _gp_case(*case)
        part, cut = matrix_market.partition_kway(mm, k)
        gp.append({"case": list(case), "part": part.tolist(), "edgecut": cut,
                   "new_order": matrix_market.find_new_order_GP(mm, k, partitioner=True).tolist()})
    json.dump({"generator": "python tests/test_reorder.py --make-golden (spmvb200_mm_partition_kway, this repository's partitioner)",
               "cases": gp}, open(GP_GOLD, "w"))
    print("wrote", GP_GOLD)

    cases = []
    for kind, seed in KINDS:
        n, i, j, a = graph(kind, seed)
        cases.append({"kind": kind, "seed": seed, "new_order": Ref().from_entries(n, n, i, j, a).order_rcm().tolist()})
    doc = {"generator": "python tests/test_reorder.py --make-golden (find_new_order_RCM of the reference via oracle/_ref)",
           "cases": cases, "poisson2D_new_order": Ref().load(MTX).order_rcm().tolist()}
    json.dump(doc, open(GOLD, "w"))
    print("wrote", GOLD)
