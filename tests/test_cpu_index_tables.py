"""Host-side checks of the packed index tables the kernels use (constants are read from the CUDA sources, so the tests follow
the code): the row pattern of Combinations::permutationsNoReplacement(n,3) (combinations.cpp:127-244), the partner row that
P3P's exchange of points 1 and 2 selects (p3p.cpp:101-121), the m-th index outside a sorted triple, and the 8-neighbourhood
direction tables of the border follower."""
import itertools
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K2 = open(os.path.join(ROOT, "rpg_monocular_pose_estimator_b200", "csrc", "k2_p3p_sweep.cu")).read()
K1 = open(os.path.join(ROOT, "rpg_monocular_pose_estimator_b200", "csrc", "k1_find_leds.cu")).read()


def _const(src, pattern):
    m = re.search(pattern, src)
    assert m, pattern
    return int(m.group(1), 16)


def reference_perm_rows(n):
    """Rows of permutationsNoReplacement(n,3), 0-based: for every lexicographic combination a<b<c the six rows
    [c b a],[c a b],[b c a],[b a c],[a b c],[a c b] (verified there by emulating combinations.cpp)."""
    rows = []
    for a, b, c in itertools.combinations(range(n), 3):
        rows += [(c, b, a), (c, a, b), (b, c, a), (b, a, c), (a, b, c), (a, c, b)]
    return rows


def test_packed_permutation_pattern_matches_reference_row_order():
    pat = _const(K2, r"\(uint32_t\)\((0x[0-9a-fA-F]+)ull >> \(6 \* r6\)\)")
    for n in range(3, 9):
        combos = list(itertools.combinations(range(n), 3))
        rows = reference_perm_rows(n)
        assert len(set(rows)) == n * (n - 1) * (n - 2)
        for j, row in enumerate(rows):
            abc = combos[j // 6]
            e = (pat >> (6 * (j % 6))) & 63
            got = (abc[e & 3], abc[(e >> 2) & 3], abc[(e >> 4) & 3])
            assert got == row, (n, j, got, row)


def test_partner_row_is_the_row_with_points_1_and_2_exchanged():
    pmap = _const(K2, r"\(\((0x[0-9a-fA-F]+) >> \(4 \* r6\)\) & 7\)")
    rows = reference_perm_rows(6)
    index = {r: i for i, r in enumerate(rows)}
    for j, (p0, p1, p2) in enumerate(rows):
        r6 = j % 6
        tj = j - r6 + ((pmap >> (4 * r6)) & 7)
        assert rows[tj] == (p1, p0, p2) and index[(p1, p0, p2)] == tj


def nth_unused(m, a, b, c):          # k2_p3p_sweep.cu: nth_unused
    m += m >= a
    m += m >= b
    m += m >= c
    return m


def test_nth_unused_enumerates_the_complement_in_order():
    for n in range(4, 17):
        for a, b, c in itertools.combinations(range(n), 3):
            rest = [i for i in range(n) if i not in (a, b, c)]
            assert [nth_unused(m, a, b, c) for m in range(n - 3)] == rest


def test_direction_tables():
    dxp = _const(K1, r"dir_dx\(int s\) \{ return \(int\)\(\((0x[0-9a-fA-F]+)u >>")
    dyp = _const(K1, r"dir_dy\(int s\) \{ return \(int\)\(\((0x[0-9a-fA-F]+)u >>")
    # OpenCV direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (image y grows downwards)
    dx = [1, 1, 0, -1, -1, -1, 0, 1]
    dy = [0, -1, -1, -1, 0, 1, 1, 1]
    for s in range(8):
        assert ((dxp >> (2 * s)) & 3) - 1 == dx[s]
        assert ((dyp >> (2 * s)) & 3) - 1 == dy[s]
