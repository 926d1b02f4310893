"""Executable argument for `stencil_half` (hsmc_b200/csrc/geom.cuh): walking, from every owned cell, the own
cell plus the 13 neighbour cells lexicographically after it visits every unordered pair of adjacent cells
exactly once -- on a periodic grid with at least four cells per axis, and on x-slabs where only owned cells
walk and ghost layers are only ever looked at.  The GPU tests check the kernels' counts against the
reference; this pins the index arithmetic, including the wrap of each axis and the slab faces."""
import itertools

import pytest

FORWARD = [(0, 0, 1)] + [(0, 1, dz) for dz in (-1, 0, 1)] + [(1, dy, dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]


def _pairs_single_gpu(nx, ny, nz):
    seen = {}
    for c in itertools.product(range(nx), range(ny), range(nz)):
        for d in FORWARD:
            n = ((c[0] + d[0]) % nx, (c[1] + d[1]) % ny, (c[2] + d[2]) % nz)
            key = frozenset((c, n))
            seen[key] = seen.get(key, 0) + 1
    return seen


def _adjacent_pairs(nx, ny, nz):
    out = set()
    for c in itertools.product(range(nx), range(ny), range(nz)):
        for d in itertools.product((-1, 0, 1), repeat=3):
            if d != (0, 0, 0):
                out.add(frozenset((c, ((c[0] + d[0]) % nx, (c[1] + d[1]) % ny, (c[2] + d[2]) % nz))))
    return out


@pytest.mark.parametrize("shape", [(4, 4, 4), (4, 6, 8), (6, 4, 10), (8, 8, 4)])
def test_forward_half_visits_every_adjacent_cell_pair_once(shape):
    assert len(FORWARD) == 13 and len(set(FORWARD)) == 13
    assert all((-d[0], -d[1], -d[2]) not in FORWARD for d in FORWARD)
    seen = _pairs_single_gpu(*shape)
    assert set(seen) == _adjacent_pairs(*shape)
    assert set(seen.values()) == {1}


@pytest.mark.parametrize("shape,world", [((8, 4, 4), 2), ((12, 4, 6), 3), ((16, 6, 4), 4)])
def test_slabs_see_every_pair_from_exactly_one_rank(shape, world):
    """rank r owns x-layers [lo, hi); its threads walk owned cells only, reading the ghost layer at hi (and
    never needing the one at lo - 1): a pair across a slab face is counted by the lower-x rank alone."""
    nx, ny, nz = shape
    per = nx // world
    seen = {}
    for r in range(world):
        lo, hi = r * per, (r + 1) * per
        for c in itertools.product(range(lo, hi), range(ny), range(nz)):
            for d in FORWARD:
                x = c[0] + d[0]
                assert lo <= x <= hi                       # owned layer or the right ghost layer, never further
                n = (x % nx, (c[1] + d[1]) % ny, (c[2] + d[2]) % nz)
                key = frozenset((c, n))
                seen[key] = seen.get(key, 0) + 1
    assert set(seen) == _adjacent_pairs(*shape) and set(seen.values()) == {1}
