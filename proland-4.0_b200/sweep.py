"""Subtree partitioning and breadth-first batch plan of a full-quadtree sweep.

Pure planning (no GPU, no torch): which (face, level-2 quad) subtrees a rank owns and which
Morton ranges it produces into which pool slots.  Tile order inside a level is Morton order
with x in the even bits -- the order TileSampler::getTiles visits children
(core/sources/proland/terrain/TileSampler.cpp:416-461), so the four children of slot s of one
level are four consecutive slots of the next.

Partition (north star / SURVEY 8e): subtrees are independent given their root's parent chain;
the planet is cut into 6 faces x 16 level-2 quads = 96 units, dealt round-robin to the ranks;
each rank recomputes the 21 tiles of levels 0..2 of every face it touches (replicated
ancestors, 2.5e-6 of the work), no halo, no data-path collective.
"""

ROOT_SLOTS = 21          # levels 0..2 of the current face: slots 0, 1..4, 5..20
UNIT_LEVEL = 2


def planet_units(faces=(1, 2, 3, 4, 5, 6)):
    return [(f, m2) for f in faces for m2 in range(4 ** UNIT_LEVEL)]


def units_of_rank(units, rank, world):
    return units[rank::world]


def region_offsets(max_level):
    """slot offset of the region holding depth d (level UNIT_LEVEL + d) of the current unit,
    d = 1..max_level-UNIT_LEVEL, and the pool capacity."""
    depths = max_level - UNIT_LEVEL
    off = [ROOT_SLOTS]
    for d in range(1, depths + 1):
        off.append(off[-1] + 4 ** d)
    return off[:-1], off[-1]


def pairs_in_units(units, max_level, count_roots=True):
    per_unit = sum(4 ** d for d in range(1, max_level - UNIT_LEVEL + 1))
    faces = len({f for f, _ in units})
    return len(units) * per_unit + (ROOT_SLOTS * faces if count_roots else 0)


def batches(units, max_level):
    """yields (face, level, morton0, n, out_slot0, parent_slot0, parent_morton0)."""
    off, _ = region_offsets(max_level)
    face_done = None
    for f, m2 in units:
        if f != face_done:
            yield f, 0, 0, 1, 0, 0, 0
            if max_level >= 1:
                yield f, 1, 0, 4, 1, 0, 0
            if max_level >= 2:
                yield f, 2, 0, 16, 5, 1, 0
            face_done = f
        for d in range(1, max_level - UNIT_LEVEL + 1):
            n = 4 ** d
            m0 = m2 << (2 * d)
            if d == 1:
                yield f, UNIT_LEVEL + d, m0, n, off[0], 5 + m2, m2
            else:
                yield f, UNIT_LEVEL + d, m0, n, off[d - 1], off[d - 2], m0 >> 2


def morton_encode(tx, ty):
    m = 0
    for b in range(24):
        m |= ((tx >> b) & 1) << (2 * b) | ((ty >> b) & 1) << (2 * b + 1)
    return m


def morton_decode(m):
    tx = ty = 0
    for b in range(24):
        tx |= ((m >> (2 * b)) & 1) << b
        ty |= ((m >> (2 * b + 1)) & 1) << b
    return tx, ty
