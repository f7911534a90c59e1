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


# ------------------------------------------------------------------ config 4: subtree sweep

class SubtreeSweep:
    """BASELINE config 4: every descendant, down to `depth` levels, of ONE tile (root_level, tx, ty)
    -- 4**depth leaf tiles plus a third as many ancestors -- partitioned by subtree.

    Slots: [0, root_level] the ancestor chain of the sweep root (levels 0 .. root_level, replicated on
    every rank), then the top tree (depths 1 .. k below the root, replicated: (4**(k+1) - 4) / 3 tiles),
    then one recycled region per depth k+1 .. depth for the unit being produced.  The 4**k units (the
    root's descendants at depth k) are dealt round-robin to the ranks; a unit is a contiguous Morton
    range at every level, so each of its levels is ONE pl_produce_range call.
    """

    def __init__(self, root_level, tx, ty, depth, unit_depth=7):
        assert 0 <= tx < (1 << root_level) and 0 <= ty < (1 << root_level)
        self.root_level, self.tx, self.ty, self.depth = root_level, tx, ty, depth
        self.k = max(depth - unit_depth, 0)           # units are the root's descendants at depth k
        self.root_morton = morton_encode(tx, ty)
        self.chain_slots = root_level + 1
        self.top_off = [self.chain_slots]              # slot offset of top-tree depth j = 1 .. k
        for j in range(1, self.k + 1):
            self.top_off.append(self.top_off[-1] + 4 ** j)
        self.unit_off = [self.top_off[-1]]             # slot offset of unit depth e = 1 .. depth - k
        for e in range(1, depth - self.k + 1):
            self.unit_off.append(self.unit_off[-1] + 4 ** e)
        self.capacity = self.unit_off[-1]
        self.top_off, self.unit_off = self.top_off[:-1], self.unit_off[:-1]

    def units(self):
        return list(range(4 ** self.k))

    def units_of_rank(self, rank, world):
        return self.units()[rank::world]

    def tiles_in_unit(self):
        return sum(4 ** e for e in range(1, self.depth - self.k + 1))

    def replicated_tiles(self):
        return self.chain_slots + sum(4 ** j for j in range(1, self.k + 1))

    def total_tiles(self):
        """every tile of the sweep once: chain + subtree"""
        return self.root_level + sum(4 ** j for j in range(self.depth + 1))

    def leaf_region(self):
        """(slot0, n) of the unit's deepest level"""
        return self.unit_off[-1] if self.unit_off else self.chain_slots - 1, 4 ** (self.depth - self.k)

    def prologue(self):
        """batches every rank produces first: the chain and the top tree.
        yields (level, morton0, n, out_slot0, parent_slot0, parent_morton0)"""
        for l in range(self.root_level + 1):
            m = self.root_morton >> (2 * (self.root_level - l))
            yield l, m, 1, l, max(l - 1, 0), m >> 2
        for j in range(1, self.k + 1):
            m0 = self.root_morton << (2 * j)
            parent_slot0 = self.root_level if j == 1 else self.top_off[j - 2]
            yield self.root_level + j, m0, 4 ** j, self.top_off[j - 1], parent_slot0, m0 >> 2

    def unit_batches(self, unit):
        """the levels of one unit, breadth first"""
        um = (self.root_morton << (2 * self.k)) | unit           # the unit root's Morton index at depth k
        unit_root_slot = self.root_level if self.k == 0 else self.top_off[self.k - 1] + unit
        for e in range(1, self.depth - self.k + 1):
            m0 = um << (2 * e)
            parent_slot0 = unit_root_slot if e == 1 else self.unit_off[e - 2]
            yield self.root_level + self.k + e, m0, 4 ** e, self.unit_off[e - 1], parent_slot0, m0 >> 2
