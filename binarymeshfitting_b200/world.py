"""Host-side LOD leaf enumeration and multi-GPU partitioning -- the caller that feeds the hot path.

Mirrors WorldOctree::init / split_leaves / split_node / node_needs_split / create_chunk
(WorldOctree.cpp:117-127, 129-173, 175-210, 239-249, 263-269) and the per-chunk overlap rule of
ChunkGenerator::extract_chunk (ChunkGenerator.cpp:98).  Pure host logic in float32 (numpy scalars) so
leaf sets, positions, sizes, levels and Morton codes are identical to the reference's; pinned by
tests/test_world.py against the compiled reference and the committed golden leaf list.
"""
import numpy as np

from . import capi

F = np.float32

# Tables::MCDX/MCDY/MCDZ (Tables.hpp:7-9): octree children are enumerated in MC-corner order
MCDX = (0, 1, 1, 0, 0, 1, 1, 0)
MCDY = (0, 0, 0, 0, 1, 1, 1, 1)
MCDZ = (0, 0, 1, 1, 0, 0, 1, 1)


class WorldProperties:
    """WorldProperties() defaults (WorldOctree.cpp:20-33)."""

    def __init__(self, **kw):
        self.split_multiplier = 1.0
        self.size_modifier = 0.0
        self.max_level = 7
        self.min_level = 1
        self.process_iters = 0
        self.chunk_resolution = 32
        self.overlap = 0.035
        self.boundary_processing = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def node_needs_split(props, focus, pos, size, level):
    """WorldOctree::node_needs_split (WorldOctree.cpp:239-249); middle = pos + size*0.5 (WorldOctreeNode.cpp:26)."""
    if level >= props.max_level:
        return False
    if level < props.min_level:
        return True
    half = F(size) * F(0.5)
    mid = (F(pos[0]) + half, F(pos[1]) + half, F(pos[2]) + half)
    # glm::distance(middle, center) = length(center - middle), dot summed x,y,z left to right
    dx, dy, dz = F(focus[0]) - mid[0], F(focus[1]) - mid[1], F(focus[2]) - mid[2]
    d = np.sqrt(F(F(dx * dx) + F(dy * dy)) + F(dz * dz), dtype=F)
    rhs = F(F(F(size) * F(props.split_multiplier)) + F(props.size_modifier)) + half
    return bool(d < rhs)


def split_leaves(props, world_size=256, focus=(0.0, 0.0, 0.0)):
    """Static LOD build.  Returns (pos_size[n,4] f32, level[n] i32, morton[n] u64) in the reference's leaf order
    (stack DFS, children pushed 0..7 and popped LIFO -- not Morton order)."""
    size = F(world_size * 2)  # WorldOctree::init doubles the size (WorldOctree.cpp:119)
    p = F(1) * size * F(-0.5)
    stack = [((p, p, p), size, 0, 1)]
    out_ps, out_lv, out_mc = [], [], []
    while stack:
        pos, sz, level, code = stack.pop()
        if not node_needs_split(props, focus, pos, sz, level):
            out_ps.append((pos[0], pos[1], pos[2], sz))
            out_lv.append(level)
            out_mc.append(code)
            continue
        c_size = F(sz * F(0.5))
        for i in range(8):
            cpos = (F(pos[0] + F(MCDX[i]) * c_size), F(pos[1] + F(MCDY[i]) * c_size), F(pos[2] + F(MCDZ[i]) * c_size))
            ccode = (code << 3) | (MCDX[i] | (MCDY[i] << 1) | (MCDZ[i] << 2))
            stack.append((cpos, c_size, level + 1, ccode))
    return np.array(out_ps, np.float32).reshape(-1, 4), np.array(out_lv, np.int32), np.array(out_mc, np.uint64)


def chunk_overlap(props, level):
    """ChunkGenerator.cpp:98."""
    iters = props.process_iters
    if level == props.max_level and (not props.boundary_processing or iters == 0):
        return F(0.0)
    return F(F(props.overlap) + F(F(0.005) * F(iters)))


def make_descs(props, pos_size, levels, mortons=None):
    d = capi.make_chunk_descs(pos_size)
    d["level"] = levels
    d["overlap"] = [chunk_overlap(props, int(l)) for l in levels]
    if mortons is not None:
        d["morton"] = mortons
    return d


def grid_chunks(n_per_axis=16, chunk_size=16.0, origin=(-128.0, -128.0, -128.0)):
    """Config 3: n^3 grid of equal chunks (x-major like the reference's own loops)."""
    i = np.arange(n_per_axis, dtype=np.float32)
    X, Y, Z = np.meshgrid(i, i, i, indexing="ij")
    ps = np.stack([F(origin[0]) + X.ravel() * F(chunk_size), F(origin[1]) + Y.ravel() * F(chunk_size),
                   F(origin[2]) + Z.ravel() * F(chunk_size), np.full(X.size, chunk_size, np.float32)], axis=1)
    return np.ascontiguousarray(ps, np.float32)


MORTON_MAX_LEVEL = 20  # 3 bits per level + the sentinel fit 64 bits


def morton_key(code):
    """Depth-normalised sort key of a WorldOctreeNode::morton_code (sentinel 1, then 3 bits per level: x | y<<1 | z<<2,
    WorldOctree.cpp:196-200).  The raw code orders level-major (deeper nodes are numerically larger); stripping the sentinel
    and left-aligning the digits orders leaves of MIXED levels along one Z-curve, so contiguous ranges are spatially compact.
    Ties (a node and its first descendants) break by level."""
    code = int(code)
    level = (code.bit_length() - 1) // 3
    return ((code ^ (1 << (3 * level))) << (3 * (MORTON_MAX_LEVEL - level)), level)


def column_key(code):
    """Sort key for heightfield worlds: the same digits, regrouped -- the (z, x) bits of all levels first (a 2-D Z-curve over the chunk
    COLUMNS), the y bits after them.  A contiguous range of this order is a bundle of whole columns (plus at most one cut column at
    each end), so the ranks of a 2-D noise terrain do not sample each other's noise sheets: with the octree's own order (z, y, x) the
    second cut is along y and two ranks compute every sheet twice."""
    code = int(code)
    level = (code.bit_length() - 1) // 3
    zx = yy = 0
    for l in range(level - 1, -1, -1):
        d = (code >> (3 * l)) & 7
        zx = (zx << 2) | (((d >> 2) & 1) << 1) | (d & 1)
        yy = (yy << 1) | ((d >> 1) & 1)
    zx <<= 2 * (MORTON_MAX_LEVEL - level)
    yy <<= MORTON_MAX_LEVEL - level
    return ((zx << MORTON_MAX_LEVEL) | yy, level)


def morton_order(mortons, columns=False):
    kf = column_key if columns else morton_key
    keys = [kf(c) for c in np.asarray(mortons, np.uint64).tolist()]
    return np.array(sorted(range(len(keys)), key=keys.__getitem__), np.int64)


def grid_mortons(n_per_axis):
    """Morton codes (reference convention: sentinel, digit = x | y<<1 | z<<2 per level) of grid_chunks' n^3 chunks, n a power of two."""
    levels = int(n_per_axis).bit_length() - 1
    assert 1 << levels == n_per_axis
    i = np.arange(n_per_axis, dtype=np.uint64)
    X, Y, Z = (a.ravel() for a in np.meshgrid(i, i, i, indexing="ij"))
    code = np.full(X.size, 1, np.uint64)
    for l in range(levels - 1, -1, -1):
        digit = ((X >> np.uint64(l)) & np.uint64(1)) | (((Y >> np.uint64(l)) & np.uint64(1)) << np.uint64(1)) | (((Z >> np.uint64(l)) & np.uint64(1)) << np.uint64(2))
        code = (code << np.uint64(3)) | digit
    return code


def partition(mortons, costs, n_parts, columns=False):
    """Multi-GPU sharding by octree node: sort leaves along the Z-curve (depth-normalised Morton key) and deal
    contiguous, cost-balanced ranges.  Returns a list of index arrays (into the original order), one per part.
    columns=True: the column-major variant of the curve (column_key) for heightfield samplers."""
    order = morton_order(mortons, columns)
    c = np.asarray(costs, np.float64)[order]
    cum = np.cumsum(c)
    total = cum[-1] if len(cum) else 0.0
    bounds = [int(np.searchsorted(cum, total * k / n_parts, side="right")) for k in range(1, n_parts)]
    parts, lo = [], 0
    for b in bounds + [len(order)]:
        b = max(b, lo)
        parts.append(order[lo:b])
        lo = b
    return parts


def border_chunks(pos_size, group):
    """Multi-GPU seam scheme ("a final host gather ... and a WorldStitcher seam pass"): every rank stitches
    the dual cells that lie entirely inside its own chunks; the cells that span ranks only involve chunks that touch
    (face, edge or corner) a chunk of another group.  Returns the indices of those border chunks, in batch order --
    the batch of the final bmf_batch_stitch(group, cross_group_only=1) pass on the gathering rank."""
    ps = np.asarray(pos_size, np.float64).reshape(-1, 4)
    group = np.asarray(group)
    smin = ps[:, 3].min()
    # lattice origin anchored to the octree like bmf_batch_stitch: step down from the largest chunk in multiples of its size
    big = int(np.argmax(ps[:, 3]))
    org = ps[big, :3] - np.ceil((ps[big, :3] - ps[:, :3].min(axis=0)) / ps[big, 3] - 1e-6) * ps[big, 3]
    lo = np.rint((ps[:, :3] - org) / smin).astype(np.int64)
    ext = np.rint(ps[:, 3] / smin).astype(np.int64)
    G = (lo + ext[:, None]).max(axis=0)
    grid = np.full(tuple(G), -1, np.int64)
    for i in range(len(ps)):
        grid[lo[i, 0]:lo[i, 0] + ext[i], lo[i, 1]:lo[i, 1] + ext[i], lo[i, 2]:lo[i, 2] + ext[i]] = group[i]
    out = []
    for i in range(len(ps)):
        a = np.maximum(lo[i] - 1, 0)
        b = np.minimum(lo[i] + ext[i] + 1, G)
        nb = grid[a[0]:b[0], a[1]:b[1], a[2]:b[2]]
        if np.any((nb != group[i]) & (nb >= 0)):
            out.append(i)
    return np.array(out, np.int64)


# ---- incremental LOD updates: the WorldWatcher policy (WorldWatcher.cpp:34-134 update, :155-212 check_leaves /
# handle_split_check / handle_group_check, :214-238 process_batch, :240-352 post_process_batch, :354-392 split_node /
# group_node_1) as one synchronous tick -- without the watcher thread, the render-thread handshake and the draw flags.
def node_needs_group(props, focus, pos, size, level, group_multiplier=2.0):
    """WorldOctree::node_needs_group (WorldOctree.cpp:251-261)"""
    if level < props.min_level:
        return False
    if level > props.max_level:
        return True
    half = F(size) * F(0.5)
    mid = (F(pos[0]) + half, F(pos[1]) + half, F(pos[2]) + half)
    dx, dy, dz = F(focus[0]) - mid[0], F(focus[1]) - mid[1], F(focus[2]) - mid[2]
    d = np.sqrt(F(F(dx * dx) + F(dy * dy)) + F(dz * dz), dtype=F)
    rhs = F(F(F(size) * F(group_multiplier)) + F(props.size_modifier)) + half
    return bool(d > rhs)


class _Node:
    __slots__ = ("pos", "size", "level", "code", "parent", "children", "leaf", "mark")

    def __init__(self, pos, size, level, code, parent):
        self.pos, self.size, self.level, self.code, self.parent = pos, size, level, code, parent
        self.children, self.leaf, self.mark = None, True, 0


class LodWatcher:
    """Leaf set of the world that follows a moving focus point.  `tick(focus)` = check_leaves + process_batch +
    post_process_batch: walks the renderables (leaves, in list order), marks at most max_gen/8 splits and any number of
    groups, applies them, and returns the nodes that have to be generated (8 children per split, the parent per group) --
    the batch WorldWatcher::update hands to ChunkGenerator::process_queue."""

    def __init__(self, props, world_size=256, focus=(0.0, 0.0, 0.0), group_multiplier=2.0, start="static"):
        """start = "static": the renderables are the leaves of WorldOctree::split_leaves for `focus`; start = "root": only the
        root is drawable, like right after WorldWatcher::init (WorldWatcher.cpp:21-31) -- the ticks then build the world."""
        self.props, self.group_multiplier = props, group_multiplier
        size = F(world_size * 2)
        p = F(1) * size * F(-0.5)
        self.root = _Node((p, p, p), size, 0, 1, None)
        self.renderables = []
        if start == "root":
            self.renderables.append(self.root)
            return
        stack = [self.root]  # WorldOctree::split_leaves order (LIFO)
        while stack:
            n = stack.pop()
            if not node_needs_split(props, focus, n.pos, n.size, n.level):
                self.renderables.append(n)
                continue
            stack.extend(self._split(n))

    def _split(self, n):
        c_size = F(n.size * F(0.5))
        n.children = []
        for i in range(8):
            cpos = (F(n.pos[0] + F(MCDX[i]) * c_size), F(n.pos[1] + F(MCDY[i]) * c_size), F(n.pos[2] + F(MCDZ[i]) * c_size))
            n.children.append(_Node(cpos, c_size, n.level + 1, (n.code << 3) | (MCDX[i] | (MCDY[i] << 1) | (MCDZ[i] << 2)), n))
        n.leaf = False
        return n.children

    def tick(self, focus, max_gen=400):
        batch, counter = [], 0
        for n in self.renderables:  # check_leaves
            if counter >= max_gen:
                break
            if node_needs_split(self.props, focus, n.pos, n.size, n.level):
                if n.leaf and not n.mark:
                    n.mark = 1
                    batch.append(n)
                counter += 8
            elif n.leaf and n.parent is not None and node_needs_group(self.props, focus, n.parent.pos, n.parent.size, n.parent.level, self.group_multiplier):
                par = n.parent
                if not par.mark and all(c.leaf and not c.mark for c in par.children):
                    par.mark = 2
                    batch.append(par)
        generate, gone = [], set()
        for n in batch:  # process_batch + post_process_batch
            if n.mark == 1:
                kids = self._split(n)
                generate.extend(kids)
                self.renderables.extend(kids)
                gone.add(id(n))
            else:
                for c in n.children:
                    gone.add(id(c))
                n.children, n.leaf = None, True
                generate.append(n)
                self.renderables.append(n)
            n.mark = 0
        if gone:
            self.renderables = [r for r in self.renderables if id(r) not in gone]
        return generate

    @staticmethod
    def arrays(nodes):
        ps = np.array([[n.pos[0], n.pos[1], n.pos[2], n.size] for n in nodes], np.float32).reshape(-1, 4)
        return ps, np.array([n.level for n in nodes], np.int32), np.array([n.code for n in nodes], np.uint64)

    def leaves(self):
        return self.arrays(self.renderables)
