#!/usr/bin/env python
"""Generate the packed Marching-Cubes triangle table used by the CUDA kernels and by the oracle.

The triangulation table is DATA that defines the reference's output order (which edge vertex each
emitted index names, DMCChunk.cpp:567-575), so it cannot be re-derived: it is read from
MarchingCubes::tri_table (MCTable.h:25-283) and re-encoded in this project's own packed form:

    tri_pack[mask8] : uint64, nibble i (i < 15) = edge id of the i-th emitted index,
                      nibble 15 = number of emitted indices (0, 3, ..., 15)

Edge numbering (the reference's): e0-3 X-edges of cells (x, y+{0,0,1,1}, z+{0,1,0,1}),
e4-7 Y-edges of cells (x+{0,0,1,1}, y, z+{0,1,0,1}), e8-11 Z-edges of cells (x+{0,0,1,1}, y+{0,1,0,1}, z).
The script also verifies the table against edge_map (MCTable.h:5-23) and against the corner rule
"edge e is crossed iff its two corner bits differ" with corner i at offset (i>>2, (i>>1)&1, i&1).

Run in the authoring container (needs /root/reference); outputs are committed.
"""
import re
import sys

REF = "/root/reference/BinaryMeshFitting/MCTable.h"
TABLES = "/root/reference/BinaryMeshFitting/Tables.hpp"


def parse():
    src = open(REF).read()
    em = re.search(r"edge_map\[\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
    edge_map = [int(t, 16) for t in re.findall(r"0x[0-9A-Fa-f]+", em)]
    tt = re.search(r"tri_table\[\]\[16\]\s*=\s*\{(.*)\};", src, re.S).group(1)
    rows = re.findall(r"\{([^{}]*)\}", tt)
    tri = [[int(t) for t in r.replace(" ", "").split(",") if t] for r in rows]
    assert len(edge_map) == 256 and len(tri) == 256 and all(len(r) == 16 for r in tri)
    return edge_map, tri


def ref_patches():
    """Tables::EdgeTable (Tables.hpp:96-354): patches separated by -1, row terminated by -2"""
    src = open(TABLES).read()
    et = re.search(r"EdgeTable\[256\]\[16\]\s*=\s*\{(.*?)\n\t\};", src, re.S).group(1)
    rows = re.findall(r"\{([^{}]*)\}", et)
    assert len(rows) == 256
    out = []
    for r in rows:
        vals = [int(t) for t in r.replace(" ", "").split(",") if t]
        masks, cur = [], 0
        for v in vals:
            if v >= 0:
                cur |= 1 << v
            else:
                if cur:
                    masks.append(cur)
                cur = 0
                if v == -2:
                    break
        out.append(sorted(masks))
    return out


def tri_patches(tri):
    """surface patches of every cell configuration = connected components of its triangles (sharing an edge vertex):
    the patches of the mesh the triangle emitter produces, so quad and triangle meshes have the same topology"""
    out = []
    for m in range(256):
        row = tri[m]
        n = row.index(-1) if -1 in row else 16
        tris = [row[i:i + 3] for i in range(0, n, 3)]
        par = list(range(len(tris)))

        def find(a):
            while par[a] != a:
                a = par[a]
            return a
        for i in range(len(tris)):
            for j in range(i):
                if set(tris[i]) & set(tris[j]):
                    par[find(i)] = find(j)
        comp = {}
        for i, t in enumerate(tris):
            r = find(i)
            comp.setdefault(r, 0)
            for e in t:
                comp[r] |= 1 << e
        # patch order: by lowest edge id, so the numbering is a function of the mask alone
        out.append(sorted(comp.values(), key=lambda v: (v & -v)))
    return out


def edge_corners(e):
    """corner bit indices (x high, z low) of the two endpoints of edge e"""
    if e < 4:  # X edge at (y,z) offset
        y, z = (e >> 1) & 1, e & 1
        return (y << 1) | z, 4 | (y << 1) | z
    if e < 8:  # Y edge at (x,z) offset
        x, z = ((e - 4) >> 1) & 1, (e - 4) & 1
        return (x << 2) | z, (x << 2) | 2 | z
    x, y = ((e - 8) >> 1) & 1, (e - 8) & 1  # Z edge at (x,y) offset
    return (x << 2) | (y << 1), (x << 2) | (y << 1) | 1


def main():
    edge_map, tri = parse()
    packed = []
    for m in range(256):
        crossed = 0
        for e in range(12):
            a, b = edge_corners(e)
            if ((m >> a) ^ (m >> b)) & 1:
                crossed |= 1 << e
        assert crossed == edge_map[m], (m, hex(crossed), hex(edge_map[m]))
        row = tri[m]
        n = row.index(-1) if -1 in row else 16
        assert n % 3 == 0 and n <= 15
        v = n << 60
        for i in range(n):
            assert (crossed >> row[i]) & 1, "table uses an uncrossed edge"
            v |= row[i] << (4 * i)
        packed.append(v)
    patches = tri_patches(tri)
    ref = ref_patches()
    differs = [m for m in range(256) if sorted(patches[m]) != ref[m]]
    assert differs == [126, 189, 219, 231], differs  # the two-diagonal-corners cases: EdgeTable joins them into one 6-edge patch
    ppack = []
    for m in range(256):
        crossed = 0
        for e in range(12):
            a, b = edge_corners(e)
            if ((m >> a) ^ (m >> b)) & 1:
                crossed |= 1 << e
        masks = patches[m]
        assert len(masks) <= 4
        union = 0
        for pm in masks:
            assert pm and not (union & pm), "patches of one cell must be disjoint"
            union |= pm
        assert union == crossed, (m, hex(union), hex(crossed))
        v = len(masks) << 60
        for i, pm in enumerate(masks):
            v |= pm << (12 * i)
        ppack.append(v)
    pbody = ",\n".join("\t" + ", ".join("0x%016xull" % v for v in ppack[i:i + 4]) for i in range(0, 256, 4))
    body = ",\n".join("\t" + ", ".join("0x%016xull" % v for v in packed[i:i + 4]) for i in range(0, 256, 4))
    text = ("/* GENERATED by tools/gen_mc_tables.py from MarchingCubes::tri_table (reference MCTable.h:25-283).\n"
            " * tri_pack[mask8]: nibble i (<15) = edge id of the i-th emitted index, nibble 15 = index count.\n"
            " * Edge ids: 0-3 X-edges (y,z offsets 00,01,10,11), 4-7 Y-edges (x,z), 8-11 Z-edges (x,y). */\n"
            "#define %s_TRI_PACK_INIT { \\\n%s }\n"
            "/* patch_pack[mask8]: the surface patches of a cell (one dual vertex per patch, Nielson's dual marching cubes) = the\n"
            " * connected components of its tri_table triangles; identical to Tables::EdgeTable (reference Tables.hpp:96-354) except\n"
            " * for masks 126, 189, 219, 231, where EdgeTable joins the two triangles into one 6-edge patch (and NumVertices[62] is\n"
            " * a typo there).  Bits 12p..12p+11 = the edges of patch p (ordered by lowest edge id), bits 60..63 = number of patches\n"
            " * (0..4).  The patches of a cell are disjoint and together cover exactly its sign-changing edges. */\n"
            "#define %s_PATCH_PACK_INIT { \\\n%s }\n")
    for path, prefix in (("binarymeshfitting_b200/csrc/mc_tables.h", "BMF"), ("oracle/mc_tables_oracle.h", "ORACLE")):
        with open(path, "w") as f:
            f.write("#pragma once\n" + text % (prefix, body.replace("\n", " \\\n"), prefix, pbody.replace("\n", " \\\n")))
    print("ok: 256 cases, max indices", max(v >> 60 for v in packed))


if __name__ == "__main__":
    sys.exit(main())
