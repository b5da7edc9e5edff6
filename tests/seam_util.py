"""Helpers for the seam-pass tests: LOD worlds sampled at voxel-node centres and a watertightness check."""
import numpy as np
from scipy.spatial import cKDTree

from binarymeshfitting_b200 import world


def lod_world(max_level, min_level=1, focus=(0.0, 0.0, 0.0), dim=32):
    props = world.WorldProperties(max_level=max_level, min_level=min_level, chunk_resolution=dim)
    ps, lv, mc = world.split_leaves(props, 256, focus)
    return ps, lv, mc


def seam_overlap(dim):
    return np.float32(-0.5) / np.float32(dim)


def world_triangles(chunks, dim):
    """chunk meshes (grid units) -> world-space triangle soup [T,3,3] (what the renderer does with overlap_pos / scale)"""
    out = []
    for ch in chunks:
        if not ch.get("n_inds"):
            continue
        p = np.asarray(ch["overlap_pos"], np.float64)[None, :] + np.asarray(ch["pos"], np.float64).reshape(-1, 3) * float(ch["scale"])
        out.append(p[np.asarray(ch["inds"], np.int64)].reshape(-1, 3, 3))
    return np.concatenate(out) if out else np.zeros((0, 3, 3))


def edge_report(tris, eps, return_open=False):
    """weld corners closer than eps, drop collapsed triangles, count how often every directed edge is matched by its
    opposite.  Returns (n_triangles_kept, n_unmatched_directed_edges, n_nonmanifold_edges)."""
    pts = np.asarray(tris, np.float64).reshape(-1, 3)
    if len(pts) == 0:
        return 0, 0, 0
    tree = cKDTree(pts)
    pairs = tree.query_pairs(eps, output_type="ndarray")
    parent = np.arange(len(pts))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in pairs:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    ids = np.array([find(i) for i in range(len(pts))]).reshape(-1, 3)
    keep = (ids[:, 0] != ids[:, 1]) & (ids[:, 1] != ids[:, 2]) & (ids[:, 0] != ids[:, 2])
    ids = ids[keep]
    e = np.concatenate([ids[:, [0, 1]], ids[:, [1, 2]], ids[:, [2, 0]]])
    fwd = {}
    for a, b in e:
        fwd[(a, b)] = fwd.get((a, b), 0) + 1
    unmatched = sum(abs(c - fwd.get((b, a), 0)) for (a, b), c in fwd.items() if a < b or (b, a) not in fwd)
    nonmanifold = sum(1 for (a, b), c in fwd.items() if c > 1)
    if return_open:
        # positions of the end points of every directed edge without an opposite
        rep = {}
        for i, r in enumerate(np.array([find(i) for i in range(len(pts))])):
            rep.setdefault(int(r), pts[i])
        open_pts = [rep[int(a)] for (a, b), c in fwd.items() if fwd.get((b, a), 0) != c] + [rep[int(b)] for (a, b), c in fwd.items() if fwd.get((b, a), 0) != c]
        return int(keep.sum()), int(unmatched), int(nonmanifold), np.array(open_pts, np.float64).reshape(-1, 3)
    return int(keep.sum()), int(unmatched), int(nonmanifold)
