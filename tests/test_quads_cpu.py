"""Quad emission, CPU side: properties of the oracle's dual-marching-cubes quad emitter (build-defined --
the reference has the MeshProcessor<4> consumer and the Nielson tables but nothing that emits a quad)."""
import numpy as np
import pytest

from oracle import oracle_binding as ob


def signed_volume(P, T):
    return np.einsum("ij,ij->i", P[T[:, 0]], np.cross(P[T[:, 1]], P[T[:, 2]])).sum() / 6.0


@pytest.mark.parametrize("kind,dim", [(ob.SPHERE, 32), (ob.SPHERE, 64), (ob.TORUS_Z, 64), (ob.CUBOID, 64)])
def test_quad_mesh_is_closed_and_wound_like_the_triangle_mesh(oracle, kind, dim):
    ch = oracle.chunk(oracle.sampler(kind), (-128, -128, -128), 256.0, dim)
    q = oracle.quads(ch["density"], ch["bits"], dim)
    I = q["inds"].reshape(-1, 4)
    assert q["n_verts"] > 0 and len(I) > 0 and I.max() < q["n_verts"]
    # closed and consistently oriented: every directed edge has its opposite exactly once
    e = np.concatenate([I[:, [0, 1]], I[:, [1, 2]], I[:, [2, 3]], I[:, [3, 0]]])
    fwd = {}
    for a, b in e:
        fwd[(a, b)] = fwd.get((a, b), 0) + 1
    assert all(c == 1 for c in fwd.values()) and all((b, a) in fwd for a, b in fwd)
    # genus from the Euler characteristic: sphere / cuboid 0, torus 1
    chi = q["n_verts"] - len(e) // 2 + len(I)
    assert chi == (0 if kind == ob.TORUS_Z else 2)
    # every vertex is used, valence = number of quad corners on it
    assert np.array_equal(np.bincount(I.ravel(), minlength=q["n_verts"]).astype(np.uint8), q["valence"])
    # same winding and (nearly) the same enclosed volume as the reference's triangle mesh of the same samples
    T = np.concatenate([I[:, [0, 1, 2]], I[:, [2, 3, 0]]])
    vq = signed_volume(q["pos"].astype(np.float64), T)
    vt = signed_volume(ch["pos"].astype(np.float64), ch["inds"].reshape(-1, 3).astype(np.int64))
    assert vq * vt > 0 and abs(vq - vt) < 0.01 * abs(vt)
    # far fewer primitives than triangles (README.md:72)
    assert len(I) * 2 <= len(ch["inds"]) // 3 + 8


def test_quads_then_meshprocessor4(oracle):
    ch = oracle.chunk(oracle.sampler(ob.SPHERE), (-128, -128, -128), 256.0, 32)
    q = oracle.quads(ch["density"], ch["bits"], 32)
    nv = q["n_verts"]
    p, c, n = oracle.smooth(q["pos"], np.ones((nv, 3), np.float32), np.zeros((nv, 3), np.float32), q["boundary"], q["valence"], q["inds"], 4, 3, False, False)
    r0 = np.linalg.norm(q["pos"] - q["pos"].mean(axis=0), axis=1)
    r1 = np.linalg.norm(p - p.mean(axis=0), axis=1)
    assert np.all(np.isfinite(p)) and not np.array_equal(p, q["pos"])
    assert abs(r1.mean() - r0.mean()) < 0.05 * r0.mean()  # the consumer accepts the emitter's output and keeps the shape
    assert np.all(c == 1.0)
