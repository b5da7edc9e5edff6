"""GPU: Sampler::gradient on the device (bmf_sampler_gradient; implicit_gradient, ImplicitSampler.hpp:38-49) bit for bit against the
oracle -- which tests/test_oracle_vs_ref.py pins to the compiled reference -- and the gradient-fed QEF placement of config 5
(bmf_params.qef = 2: plane normals = normalize(sampler.gradient(dual_p)), the line commented out at MeshProcessor.cpp:224) against
its oracle twin.  The placement POLICY is build-defined (the reference never calls its solver); the gradient arithmetic is pinned."""
import os
import subprocess

import numpy as np
import pytest

from binarymeshfitting_b200 import capi
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CSG = dict(csg_op=ob.CSG_SUBTRACT, csg_kind_a=ob.SPHERE, csg_kind_b=ob.CUBOID, csg_world_size_a=256.0, csg_world_size_b=300.0,
           csg_offset_a=(0.0, 0.0, 0.0), csg_offset_b=(20.0, -10.0, 5.0))


def points(seed, m):
    rng = np.random.default_rng(seed)
    p = (rng.random((m, 3), dtype=np.float32) * 300 - 150).astype(np.float32)
    p[:8] = [[0, 0, 0], [64, 0, 0], [0, 32, 0], [0, 0, -32], [32, 32, 32], [1e-3, -1e-3, 0], [150, 150, 150], [-64, 0, 25.6]]  # centre, surfaces, corners
    return p


@pytest.mark.parametrize("kind", [ob.SPHERE, ob.TORUS_Z, ob.CUBOID, ob.PLANE_Y])
@pytest.mark.parametrize("h", [0.01, 0.001, 1.0])
def test_implicit_gradient_bit_exact(gpu, oracle, kind, h):
    gpu.set_sampler(kind)
    p = points(kind, 4000)
    g = gpu.sampler_gradient(p, h)
    o = np.stack([oracle.implicit_gradient(kind, q, h) for q in p])
    np.testing.assert_array_equal(g.view(np.uint32), o.view(np.uint32))
    assert np.abs(g).max() > 0


def test_csg_and_noise_sampler_gradients(gpu, oracle):
    p = points(9, 2000)
    gpu.set_sampler(ob.CSG, **CSG)
    np.testing.assert_array_equal(gpu.sampler_gradient(p).view(np.uint32), oracle.sampler_gradient(oracle.sampler(ob.CSG, **CSG), p).view(np.uint32))
    # the value callback of every noise sampler is NoiseSamplers::noise3d == 0 (NoiseSampler.cpp:99-102): gradient (0,0,0), like the reference's
    for kind in (ob.TERRAIN2D_PERT, ob.TERRAIN3D_PERT):
        gpu.set_sampler(kind)
        g = gpu.sampler_gradient(p)
        assert not g.any() and np.array_equal(g, oracle.sampler_gradient(oracle.sampler(kind), p))
    gpu.set_sampler(ob.HOST_DENSITY)
    with pytest.raises(capi.BmfError, match="HOST_DENSITY"):
        gpu.sampler_gradient(p)


@pytest.mark.parametrize("case", ["sphere64", "csg128"])
def test_gradient_fed_qef_matches_oracle_twin(gpu, oracle, case):
    """config 5: implicit CSG shapes with gradients -- QEF vertex placement from gradient normals, triangles"""
    if case == "sphere64":
        kind, kw, pos, size, dim, overlap, iters = ob.SPHERE, {}, (-128.0, -128.0, -128.0), 256.0, 64, 0.0, 2
    else:
        kind, kw, pos, size, dim, overlap, iters = ob.CSG, CSG, (-128.0, -128.0, -128.0), 256.0, 128, 0.055, 4
    gpu.set_sampler(kind, **kw)
    descs = capi.make_chunk_descs([[pos[0], pos[1], pos[2], size]], overlaps=overlap)
    gpu.submit(descs, dim, iters=iters, qef=2)
    gpu.wait()
    g = gpu.copy_chunk(0, want=("verts", "inds"))
    s = oracle.sampler(kind, **kw)
    o = oracle.chunk(s, pos, size, dim, overlap, iters=iters, qef=2)
    assert g["n_verts"] == o["n_verts"] > 1000 and np.array_equal(g["inds"], o["inds"])
    d = np.abs(g["verts"]["p"] - o["pos"])
    assert d.max() <= 1e-4 * (dim - 1), "gradient-QEF positions differ by %g grid units" % d.max()  # north_star bar: 1e-4 of chunk extent
    plain = oracle.chunk(s, pos, size, dim, overlap, iters=iters)["pos"]
    face = oracle.chunk(s, pos, size, dim, overlap, iters=iters, qef=1)["pos"]
    assert np.abs(o["pos"] - plain).max() > 1e-3 and np.abs(o["pos"] - face).max() > 1e-3  # a different placement from qef = 0 and qef = 1


def test_qef2_is_refused_for_noise_samplers(gpu):
    gpu.set_sampler(ob.TERRAIN2D_PERT)
    with pytest.raises(capi.BmfError, match="analytic"):
        gpu.submit(capi.make_chunk_descs([[0, 0, 0, 32.0]]), 32, iters=2, qef=2)
    with pytest.raises(capi.BmfError, match="qef must be"):
        gpu.submit(capi.make_chunk_descs([[0, 0, 0, 32.0]]), 32, iters=2, qef=3)


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 11])
def test_cpp_mirror_block_gradient_equals_host_callback(kind):
    """C++ mirror: sampler_gradient_block (device) against Sampler::gradient called per point on the host, as the reference would"""
    exe = os.path.join(ROOT, "binarymeshfitting_b200", "host", "host_test")
    r = subprocess.run([exe, "gradient", str(kind), "5000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "mismatches=0" in r.stdout, r.stdout
