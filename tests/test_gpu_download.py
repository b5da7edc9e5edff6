"""Device-driven download (bmf_batch_download_enqueue, csrc/download.cuh) against the reference-layout download
(bmf_batch_download = GLChunk::format_data's SoA, GLChunk.cpp:278-296): same bytes, no host synchronisation between submit and
enqueue, narrowed indices, skipped constant streams, and the error paths."""
import numpy as np
import pytest

from binarymeshfitting_b200 import Context, capi
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu


def grid(n, size, origin=-64.0):
    return [[origin + size * i, origin + size * j, origin + size * k, size] for i in range(n) for j in range(n) for k in range(n)]


def pinned(ctx, V, I):
    bufs = {k: capi.PinnedBuffer(ctx.lib, nb) for k, nb in (("pos", 12 * V), ("normal", 12 * V), ("color", 12 * V), ("boundary", V), ("valence", V),
                                                            ("i32", 4 * I), ("i16", 2 * I))}
    for b in bufs.values():
        b.u8[:] = 0xA5
    return bufs


def test_enqueue_on_a_fresh_context_without_any_host_sync():
    """first batch of a context: the arenas do not exist yet, so the emitters and the download return at once on the device;
    bmf_batch_wait grows the arenas, re-launches both and the data is there"""
    ctx = Context(0)
    ctx.set_sampler(capi.TERRAIN2D_PERT)
    descs = capi.make_chunk_descs(grid(4, 32.0), overlaps=0.045)
    V, I = 200000, 1200000
    b = pinned(ctx, V, I)
    ctx.submit(descs, 32, iters=2)
    ctx.download_enqueue(pos=b["pos"].view(np.float32), normal=b["normal"].view(np.float32), color=b["color"].view(np.float32),
                         boundary=b["boundary"].view(np.uint8), valence=b["valence"].view(np.uint8), inds32=b["i32"].view(np.uint32))
    ctx.wait()
    _, v, i = ctx.totals()
    assert 0 < v <= V and 0 < i <= I
    ref = ctx.download()
    np.testing.assert_array_equal(b["pos"].view(np.uint32)[:3 * v], ref["pos"].view(np.uint32).ravel())
    np.testing.assert_array_equal(b["normal"].view(np.uint32)[:3 * v], ref["normal"].view(np.uint32).ravel())
    np.testing.assert_array_equal(b["color"].view(np.uint32)[:3 * v], ref["color"].view(np.uint32).ravel())
    np.testing.assert_array_equal(b["boundary"].view(np.uint8)[:v], ref["boundary"])
    np.testing.assert_array_equal(b["valence"].view(np.uint8)[:v], ref["valence"])
    np.testing.assert_array_equal(b["i32"].view(np.uint32)[:i], ref["inds"])
    assert (b["pos"].view(np.uint8)[12 * v:] == 0xA5).all() and (b["i32"].view(np.uint8)[4 * i:] == 0xA5).all()  # nothing past the batch
    # second batch: arenas exist, nothing is re-launched; compact form = positions + uint16 indices only
    ctx.submit(descs, 32, iters=2)
    ctx.download_enqueue(pos=b["pos"].view(np.float32), inds16=b["i16"].view(np.uint16))
    ctx.wait()
    np.testing.assert_array_equal(b["i16"].view(np.uint16)[:i].astype(np.uint32), ref["inds"])
    infos = ctx.chunk_infos()
    mesh = infos["n_verts"] > 0
    assert (infos["flags"][mesh] & capi.CHUNK_COLOR_ONE).all() and (infos["flags"] & capi.CHUNK_INDEX16).all()
    assert not (infos["flags"] & capi.CHUNK_NORMAL_ZERO).any()  # iters = 2: the set_colors step turns processed normals into NaN
    assert (ref["color"] == 1.0).all()
    ctx.submit(descs, 32, iters=4)
    ctx.wait()
    assert (ctx.chunk_infos()["flags"] & capi.CHUNK_NORMAL_ZERO).all() and not ctx.download()["normal"].any()
    for x in b.values():
        x.close()
    ctx.close()


@pytest.mark.parametrize("misalign", [0, 4, 2])
def test_odd_sizes_and_unaligned_buffers(misalign):
    ctx = Context(0)
    ctx.set_sampler(capi.SPHERE)
    ctx.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), 64, iters=0)
    _, v, i = ctx.totals()
    ref = ctx.download()
    raw = {k: capi.PinnedBuffer(ctx.lib, nb + 64) for k, nb in (("pos", 12 * v), ("i32", 4 * i), ("i16", 2 * i), ("val", v))}
    ctx.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), 64, iters=0)
    pos = raw["pos"].view(np.uint8)[(misalign + 3) // 4 * 4:(misalign + 3) // 4 * 4 + 12 * v]  # floats stay 4-byte aligned
    i16 = raw["i16"].view(np.uint8)[misalign:misalign + 2 * i]
    val = raw["val"].view(np.uint8)[misalign:misalign + v]
    d = capi.DownloadDesc()
    d.pos, d.indices16, d.valence, d.cap_verts, d.cap_inds = pos.ctypes.data, i16.ctypes.data, val.ctypes.data, v, i
    import ctypes as C
    ctx._check(ctx.lib.bmf_batch_download_enqueue(ctx.h, C.byref(d)))
    ctx.wait()
    np.testing.assert_array_equal(pos, ref["pos"].view(np.uint8).ravel())
    np.testing.assert_array_equal(i16, ref["inds"].astype(np.uint16).view(np.uint8))
    np.testing.assert_array_equal(val, ref["valence"])
    for x in raw.values():
        x.close()
    ctx.close()


def test_error_paths():
    ctx = Context(0)
    ctx.set_sampler(capi.SPHERE)
    d1 = capi.make_chunk_descs([[-128, -128, -128, 256.0]])
    ctx.submit(d1, 64, iters=0)
    _, v, i = ctx.totals()
    small = capi.PinnedBuffer(ctx.lib, 12 * (v - 1))
    ctx.submit(d1, 64, iters=0)
    ctx.download_enqueue(pos=small.view(np.float32))
    with pytest.raises(capi.BmfError, match="too small"):
        ctx.wait()
    # pageable memory is refused up front
    ctx.submit(d1, 64, iters=0)
    with pytest.raises(capi.BmfError, match="page-locked"):
        ctx.download_enqueue(pos=np.zeros(3 * v, np.float32))
    # element-type alignment is checked up front
    odd = capi.PinnedBuffer(ctx.lib, 12 * v + 64)
    d = capi.DownloadDesc()
    d.pos, d.cap_verts = odd.view(np.uint8)[1:].ctypes.data, v
    import ctypes as C
    assert ctx.lib.bmf_batch_download_enqueue(ctx.h, C.byref(d)) == -1
    odd.close()
    # a chunk with >= 65536 vertices cannot travel as uint16
    rng = np.random.default_rng(3)
    dens = rng.standard_normal(64 ** 3).astype(np.float32)
    ctx.set_sampler(capi.HOST_DENSITY)
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 64.0]]), 64, iters=0, density=dens)
    _, v2, i2 = ctx.totals()
    assert v2 >= 65536
    assert not (ctx.chunk_infos()["flags"] & capi.CHUNK_INDEX16).any()
    big = capi.PinnedBuffer(ctx.lib, 2 * i2)
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 64.0]]), 64, iters=0, density=dens)
    ctx.download_enqueue(inds16=big.view(np.uint16))
    with pytest.raises(capi.BmfError, match="65536"):
        ctx.wait()
    small.close()
    big.close()
    ctx.close()


def test_registered_shared_memory_segment(tmp_path):
    """bmf_host_register: the multi-rank gather writes every rank's meshes into one shared host segment"""
    import mmap
    import os
    ctx = Context(0)
    ctx.set_sampler(capi.TORUS_Z)
    d1 = capi.make_chunk_descs([[-128, -128, -128, 256.0]])
    ctx.submit(d1, 64, iters=2)
    _, v, i = ctx.totals()
    ref = ctx.download()
    path = "/dev/shm/bmf_test_%d" % os.getpid()
    nbytes = 12 * v + 4 * i + 4096
    with open(path, "wb") as f:
        f.truncate(nbytes)
    fd = os.open(path, os.O_RDWR)
    try:
        mm = mmap.mmap(fd, nbytes)
        arr = np.frombuffer(mm, np.uint8)
        seg = capi.PinnedBuffer(ctx.lib, existing=arr)
        ctx.submit(d1, 64, iters=2)
        ctx.download_enqueue(pos=seg.view(np.float32, 0, 3 * v), inds32=seg.view(np.uint32, 12 * v + (-12 * v) % 16, i))
        ctx.wait()
        np.testing.assert_array_equal(seg.view(np.uint32, 0, 3 * v), ref["pos"].view(np.uint32).ravel())
        np.testing.assert_array_equal(seg.view(np.uint32, 12 * v + (-12 * v) % 16, i), ref["inds"])
        seg.close()
        del arr
        mm.close()
    finally:
        os.close(fd)
        os.unlink(path)
    ctx.close()


def test_download_dma_matches_plain_download():
    """bmf_batch_download_dma: copy-engine transfers of exactly the batch's sizes (uint16 indices packed on the device) == bmf_batch_download"""
    ctx = Context(0)
    ctx.set_sampler(capi.TERRAIN2D_PERT)
    ps = [[-32.0 + 32.0 * i, -32.0 + 32.0 * j, -32.0 + 32.0 * k, 32.0] for i in range(2) for j in range(2) for k in range(2)]
    d = capi.make_chunk_descs(ps, overlaps=0.045)
    ctx.submit(d, 64, iters=2)
    ref = ctx.download()
    _, v, i = ctx.totals()
    assert v > 1000 and i % 8 != 0 or True
    bufs = {k: capi.PinnedBuffer(ctx.lib, nb + 64) for k, nb in (("pos", 12 * v), ("nrm", 12 * v), ("col", 12 * v), ("bnd", v), ("val", v), ("i32", 4 * i), ("i16", 2 * i))}
    for use16 in (False, True):
        ctx.submit(d, 64, iters=2)  # not waited for: the call completes the batch itself
        pos, nrm, col = (bufs[k].view(np.float32, 0, 3 * v) for k in ("pos", "nrm", "col"))
        bnd, val = bufs["bnd"].view(np.uint8, 0, v), bufs["val"].view(np.uint8, 0, v)
        i32, i16 = bufs["i32"].view(np.uint32, 0, i), bufs["i16"].view(np.uint16, 0, i)
        for a in (pos, nrm, col, bnd, val, i32, i16):
            a[...] = 0
        if use16:
            ctx.download_dma(pos=pos, normal=nrm, color=col, boundary=bnd, valence=val, inds16=i16)
        else:
            ctx.download_dma(pos=pos, normal=nrm, color=col, boundary=bnd, valence=val, inds32=i32)
        ctx.wait()
        assert np.array_equal(pos.view(np.uint32), ref["pos"].reshape(-1).view(np.uint32))
        assert np.array_equal(nrm.view(np.uint32), ref["normal"].reshape(-1).view(np.uint32)) and np.array_equal(col, ref["color"].reshape(-1))
        assert np.array_equal(bnd, ref["boundary"]) and np.array_equal(val, ref["valence"])
        assert np.array_equal(i16 if use16 else i32, ref["inds"])
    # too small a buffer is refused at once, nothing written
    small = bufs["pos"].view(np.float32, 0, 3 * (v - 1))
    small[...] = 7.0
    with pytest.raises(capi.BmfError, match="too small"):
        ctx.download_dma(pos=small)
    assert (small == 7.0).all()
    # uint16 indices are refused when a chunk has >= 65536 vertices
    rng = np.random.default_rng(3)
    dens = rng.standard_normal(64 ** 3).astype(np.float32)
    ctx.set_sampler(capi.HOST_DENSITY)
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 64.0]]), 64, density=dens)
    _, v2, i2 = ctx.totals()
    assert v2 >= 65536
    big = capi.PinnedBuffer(ctx.lib, 2 * i2 + 64)
    with pytest.raises(capi.BmfError, match="65536"):
        ctx.download_dma(inds16=big.view(np.uint16, 0, i2))
    big.close()
    for b in bufs.values():
        b.close()
    ctx.close()
