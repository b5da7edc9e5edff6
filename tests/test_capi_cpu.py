"""CPU: the C-ABI library loads and exports every symbol include/bmf_b200.h declares; argument validation and the
no-fallback rule (no compute call is made without a GPU)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import pytest

from binarymeshfitting_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "bmf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bmf_[a-z0-9_]+)\s*\(", txt)))


def test_header_functions_are_exported():
    lib = capi.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libbmf_b200.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with include/bmf_b200.h"
    assert b"sm_100a" in lib.bmf_version()


def test_struct_layouts_match_header():
    assert C.sizeof(capi.SamplerDesc) == 80 and C.sizeof(capi.ChunkDesc) == 32
    assert C.sizeof(capi.Params) == 36 and C.sizeof(capi.ChunkInfo) == 56 and C.sizeof(capi.DownloadDesc) == 72


def test_sampler_defaults_are_the_reference_world_defaults():
    lib = capi.load_library()
    s = capi.SamplerDesc()
    lib.bmf_sampler_defaults(C.byref(s), capi.TERRAIN2D_PERT)
    # WorldOctree.cpp:47-54
    assert (s.kind, s.world_size, s.octaves, s.seed) == (11, 256.0, 13, 1337)
    assert abs(s.g_scale - 0.25) < 1e-9 and abs(s.height - 75.0) < 1e-9 and abs(s.amp - 0.87) < 1e-7
    assert abs(s.frequency - 0.585) < 1e-7 and abs(s.gain - 0.488) < 1e-7


def test_sass_is_sm_100a_and_has_the_kernels():
    so = os.path.join(ROOT, "binarymeshfitting_b200", "libbmf_b200.so")
    r = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    r = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True)
    names = r.stdout
    for k in ("k_terrain2d_sheet", "k_terrain2d_bits", "k_terrain3d", "k_sample_implicit", "k_pack_density", "k_count", "k_scan_chunks", "k_bases", "k_verts3",
              "k_inds3", "k_valence_offsets", "k_adj_fill", "k_smooth_chunks", "k_dual", "k_primal", "k_qef_batch", "k_qef_place",
              "k_seam_layers", "k_seam_cull", "k_seam_classify", "k_seam_pass", "k_q_count", "k_q_bases", "k_q_emit", "k_format_unwind", "k_ubench_issue",
              "k_chunk_count", "k_chunk_emit", "k_publish", "k_pack_indices16", "k_download", "k_sampler_gradient", "k_dual_gradient", "k_color_map", "k_collapse_bad_quads"):
        assert k in names, k
    assert "UBLKCP" in names, "the per-chunk kernels stage their chunk with cp.async.bulk (TMA): its SASS mnemonic must be present"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BmfError):
        capi.Context(0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.BmfError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_reference_arm_contract_line():
    """bench.py --impl reference prints one JSON line with the contract keys (tiny workload, CPU only)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--chunks-per-axis", "4"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference"
    if "unavailable" in d:
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert k in d
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
