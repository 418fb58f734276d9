"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/facetconv_b200.h declares, and refuses to compute without a device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "facetconv_b200.h")).read()
    return sorted(set(re.findall(r"FGC_API\s+[\w\s\*]+?\b(fgc_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    assert len(syms) >= 35
    for must in ("fgc_conv_fwd", "fgc_conv_bwd", "fgc_build_reverse_adj", "fgc_pool_max", "fgc_upsample",
                 "fgc_mlp_head_fwd", "fgc_normalize_rows", "fgc_vertex_update_edges", "fgc_vertex_update_ms",
                 "fgc_conv_fwd_host", "fgc_conv_fwd_bwd_host", "fgc_conv_fwd_planned", "fgc_conv_bwd_planned",
                 "fgc_build_conv_plan", "fgc_build_faces_adj", "fgc_build_edge_maps", "fgc_face_features",
                 "fgc_normalize_rows_segmented", "fgc_mlp_head_workspace"):
        assert must in syms


def test_library_builds_loads_and_exports_every_symbol():
    from facet_graph_convolution_b200.build import build
    from facet_graph_convolution_b200 import _lib
    path = build()
    h = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(h, s), "missing export %s" % s
    # the ctypes signature table covers the whole header too
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)
    assert _lib.lib().fgc_version() >= 100


def test_no_cpu_fallback():
    import torch
    from facet_graph_convolution_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().fgc_device_count() == 0
    x = torch.zeros(1, 4, 6)
    adj = torch.zeros(1, 4, 3, dtype=torch.int32)
    with pytest.raises(_lib.FacetConvError):
        ops.conv_fwd(x, adj, torch.zeros(2, 5, 6), torch.zeros(5), torch.zeros(2, 6), torch.zeros(2, 6), torch.zeros(2))
    with pytest.raises(_lib.FacetConvError):
        _lib.require_device()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "facet_graph_convolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_reference_arm_of_bench_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times beside the B200 arm) needs no GPU: bounded steps
    of the oracle port of the reference network on one patch of the C3 mesh (default), or of the reference-order
    layer port (--config c2); one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    for extra, metric in ((["--grid", "200", "--block", "40"], "facets/sec (denoise inference, fp32)"),
                          (["--config", "c2", "--cpu-sample", "2000"], "facets/sec")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                            "--warmup", "1"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["metric"] == metric and line["value"] > 0
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
        assert line["e2e"] == {"value": line["value"], "unit": "facets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["gpu_launches"] == 0
        assert "workload" in line["config"]
