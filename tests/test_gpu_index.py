"""GPU index builders against the reference's own outputs: bit-exact.
tests/golden/index_layouts.npz holds adjacency, vertex-face and edge maps produced by the reference's
getFacesLargeAdj / getVerticesFaces / getEdgeMap (Code/utils.py:243-295, 370-395, 91-183) run unmodified
(oracle/make_golden.py); larger meshes are checked against the NumPy restatement the fixture pins."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def test_builders_reproduce_reference_fixtures_bit_for_bit():
    from facet_graph_convolution_b200 import ops
    g = golden("index_layouts")
    for tag in ("ico2", "torus", "open"):
        F = g[tag + "_F"]
        adj16, vf = ops.build_faces_adj(T(F), K=16, kv=25)
        adj10, _ = ops.build_faces_adj(T(F), K=10)          # K = 10 truncates rows (reference :280-291)
        assert np.array_equal(adj16.cpu().numpy(), g[tag + "_adj16"]), tag
        assert np.array_equal(adj10.cpu().numpy(), g[tag + "_adj10"]), tag
        assert np.array_equal(vf.cpu().numpy(), g[tag + "_vf"]), tag
        e_map, v_e = ops.build_edge_maps(T(F), 20)
        assert np.array_equal(e_map.cpu().numpy(), g[tag + "_emap"]), tag
        assert np.array_equal(v_e.cpu().numpy(), g[tag + "_vemap"]), tag


def test_builders_match_numpy_restatement_on_larger_meshes():
    from facet_graph_convolution_b200 import mesh, ops
    cases = [mesh.icosphere(4)[1], mesh.grid_mesh(40, 24, torus=True, morton=True)[1],
             mesh.grid_mesh(37, 21, torus=False, morton=False)[1]]
    for F in cases:
        F = np.asarray(F, np.int32)
        for K in (8, 16, 23):
            adj, _ = ops.build_faces_adj(T(F), K=K)
            assert np.array_equal(adj.cpu().numpy(), mesh.faces_large_adj(F, K)), (F.shape, K)
        _, vf = ops.build_faces_adj(T(F), kv=25)
        assert np.array_equal(vf.cpu().numpy(), mesh.vertex_faces(F, 25))
        e_ref, v_ref = mesh.edge_maps(F, 20)
        e_map, v_e = ops.build_edge_maps(T(F), 20)
        assert np.array_equal(e_map.cpu().numpy(), e_ref) and np.array_equal(v_e.cpu().numpy(), v_ref)


def test_fake_rows_and_errors():
    from facet_graph_convolution_b200 import _lib, mesh, ops
    F = np.asarray(mesh.icosphere(2)[1], np.int32)
    Fp = np.concatenate([F, np.full((16, 3), -1, np.int32)])          # fake nodes appended by the pyramid padding
    _, vf = ops.build_faces_adj(T(Fp), kv=25, nv=int(F.max()) + 1)
    assert np.array_equal(vf.cpu().numpy(), mesh.vertex_faces(Fp, 25, nv=int(F.max()) + 1))
    with pytest.raises(_lib.FacetConvError):
        ops.build_faces_adj(T(F), kv=2)                                # a vertex has more than 2 faces
    with pytest.raises(_lib.FacetConvError):
        ops.build_faces_adj(T(F), K=16, nv=5)                          # vertex ids outside 0..nv-1


def test_face_features_match_float64_reference_arithmetic():
    """normal | barycentre/diag rows: the reference computes them in float64 NumPy; tolerance 1e-6 on O(1) values."""
    from facet_graph_convolution_b200 import mesh, ops
    V, F = mesh.icosphere(3)
    V = mesh.add_vertex_noise(V, F, 0.3, 1).astype(np.float32)
    F = np.asarray(F, np.int32)
    got = ops.face_features(T(V), T(F)).cpu().numpy()
    ref = mesh.face_features(V.astype(np.float64), F)
    assert np.abs(got - ref).max() < 1e-6
    Fp = np.concatenate([F, np.full((4, 3), -1, np.int32)])
    gp = ops.face_features(T(V), T(Fp)).cpu().numpy()
    assert np.array_equal(gp[: F.shape[0]], got) and not gp[F.shape[0]:].any()
