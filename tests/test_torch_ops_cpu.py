"""The `fgc::` torch.library ops exist, carry schemas and infer shapes on fake tensors without a GPU."""
import pytest
import torch


def test_ops_registered_with_schemas():
    from facet_graph_convolution_b200 import torch_ops
    torch_ops.register()
    torch_ops.register()          # idempotent
    for name in torch_ops.OP_NAMES:
        op = getattr(torch.ops.fgc, name)
        assert "fgc::" + name in str(op.default._schema)
    s = str(torch.ops.fgc.conv_fwd.default._schema)
    assert "Tensor x" in s and "Tensor adj" in s and "bool bias_mask" in s and "-> Tensor" in s


def test_shape_inference_on_fake_tensors():
    from torch._subclasses.fake_tensor import FakeTensorMode
    from facet_graph_convolution_b200 import torch_ops
    torch_ops.register()
    B, N, K, Cin, Cout, M = 2, 64, 16, 32, 64, 9
    with FakeTensorMode():
        x = torch.empty(B, N, Cin)
        adj = torch.empty(B, N, K, dtype=torch.int32)
        W0, b = torch.empty(M, Cout, Cin), torch.empty(Cout)
        u, v, c = torch.empty(M, Cin), torch.empty(M, Cin), torch.empty(M)
        y = torch.ops.fgc.conv_fwd(x, adj, W0, b, u, v, c, True, 0, 0.1)
        assert tuple(y.shape) == (B, N, Cout)
        g = torch.ops.fgc.conv_bwd(y, x, adj, W0, u, v, c, True)
        assert [tuple(t.shape) for t in g] == [(B, N, Cin), (M, Cout, Cin), (Cout,), (M, Cin), (M, Cin), (M,)]
        assert tuple(torch.ops.fgc.pool_max(y, 4).shape) == (B, N // 4, Cout)
        assert tuple(torch.ops.fgc.upsample(y, 4).shape) == (B, 4 * N, Cout)
        xc = torch.empty(B, N // 4, Cin)
        assert tuple(torch.ops.fgc.conv_fwd_up(xc, adj, W0, b, u, v, c, 2, True, 0, 0.1).shape) == (B, N, Cout)
        assert tuple(torch.ops.fgc.normalize_rows(torch.empty(B, N, 3)).shape) == (B, N, 3)
        h = torch.ops.fgc.mlp_head(torch.empty(B, N, 32), torch.empty(32, 1024), torch.empty(1024), torch.empty(1024, 3),
                                   torch.empty(3), 0.1)
        assert tuple(h.shape) == (B, N, 3)
        assert tuple(torch.ops.fgc.gather_rows(x, adj).shape) == (B, N, K, Cin)
        p0, p1 = torch.empty(1, 50, 3), torch.empty(1, 40, 3)
        i0 = torch.empty(10, dtype=torch.int32)
        loss, gp0 = torch.ops.fgc.point_set_loss(p0, p1, i0, i0, 1)
        assert tuple(loss.shape) == (1,) and tuple(gp0.shape) == (1, 50, 3)
        xv, nr = torch.empty(30, 3), torch.empty(16, 3)
        fc, vf = torch.empty(64, 3, dtype=torch.int32), torch.empty(30, 25, dtype=torch.int32)
        assert tuple(torch.ops.fgc.vertex_update_ms(xv, nr, fc, vf, 1, 2, 20).shape) == (30, 3)
        gx, gn = torch.ops.fgc.vertex_update_ms_bwd(xv, xv, nr, fc, vf, 1, 2, 20)
        assert tuple(gx.shape) == (30, 3) and tuple(gn.shape) == (16, 3)


def test_cpu_tensors_fail_loudly():
    from facet_graph_convolution_b200 import torch_ops
    torch_ops.register()
    x = torch.zeros(1, 16, 32)
    adj = torch.zeros(1, 16, 4, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        torch.ops.fgc.conv_fwd(x, adj, torch.zeros(9, 32, 32), torch.zeros(32), torch.zeros(9, 32), torch.zeros(9, 32),
                               torch.zeros(9), True, 0, 0.1)


def test_group_index_is_a_stable_csr():
    """ops._group_index (the index lists of fgc_vertex_update_ms_bwd): positions grouped by key, ascending inside a
    group, invalid positions dropped -- against a plain Python grouping."""
    import numpy as np
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(2)
    keys = rs.randint(-1, 7, size=(40, 5)).astype(np.int32)
    valid = (keys >= 0) & (keys < 6)
    ptr, ids = ops._group_index(torch.from_numpy(keys), torch.from_numpy(valid), 6)
    ptr, ids = ptr.numpy(), ids.numpy()
    flat = keys.reshape(-1)
    assert ptr[0] == 0 and ptr[-1] == valid.sum() and ptr.dtype == np.int32 and ids.dtype == np.int32
    for k in range(6):
        want = [i for i in range(flat.size) if flat[i] == k]
        assert ids[ptr[k]:ptr[k + 1]].tolist() == want
