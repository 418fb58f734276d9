"""Parity on the benchmarked workload itself: patches of bench.py's C3 mesh (1000x1000-quad height field cut into 100x100-quad
blocks + 3-quad halo, 22 464 level-0 rows each, the network's random-init parameters of bench.net_params) through the public
inference API against the oracle's fp64 closed form of the reference network (Code/model.py:837-946 + utils.py:1700-1715):
normals max-abs <= 1e-4, mean angular difference <= 0.01 degrees (BASELINE.json north_star); and the size-independent
property the 100-patch launches rely on: a patch inside a stacked launch gets bit for bit the rows it gets alone."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _forward(fm, ops, store, x, adjs, counts):
    with torch.no_grad(), fm.variable_store(store):
        y = fm.get_model_reg_multi_scale(x, adjs, 1.0)
        return ops.normalize_rows_segmented(y, counts)


def test_c3_patches_match_the_closed_form_and_batches_reproduce_single_patches():
    import bench
    from facet_graph_convolution_b200 import model as fm, ops, patches
    dev = torch.device("cuda:0")
    ids = [0, 37, 99]                                   # a corner, an interior and the last block of the 10 x 10 deal
    mine, _ = bench.make_c3_patches(1000, 100, ids)
    params = bench.net_params()
    store = fm.VariableStore(dev, params=params)
    pd = cf.split_net_params(params)

    def dev_group(group):
        xb, ab = patches.batch_patches(mine, group)
        cnt = torch.tensor([mine[i].x.shape[0] for i in group], dtype=torch.int32, device=dev)
        return torch.from_numpy(xb).to(dev), [torch.from_numpy(a).to(dev) for a in ab], cnt

    stacked = _forward(fm, ops, store, *dev_group([0, 1, 2])).cpu().numpy()
    for k in range(3):
        alone = _forward(fm, ops, store, *dev_group([k])).cpu().numpy()[0]
        n = mine[k].x.shape[0]
        assert np.array_equal(stacked[k, :n], alone[:n])            # bit for bit
        if k == 1:
            p = mine[k]
            ref = cf.normalize_tensor(cf.net_forward(p.x[None].astype(np.float64), [a[None] for a in p.adjs], pd))[0]
            real = np.arange(n) < p.num_real
            err = np.abs(alone[:n][real] - ref[real]).max()
            ang = np.degrees(np.arccos(np.clip((alone[:n][real] * ref[real]).sum(1) /
                                               (np.linalg.norm(alone[:n][real], axis=1) * np.linalg.norm(ref[real], axis=1) + 1e-30),
                                               -1, 1))).mean()
            assert err <= 1e-4, err
            assert ang <= 0.01, ang
