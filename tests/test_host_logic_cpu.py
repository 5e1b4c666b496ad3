"""Host-side logic that needs no GPU: the per-step zero arena, the geometry splitter of the BR step,
the scatter-plan dispatch rule, and the loud failure of CUDA-only pieces on CPU."""
import pytest
import torch

from backtoreality_b200 import step_arena
from backtoreality_b200.backbone_module import Pointnet2Backbone


def test_step_arena_hands_out_zeroed_views_and_falls_back():
    dev = torch.device("cpu")
    # inactive: plain torch.zeros
    a = step_arena.zeros((3, 4), torch.float32, dev)
    assert a.shape == (3, 4) and float(a.abs().sum()) == 0
    step_arena.begin(dev, min_bytes=4096)
    try:
        x = step_arena.zeros(100, torch.float64, dev)
        y = step_arena.zeros((10, 10), torch.float32, dev)
        assert x.dtype == torch.float64 and y.shape == (10, 10)
        assert x.data_ptr() % 256 == y.data_ptr() % 256      # 256-byte aligned carving
        assert x.data_ptr() != y.data_ptr()
        x.fill_(7.0); y.fill_(3.0)
        big = step_arena.zeros(1 << 20, torch.float32, dev)  # does not fit: plain fill this time
        assert float(big.abs().sum()) == 0
    finally:
        step_arena.end(dev)
    step_arena.begin(dev, min_bytes=4096)                    # next step: grown, cleared, rewound
    try:
        x2 = step_arena.zeros(100, torch.float64, dev)
        assert float(x2.abs().sum()) == 0
        big2 = step_arena.zeros(1 << 20, torch.float32, dev)
        assert float(big2.abs().sum()) == 0
    finally:
        step_arena.end(dev)
    z = step_arena.zeros_like(torch.ones(5))
    assert z.shape == (5,) and float(z.sum()) == 0


def test_split_geometry_slices_batches_and_routes_split_plans():
    hook = lambda: None
    lv = lambda np_: {"inds": torch.arange(4 * np_).view(4, np_), "new_xyz": torch.zeros(4, np_, 3),
                      "idx": torch.zeros(4, np_, 2, dtype=torch.int32), "event": None, "sm_limit": 7,
                      "cidx#0": torch.tensor([1]), "cidx#1": torch.tensor([2]),
                      "cmeta#0": torch.tensor([10]), "cmeta#1": torch.tensor([20])}
    levels = [lv(8), lv(4), lv(2), dict(lv(1), fp1_idx=torch.arange(8).view(4, 2), after_forward=hook)]
    halves = Pointnet2Backbone.split_geometry(levels, 2)
    assert len(halves) == 2 and len(halves[0]) == 4
    assert torch.equal(halves[1][0]["inds"], levels[0]["inds"][2:])
    assert halves[0][0]["new_xyz"].shape == (2, 8, 3) and halves[0][0]["sm_limit"] == 7
    assert int(halves[0][1]["cidx"]) == 1 and int(halves[1][1]["cidx"]) == 2
    assert int(halves[1][2]["cmeta"]) == 20 and "cidx#0" not in halves[0][0]
    assert torch.equal(halves[1][3]["fp1_idx"], torch.tensor([[4, 5], [6, 7]]))
    assert halves[0][3]["after_forward"] is hook and "after_forward" not in halves[1][3]


def test_scatter_plan_dispatch_rule():
    from backtoreality_b200 import _ext
    # SA2-shaped grouping (32768 entries over 2048 targets, 1024 rows): dense inverse index -> plan
    assert _ext._plan_pays(1024 * 32, 2048, 1, 8 * 128)
    # SA1-shaped (131072 entries over 40000 targets): mostly empty lists -> atomic scatter
    assert not _ext._plan_pays(2048 * 64, 40000, 1, 8 * 4)
    # too few rows to fill the GPU
    assert not _ext._plan_pays(1024 * 32, 2048, 1, 16)
    # more than 4096 targets with dense lists: the one-row-per-CTA plan kernel
    assert _ext._plan_pays(3 * 100000, 8192, 3, 8 * 64)


def test_flat_adam_refuses_cpu_parameters():
    from backtoreality_b200.flat_adam import FlatAdam
    with pytest.raises(RuntimeError, match="CUDA only"):
        FlatAdam([torch.nn.Parameter(torch.zeros(4))])
