"""`pointnet2._ext` -- the nine functions the reference binds in _ext_src/src/bindings.cpp:11-24,
served by backtoreality_b200 (libb2r.so, sm_100a)."""
from backtoreality_b200._ext import (  # noqa: F401
    ball_query,
    furthest_point_sampling,
    gather_points,
    gather_points_grad,
    group_points,
    group_points_grad,
    three_interpolate,
    three_interpolate_grad,
    three_nn,
)
