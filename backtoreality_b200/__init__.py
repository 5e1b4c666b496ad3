"""backtoreality_b200 -- B200-native (sm_100a) PointNet++ set-abstraction hot path.

Drop-in for the `pointnet2` extension + modules of wyf-ACCEPT/BackToReality (VoteNet and
GroupFree3D): see DESIGN.md for the scope and INTEGRATION.md for the binding.
Importing this package does not load CUDA code; the first op call loads libb2r.so and raises if
it has not been built (there is no fallback path).
"""
__version__ = "0.1.0"
