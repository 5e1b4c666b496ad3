"""Vendor the handful of reference PYTHON files that INTEGRATION.md option A runs unmodified on
top of libb2r.so into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun; SURVEY.md 7
step 0).  Build-container only: it reads /root/reference.  Nothing here is product source -- the
files are the reference's own, byte for byte, used by tests/test_option_a_gpu.py as the caller
of the drop-in boundary.

    python scripts/vendor_reference.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B2R_REFERENCE_ROOT", "/root/reference")
FILES = [
    "detection/Votenet/pointnet2/pointnet2_utils.py",
    "detection/Votenet/pointnet2/pointnet2_modules.py",
    "detection/Votenet/pointnet2/pytorch_utils.py",
    "detection/Votenet/models/backbone_module.py",
    "detection/Votenet/models/voting_module.py",
    "detection/Votenet/models/proposal_module.py",
    "detection/GroupFree3D/pointnet2/pointnet2_utils.py",
    "detection/GroupFree3D/pointnet2/pointnet2_modules.py",
    "detection/GroupFree3D/pointnet2/pytorch_utils.py",
    "detection/GroupFree3D/models/backbone_module.py",
]


def vendor(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("no reference tree at %s: nothing vendored" % REF)
        return False
    for rel in FILES:
        dst = os.path.join(ROOT, "baseline", "_ref", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        if verbose:
            print("vendored", rel)
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
