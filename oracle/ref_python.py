"""Import the reference's UNMODIFIED Python stack on CPU, over the oracle `_ext` facade.

TEST INFRASTRUCTURE ONLY, and BUILD-CONTAINER ONLY: /root/reference does not exist on the
GPU box, so nothing in the `-m gpu` tests, smoke() or bench.py may call this.  It is used by
tests/golden/make_golden.py (fixture generation) and by `-m "not gpu"` tests that skip when the
reference tree is absent.

The reference does `import pointnet2._ext as _ext` (pointnet2_utils.py:25-33) and bare
`import pointnet2_utils`, `import pytorch_utils` (pointnet2_modules.py:16-22), so we inject a
synthetic `pointnet2` package whose `_ext` attribute is oracle.fake_ext, and load the reference
files under private module names (so they never shadow this repo's own modules).
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("B2R_REFERENCE_ROOT", "/root/reference")
_V = os.path.join(REF_ROOT, "detection", "Votenet")
_G = os.path.join(REF_ROOT, "detection", "GroupFree3D")


def available(ref_root=None):
    v = _V if ref_root is None else os.path.join(ref_root, "detection", "Votenet")
    return os.path.isfile(os.path.join(v, "pointnet2", "pointnet2_utils.py"))


def _load(name, path, aliases=()):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    for a in aliases:
        sys.modules[a] = mod
    spec.loader.exec_module(mod)
    return mod


class RefStack:
    """Holds the reference modules for one flavour ('votenet' or 'groupfree3d')."""

    def __init__(self, flavour="votenet", ref_root=None, ext=None):
        """ref_root: a tree laid out like the reference (default /root/reference; the GPU tests
        pass the vendored copy under baseline/_ref).  ext: the module to serve as `pointnet2._ext`
        (default: the CPU oracle facade; the GPU tests pass the product's shim, i.e. the
        reference's unmodified Python then runs on libb2r.so -- INTEGRATION.md option A)."""
        if ext is None:
            from . import fake_ext
        else:
            fake_ext = ext

        if not available(ref_root):
            raise RuntimeError("reference tree not present at %s" % (ref_root or REF_ROOT))
        base = REF_ROOT if ref_root is None else ref_root
        root = os.path.join(base, "detection", "Votenet" if flavour == "votenet" else "GroupFree3D")
        saved = {k: sys.modules.get(k) for k in
                 ("pointnet2", "pointnet2._ext", "pointnet2_utils", "pytorch_utils",
                  "pointnet2_modules")}
        saved_path = list(sys.path)
        try:
            pkg = types.ModuleType("pointnet2")
            pkg.__path__ = []
            pkg._ext = fake_ext
            sys.modules["pointnet2"] = pkg
            sys.modules["pointnet2._ext"] = fake_ext
            tag = "_b2r_ref_%s_%s_" % (flavour, "cpu" if ext is None else "gpu")
            self.pytorch_utils = _load(tag + "pytorch_utils",
                                       os.path.join(root, "pointnet2", "pytorch_utils.py"),
                                       aliases=("pytorch_utils",))
            self.pointnet2_utils = _load(tag + "pointnet2_utils",
                                         os.path.join(root, "pointnet2", "pointnet2_utils.py"),
                                         aliases=("pointnet2_utils",))
            self.pointnet2_modules = _load(tag + "pointnet2_modules",
                                           os.path.join(root, "pointnet2", "pointnet2_modules.py"),
                                           aliases=("pointnet2_modules",))
            self.backbone_module = _load(tag + "backbone_module",
                                         os.path.join(root, "models", "backbone_module.py"))
            if flavour == "groupfree3d":
                # query sampling between backbone and decoder; the file does
                # sys.path.append + `import pointnet2_utils`, served by the alias above
                # (absent from the vendored copy under baseline/_ref the GPU tests use)
                gfm = os.path.join(root, "models", "modules.py")
                self.gf_modules = _load(tag + "gf_modules", gfm) if os.path.isfile(gfm) else None
            if flavour == "votenet":
                self.voting_module = _load(tag + "voting_module",
                                           os.path.join(root, "models", "voting_module.py"))
                self.proposal_module = _load(tag + "proposal_module",
                                             os.path.join(root, "models", "proposal_module.py"))
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
            sys.path[:] = saved_path


class cuda_is_identity:
    """The reference hard-codes `.cuda()` in a few places on this path (decode_scores,
    proposal_module.py:40; Pointnet2Backbone_jitter.forward, backbone_module.py:260): inside this
    context Tensor.cuda() returns the tensor itself so those lines run on the CPU oracle."""

    def __enter__(self):
        import torch
        self._saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._saved
