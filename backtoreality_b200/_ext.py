"""Drop-in for the reference's pybind module `pointnet2._ext`.

Same nine function names, argument order, return types and precondition errors as
/root/reference/detection/Votenet/pointnet2/_ext_src/src/bindings.cpp:11-24 (wrappers in
sampling.cpp, ball_query.cpp, group_points.cpp, interpolate.cpp), but every op runs in
libb2r.so (hand-written sm_100a CUDA behind the C ABI of include/b2r.h).  The reference's own
`pointnet2_utils.py` works unchanged on top of this module (the repo-root
package `pointnet2/` re-exports it under the reference's import path).

Error behaviour mirrors `_ext_src/include/utils.h:10-30`: non-contiguous / wrong dtype / mixed
device inputs raise RuntimeError with the reference's messages; CPU tensors raise
"CPU not supported" (sampling.cpp:38-40).  A failing kernel launch raises instead of calling
exit(-1) (cuda_utils.h:35-44).
"""
import os

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---- instrumentation used by bench.py (never changes what runs) ------------------------------
LAUNCHES = 0   # number of libb2r kernel launches issued through this module
TIMED = {}     # op name -> list of (start_event, stop_event, algorithmic_bytes); only TIME_OPS
TIME_OPS = set()

# scenes with at least this many points take the grid-accelerated ball query (same output)
BALL_QUERY_GRID_MIN_N = 8192


class _timed:
    """Counts the launch and, for ops listed in TIME_OPS, brackets it with CUDA events recorded
    on the launching (current) stream."""

    def __init__(self, name, algo_bytes=0):
        self.name = name
        self.bytes = algo_bytes
        self.ev = None

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += 1
        if self.name in TIME_OPS:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *a):
        if self.ev is not None:
            self.ev[1].record()
            TIMED.setdefault(self.name, []).append((self.ev[0], self.ev[1], self.bytes))


def _chk_contig(t, name):
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def _chk_float(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)


def _chk_int(t, name):
    if t.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)


def _chk_cuda(lead, others):
    if not lead.is_cuda:
        raise RuntimeError("CPU not supported")
    for t, name in others:
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)
        if t.device != lead.device:
            raise RuntimeError("%s must be on the same device as the first argument" % name)


class _on_device:
    """Make the tensor's device current for the duration of the launch (the reference relies on
    the caller's current device; nn.DataParallel replicas set it per thread)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and self.idx != cur:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# B2R_FPS_LEGACY=1 keeps round 1's kernel (csrc/fps.cu: every point updated every iteration,
# warp -> CTA -> cluster reduction) for A/B measurements; the indices are identical either way
FPS_LEGACY = os.environ.get("B2R_FPS_LEGACY", "0") not in ("0", "")


def fps_presort(points):
    """The spatial sort of furthest_point_sampling (include/b2r.h: b2r_fps_sort) run ahead of time:
    returns the workspace to pass as `presorted=`.  Depends on the coordinates only."""
    _chk_contig(points, "points")
    _chk_float(points, "points")
    _chk_cuda(points, [])
    B, N = points.size(0), points.size(1)
    l = _lib.lib()
    nbytes = int(l.b2r_fps_workspace_bytes(B, N))
    ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=points.device)
    with _on_device(points):
        _lib.check(l.b2r_fps_sort(points.data_ptr(), B, N, ws.data_ptr(), nbytes, _stream()), "fps_presort")
    return ws


def furthest_point_sampling(points, nsamples, cluster=0, presorted=None):
    """(B,N,3) f32 -> (B,nsamples) i32.  Replaces sampling.cpp:70-91.  `cluster` (not in the
    reference's signature, default 0 = lowest latency) caps the CTAs one scene holds; the
    indices do not depend on it (include/b2r.h: b2r_fps_ws / b2r_fps_ex).  presorted: the
    workspace `fps_presort(points)` returned for the SAME points (skips the sort)."""
    _chk_contig(points, "points")
    _chk_float(points, "points")
    _chk_cuda(points, [])
    B, N = points.size(0), points.size(1)
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    with _on_device(points), _timed("furthest_point_sampling"):
        l = _lib.lib()
        if FPS_LEGACY:
            _lib.check(l.b2r_fps_ex(points.data_ptr(), B, N, int(nsamples), out.data_ptr(),
                                    int(cluster), _stream()), "furthest_point_sampling")
        else:
            nbytes = int(l.b2r_fps_workspace_bytes(B, N))
            fn = l.b2r_fps_ws
            if presorted is not None:
                ws, fn = presorted, l.b2r_fps_ws_presorted
                if ws.numel() < nbytes:
                    raise RuntimeError("furthest_point_sampling: presorted workspace too small")
            else:
                ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=points.device)
            _lib.check(fn(points.data_ptr(), B, N, int(nsamples), out.data_ptr(),
                          int(cluster), ws.data_ptr(), nbytes, _stream()),
                       "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """(B,C,N), (B,M) -> (B,C,M).  Replaces sampling.cpp:20-43."""
    _chk_contig(points, "points")
    _chk_contig(idx, "idx")
    _chk_float(points, "points")
    _chk_int(idx, "idx")
    _chk_cuda(points, [(idx, "idx")])
    B, C, N = points.shape
    M = idx.size(1)
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with _on_device(points), _timed("gather_points"):
        _lib.check(_lib.lib().b2r_gather_fwd(points.data_ptr(), idx.data_ptr(), B, C, N, M,
                                             out.data_ptr(), _stream()), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,M), (B,M), n -> (B,C,n).  Replaces sampling.cpp:45-69."""
    _chk_contig(grad_out, "grad_out")
    _chk_contig(idx, "idx")
    _chk_float(grad_out, "grad_out")
    _chk_int(idx, "idx")
    _chk_cuda(grad_out, [(idx, "idx")])
    B, C, M = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out), _timed("gather_points_grad"):
        _lib.check(_lib.lib().b2r_gather_bwd(grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), M,
                                             out.data_ptr(), _stream()), "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,M,3), (B,N,3) -> (B,M,nsample) i32.  Replaces ball_query.cpp:13-37."""
    _chk_contig(new_xyz, "new_xyz")
    _chk_contig(xyz, "xyz")
    _chk_float(new_xyz, "new_xyz")
    _chk_float(xyz, "xyz")
    _chk_cuda(new_xyz, [(xyz, "xyz")])
    B, M = new_xyz.size(0), new_xyz.size(1)
    N = xyz.size(1)
    out = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    with _on_device(new_xyz), _timed("ball_query"):
        if N >= BALL_QUERY_GRID_MIN_N:
            # large scenes: hashed uniform grid, 27 cells per centre instead of all N points
            # (identical output; b2r_ball_query_grid in include/b2r.h)
            nbytes = _lib.lib().b2r_ball_query_workspace_bytes(B, N)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=new_xyz.device)
            _lib.check(_lib.lib().b2r_ball_query_grid(new_xyz.data_ptr(), xyz.data_ptr(), B, N, M,
                                                      float(radius), int(nsample), out.data_ptr(),
                                                      ws.data_ptr(), nbytes, _stream()),
                       "ball_query_grid")
            global LAUNCHES
            LAUNCHES += 3   # count, scan, fill (+ the query counted by _timed)
        else:
            _lib.check(_lib.lib().b2r_ball_query(new_xyz.data_ptr(), xyz.data_ptr(), B, N, M,
                                                 float(radius), int(nsample), out.data_ptr(),
                                                 _stream()), "ball_query")
    return out


def group_points(points, idx):
    """(B,C,N), (B,NP,NS) -> (B,C,NP,NS).  Replaces group_points.cpp:17-40."""
    _chk_contig(points, "points")
    _chk_contig(idx, "idx")
    _chk_float(points, "points")
    _chk_int(idx, "idx")
    _chk_cuda(points, [(idx, "idx")])
    B, C, N = points.shape
    NP, NS = idx.size(1), idx.size(2)
    out = torch.empty((B, C, NP, NS), dtype=torch.float32, device=points.device)
    with _on_device(points), _timed("group_points"):
        _lib.check(_lib.lib().b2r_group_fwd(points.data_ptr(), idx.data_ptr(), B, C, N, NP, NS,
                                            out.data_ptr(), _stream()), "group_points")
    return out


# B2R_SCATTER_ATOMIC=1 keeps round 1's RED.ADD.F32 scatter kernels (b2r_group_bwd /
# b2r_three_interp_bwd: unordered sums, atomics-bound) for A/B measurements
SCATTER_ATOMIC = os.environ.get("B2R_SCATTER_ATOMIC", "0") not in ("0", "")
# B2R_DETERMINISTIC=1: always take the plan path (run-to-run bit-identical gradients of
# group_points / three_interpolate), also where the atomic scatter would be faster
DETERMINISTIC = os.environ.get("B2R_DETERMINISTIC", "0") not in ("0", "")
_PLAN_CACHE = {}   # (ptr, version, shape, N, weight ptr/version) -> plan tensor; a few entries


def scatter_plan(idx, N, weight=None):
    """Inverse index of `idx` for the deterministic backward kernels (include/b2r.h:
    b2r_scatter_plan).  idx (B,NP,NS) for grouping, (B,n,3) + weight for three_interpolate.  The
    same index tensor is scattered through twice per QueryAndGroup (xyz and features), so plans
    are cached on (storage, version) outside CUDA-graph capture."""
    B = idx.size(0)
    E = idx.numel() // max(B, 1)
    K = 3 if weight is not None else 1
    capturing = torch.cuda.is_current_stream_capturing()
    key = (idx.data_ptr(), idx._version, tuple(idx.shape), int(N),
           None if weight is None else (weight.data_ptr(), weight._version))
    if not capturing and key in _PLAN_CACHE:
        return _PLAN_CACHE[key]
    l = _lib.lib()
    nbytes = int(l.b2r_scatter_plan_bytes(B, E, int(N), K, int(weight is not None)))
    plan = torch.empty((nbytes,), dtype=torch.uint8, device=idx.device)
    _lib.check(l.b2r_scatter_plan(idx.data_ptr(), None if weight is None else weight.data_ptr(), B, E,
                                  int(N), K, plan.data_ptr(), nbytes, _stream()), "scatter_plan")
    if not capturing:
        if len(_PLAN_CACHE) >= 8:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        # the key holds raw pointers: keep the tensors alive with the plan so they cannot be
        # recycled under the same address + version
        _PLAN_CACHE[key] = plan
        plan._b2r_keepalive = (idx, weight)
    return plan


def _plan_pays(entries, targets, per_source, rows):
    """The plan kernels walk every (source tile, target) pair of a row and run one CTA per (b,c) row:
    they beat the atomic scatter when lists are not mostly empty (entries >= 2 x tiles x targets) and
    there are rows enough to fill the GPU; otherwise the atomic kernels stay (measured:
    profiles/r02/movers_roofline.log)."""
    if DETERMINISTIC:
        return True
    if (per_source == 3 and targets <= 4096) or targets <= 1024:
        # multi-channel gather: 4 rows per CTA share one walk over the lists
        return rows >= 512 and entries >= targets
    tiles = -(-(entries // per_source) // 16384)
    return entries >= 2 * tiles * targets and rows >= 96


def group_points_grad(grad_out, idx, n):
    """(B,C,NP,NS), (B,NP,NS), n -> (B,C,n).  Replaces group_points.cpp:42-65."""
    _chk_contig(grad_out, "grad_out")
    _chk_contig(idx, "idx")
    _chk_float(grad_out, "grad_out")
    _chk_int(idx, "idx")
    _chk_cuda(grad_out, [(idx, "idx")])
    B, C, NP, NS = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out), _timed("group_points_grad"):
        if SCATTER_ATOMIC or B * C * NP * NS == 0 or int(n) == 0 or not _plan_pays(NP * NS, int(n), 1, B * C):
            _lib.check(_lib.lib().b2r_group_bwd(grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), NP,
                                                NS, out.data_ptr(), _stream()), "group_points_grad")
        else:
            plan = scatter_plan(idx, int(n))
            _lib.check(_lib.lib().b2r_group_bwd_plan(grad_out.data_ptr(), plan.data_ptr(), B, C, int(n),
                                                     NP, NS, out.data_ptr(), _stream()),
                       "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """(B,n,3), (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32].  Replaces interpolate.cpp:19-45."""
    _chk_contig(unknowns, "unknowns")
    _chk_contig(knows, "knows")
    _chk_float(unknowns, "unknowns")
    _chk_float(knows, "knows")
    _chk_cuda(unknowns, [(knows, "knows")])
    B, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with _on_device(unknowns), _timed("three_nn"):
        _lib.check(_lib.lib().b2r_three_nn(unknowns.data_ptr(), knows.data_ptr(), B, n, m,
                                           dist2.data_ptr(), idx.data_ptr(), _stream()),
                   "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,C,m), (B,n,3), (B,n,3) -> (B,C,n).  Replaces interpolate.cpp:47-75."""
    _chk_contig(points, "points")
    _chk_contig(idx, "idx")
    _chk_contig(weight, "weight")
    _chk_float(points, "points")
    _chk_int(idx, "idx")
    _chk_float(weight, "weight")
    _chk_cuda(points, [(idx, "idx"), (weight, "weight")])
    B, C, m = points.shape
    n = idx.size(1)
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with _on_device(points), _timed("three_interpolate"):
        _lib.check(_lib.lib().b2r_three_interp_fwd(points.data_ptr(), idx.data_ptr(),
                                                   weight.data_ptr(), B, C, m, n, out.data_ptr(),
                                                   _stream()), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,C,n), (B,n,3), (B,n,3), m -> (B,C,m).  Replaces interpolate.cpp:76-104."""
    _chk_contig(grad_out, "grad_out")
    _chk_contig(idx, "idx")
    _chk_contig(weight, "weight")
    _chk_float(grad_out, "grad_out")
    _chk_int(idx, "idx")
    _chk_float(weight, "weight")
    _chk_cuda(grad_out, [(idx, "idx"), (weight, "weight")])
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out), _timed("three_interpolate_grad"):
        if SCATTER_ATOMIC or B * C * n == 0 or int(m) == 0 or not _plan_pays(3 * n, int(m), 3, B * C):
            _lib.check(_lib.lib().b2r_three_interp_bwd(grad_out.data_ptr(), idx.data_ptr(),
                                                       weight.data_ptr(), B, C, n, int(m),
                                                       out.data_ptr(), _stream()),
                       "three_interpolate_grad")
        else:
            plan = scatter_plan(idx, int(m), weight)
            _lib.check(_lib.lib().b2r_three_interp_bwd_plan(grad_out.data_ptr(), plan.data_ptr(), B, C,
                                                            n, int(m), out.data_ptr(), _stream()),
                       "three_interpolate_grad")
    return out


# ---- beyond the reference's nine: the fused QueryAndGroup tail (include/b2r.h) -------------
def query_group(xyz, new_xyz, features, idx, radius, normalize_xyz):
    """One-pass group(xyz)-new_xyz(/radius) ++ group(features) -> (B,3+C,NP,NS)."""
    _chk_cuda(xyz, [(new_xyz, "new_xyz"), (idx, "idx")])
    for t, nme in ((xyz, "xyz"), (new_xyz, "new_xyz")):
        _chk_contig(t, nme)
        _chk_float(t, nme)
    _chk_contig(idx, "idx")
    _chk_int(idx, "idx")
    C = 0
    fptr = None
    if features is not None:
        _chk_contig(features, "features")
        _chk_float(features, "features")
        _chk_cuda(xyz, [(features, "features")])
        C = features.size(1)
        fptr = features.data_ptr()
    B, N = xyz.size(0), xyz.size(1)
    NP, NS = idx.size(1), idx.size(2)
    out = torch.empty((B, 3 + C, NP, NS), dtype=torch.float32, device=xyz.device)
    # algorithmic bytes: output write + compulsory reads of features / xyz / centres / idx
    L = NP * NS
    algo = B * (4 * (3 + C) * L + min(4 * C * N, 4 * C * L) + min(12 * N, 12 * L) + 12 * NP + 4 * L)
    with _on_device(xyz), _timed("query_group", algo):
        _lib.check(_lib.lib().b2r_query_group_fwd(xyz.data_ptr(), new_xyz.data_ptr(), fptr,
                                                  idx.data_ptr(), B, C, N, NP, NS, float(radius),
                                                  1 if normalize_xyz else 0, out.data_ptr(),
                                                  _stream()), "query_group")
    return out


def query_group_grad(grad_out, idx, N, C, radius, normalize_xyz, need_xyz, need_new_xyz,
                     need_features):
    """Backward of query_group: returns (grad_xyz|None, grad_new_xyz|None, grad_features|None)."""
    _chk_contig(grad_out, "grad_out")
    _chk_float(grad_out, "grad_out")
    _chk_cuda(grad_out, [(idx, "idx")])
    B, _, NP, NS = grad_out.shape
    dev = grad_out.device
    gx = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if need_xyz else None
    gn = torch.empty((B, NP, 3), dtype=torch.float32, device=dev) if need_new_xyz else None
    gf = torch.empty((B, C, N), dtype=torch.float32, device=dev) if (need_features and C) else None
    with _on_device(grad_out), _timed("query_group_grad"):
        _lib.check(_lib.lib().b2r_query_group_bwd(
            grad_out.data_ptr(), idx.data_ptr(), B, C, int(N), NP, NS, float(radius),
            1 if normalize_xyz else 0,
            gx.data_ptr() if gx is not None else None,
            gn.data_ptr() if gn is not None else None,
            gf.data_ptr() if gf is not None else None, _stream()), "query_group_grad")
    return gx, gn, gf
