"""ctypes binding of ``libfbbev_b200.so`` (C ABI: ``include/fbbev_b200.h``).

PyTorch is plumbing here: it owns device memory and streams; every kernel on
the hot path is in the shared library.  There is NO fallback -- if the library
is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FBBEV_LIB: kernel-tuning aid (tools/), points at a variant build of the same ABI
LIB_PATH = os.environ.get("FBBEV_LIB") or os.path.join(
    _HERE, "lib", "libfbbev_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_p = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/fbbev_b200.h declaration order
_SIGNATURES = {
    "fbbev_abi_version": (ctypes.c_int, []),
    "fbbev_debug_launch_count": (ctypes.c_longlong, []),
    "fbbev_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "fbbev_bev_pool_v2_fwd": (ctypes.c_int, [_p] * 7 + [_i32, _i32, _p, _p]),
    "fbbev_bev_pool_v2_dense_workspace_bytes": (
        _sz, [_i32, _i64, _i32, _i32, _i32]),
    "fbbev_bev_pool_v2_fwd_dense": (
        ctypes.c_int,
        [_p] * 7 + [_i32, _p, _i32, _i32, _i32, _i64, _p, _p, _sz, _p]),
    "fbbev_bev_pool_v2_plan": (
        ctypes.c_int,
        [_p, _p, _p, _i32, _p, _i32, _i32, _i32, _i64, _p, _sz, _p]),
    "fbbev_bev_pool_v2_fwd_dense_planned": (
        ctypes.c_int,
        [_p] * 7 + [_i32, _i32, _i32, _i32, _i64, _p, _p, _sz, _p]),
    "fbbev_bev_pool_v2_sums_planned": (
        ctypes.c_int, [_p] * 7 + [_i32, _i32, _i32, _i32, _i64, _p, _sz, _p]),
    "fbbev_bev_pool_v2_zmean_planned": (
        ctypes.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i64, _i32, _p, _p, _sz,
                       _p]),
    "fbbev_bev_pool_v2_write_planned": (
        ctypes.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i64, _i32, _p, _p, _p,
                       _sz, _p]),
    "fbbev_bev_pool_v2_bwd": (ctypes.c_int, [_p] * 8 + [_i32, _i32, _p, _p, _p]),
    "fbbev_bev_pool_v2_bwd_bczyx": (
        ctypes.c_int, [_p] * 8 + [_i32, _i32, _i64, _p, _p, _p]),
    "fbbev_voxel_prepare_workspace_bytes": (_sz, [_i64, _i64]),
    "fbbev_voxel_prepare": (
        ctypes.c_int, [_p] + [_i32] * 5 + [_p] * 3 + [_p] * 6 +
        [_p, _sz, _i32, _p, _sz, _p]),
    "fbbev_voxel_prepare_cams": (
        ctypes.c_int, [_p] * 8 + [_i32] * 6 + [_p] * 3 + [_p] * 6 +
        [_p, _sz, _i32, _p, _sz, _p]),
    "fbbev_voxel_prepare_can_plan": (ctypes.c_int, [_i32, _i64]),
    "fbbev_voxel_prepare_sparse": (
        ctypes.c_int, [_p, _p, ctypes.c_float] + [_i32] * 5 + [_p] * 3 +
        [_p] * 6 + [_p, _sz, _i32, _p, _sz, _p]),
    "fbbev_voxel_prepare_cams_sparse": (
        ctypes.c_int, [_p] * 8 + [_i32, _p, ctypes.c_float] + [_i32] * 5 +
        [_p] * 3 + [_p] * 6 + [_p, _sz, _i32, _p, _sz, _p]),
    "fbbev_lift_tail_fwd": (
        ctypes.c_int, [_p, _i64, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "fbbev_bev_mask_fold_workspace_bytes": (_sz, [_i32, _i32]),
    "fbbev_bev_mask_fold": (
        ctypes.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "fbbev_point_sampling": (
        ctypes.c_int, [_p] * 3 + [_i32] * 3 + [_p] * 5 + [_i32] * 3 +
        [ctypes.c_float] * 4 + [_p] * 4),
    "fbbev_bev_query_init": (ctypes.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "fbbev_tokens_to_map": (ctypes.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "fbbev_msda_fwd": (ctypes.c_int, [_p] * 5 + [_i32] * 7 + [_p, _p]),
    "fbbev_msda_bwd": (ctypes.c_int, [_p] * 6 + [_i32] * 7 + [_p] * 4),
    "fbbev_msda_fused_fwd": (ctypes.c_int, [_p] * 6 + [_i32] * 8 + [_p, _p]),
    "fbbev_da_sca_workspace_bytes": (_sz, [_i32, _i32]),
    "fbbev_da_sca_fwd": (ctypes.c_int,
                         [_p] * 10 + [_i32] * 10 + [_p, _p, _sz, _i32, _p]),
    "fbbev_da_sca_prologue": (ctypes.c_int,
                              [_p] + [_i32] * 9 + [_p, _p, _sz, _p]),
    "fbbev_history_warp": (ctypes.c_int, [_p, _i64, _p] + [_i32] * 5 +
                           [_p, _i32, _i32, _p]),
    "fbbev_linear_packed_bytes": (ctypes.c_size_t, [_i32, _i32]),
    "fbbev_linear_pack": (ctypes.c_int, [_p, _i32, _i32, _p, _p]),
    "fbbev_linear_fwd_split": (ctypes.c_int, [
        _p, _i64, _p, _i64, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _i64, _p,
        _i64, _p]),
    "fbbev_linear_fwd": (ctypes.c_int, [
        _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _i32,
        ctypes.c_float, _p, _i64, _p]),
    "fbbev_ffn_supported": (ctypes.c_int, [_i32, _i32]),
    "fbbev_ffn_fwd": (ctypes.c_int, [
        _p, _i64, _p, _p, _p, _p, _p, _i64, _p, _p, _i64, _i32, _i32,
        ctypes.c_float, _p, _i64, _p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

ABI_VERSION = 4
_lib = None


class FbbevError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libfbbev_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise FbbevError("building libfbbev_b200.so failed")
    return LIB_PATH


def lib():
    """Load the library (never builds implicitly; never falls back)."""
    global _lib
    if _proxy is not None and _lib is not None:
        return _proxy
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FbbevError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ "
                f"as g; g.build()'` (or `make -C {CSRC_DIR}`).  The CUDA "
                "library is the product; there is no CPU or eager fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
        if _lib.fbbev_abi_version() != ABI_VERSION:
            raise FbbevError("libfbbev_b200.so ABI version mismatch")
    return _lib


class KernelTimer:
    """Diagnostics: bracket every C-ABI launch with CUDA events on the calling
    stream (``with KernelTimer() as t: ...; t.records``).  bench.py uses it to
    report per-kernel times / roofline fractions measured live in the run; it
    is off (zero overhead: ``lib()`` returns the CDLL itself) otherwise."""

    _SKIP = ("fbbev_abi_version", "fbbev_debug_launch_count",
             "fbbev_error_string", "fbbev_linear_packed_bytes",
             "fbbev_bev_pool_v2_dense_workspace_bytes",
             "fbbev_voxel_prepare_workspace_bytes",
             "fbbev_da_sca_workspace_bytes")

    def __init__(self):
        self.records = []   # (name, args, start_event, end_event)
        self._real = None

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if name in self._SKIP or not name.startswith("fbbev_"):
            return fn

        def timed(*args):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, args, e0, e1))
            return rc
        return timed

    def __enter__(self):
        global _proxy
        self._real = lib()
        _proxy = self
        return self

    def __exit__(self, *exc):
        global _proxy
        _proxy = None
        return False


_proxy = None


def check(code, what):
    if code != 0:
        msg = lib().fbbev_error_string(code).decode()
        raise FbbevError(f"{what} failed: {msg} (code {code})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise FbbevError(
                "fb-bev_b200 ops run on CUDA tensors only (no CPU fallback); "
                f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise FbbevError("all tensors must be on the same CUDA device")
    return dev


def matmul_order_flags(*mats):
    """FBBEV_ORDER_SEQ_* bits for the 3x3 matrix stacks a geometry kernel
    applies per point, one bit per argument.

    torch's broadcast ``M.view(.., 1, 1, 1, 3, 3).matmul(pts)`` rounds each
    output as ``fma(m1, x1, m0*x0) + m2*x2`` on B200 -- except when M is ONE
    matrix in the column-major layout ``torch.inverse`` returns: the expand is
    then a stride-0 view, cuBLAS receives op 't' and its kernel for that case
    uses the sequential FMA chain (tools/micro/matmul_order2.py).  The flags
    are derived from the tensors exactly as the eager chain would see them, so
    they must be taken BEFORE any ``.contiguous()``."""
    flags = 0
    for i, m in enumerate(mats):
        if m.numel() == 9 and m.stride(-1) != 1:
            flags |= 1 << i
    return flags


def c_floats(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])
