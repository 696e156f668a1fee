"""ctypes binding of the C parity oracle (``oracle/fbbev_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  All functions take and
return numpy arrays on the host.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfbbev_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C oracle (and oracle/_ref when the reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) <
            os.path.getmtime(os.path.join(_HERE, "fbbev_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "libfbbev_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_voxel_prepare.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def voxel_prepare(coor, lo, iv, gs, rank_mode=0):
    """view_transformer.py:547-605.  coor (B,N,D,H,W,3) fp32.

    Returns (ranks_bev, ranks_depth, ranks_feat, interval_starts,
    interval_lengths) as int32 arrays, or five ``None`` when nothing is kept
    (the reference's empty case, :598-599).
    """
    coor = _f32(coor)
    B, N, D, H, W, three = coor.shape
    assert three == 3
    n = B * N * D * H * W
    lo, iv, gs = _f32(lo), _f32(iv), _f32(gs)
    rb = np.empty(n, np.int32)
    rd = np.empty(n, np.int32)
    rf = np.empty(n, np.int32)
    st = np.empty(n, np.int32)
    ln = np.empty(n, np.int32)
    nk = ctypes.c_int64(0)
    ni = ctypes.c_int64(0)
    rc = lib().oracle_voxel_prepare(
        _p(coor, _f32p), B, N, D, H, W, _p(lo, _f32p), _p(iv, _f32p),
        _p(gs, _f32p), int(rank_mode), _p(rb, _i32p), _p(rd, _i32p),
        _p(rf, _i32p), _p(st, _i32p), _p(ln, _i32p), ctypes.byref(nk),
        ctypes.byref(ni))
    if rc != 0:
        raise MemoryError("oracle_voxel_prepare failed")
    nk, ni = nk.value, ni.value
    if ni == 0:
        return None, None, None, None, None
    return (rb[:nk].copy(), rd[:nk].copy(), rf[:nk].copy(), st[:ni].copy(),
            ln[:ni].copy())


def bev_pool_v2_fwd(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                    bev_feat_shape, interval_starts, interval_lengths):
    """bev_pool_cuda.cu:18-45 on a zero-filled (B,Z,Y,X,C) volume
    (QuickCumsumCuda.forward, bev_pool.py:15-39)."""
    depth, feat = _f32(depth), _f32(feat)
    c = feat.shape[-1]
    out = np.zeros(bev_feat_shape, np.float32)
    assert out.shape[-1] == c
    rd, rf, rb = _i32(ranks_depth), _i32(ranks_feat), _i32(ranks_bev)
    st, ln = _i32(interval_starts), _i32(interval_lengths)
    lib().oracle_bev_pool_v2_fwd(
        c, len(st), _p(depth, _f32p), _p(feat, _f32p), _p(rd, _i32p),
        _p(rf, _i32p), _p(rb, _i32p), _p(st, _i32p), _p(ln, _i32p),
        _p(out, _f32p))
    return out


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths,
                scratch=None, out=None):
    """The op as shipped (bev_pool.py:84-90): zeros + kernel + permute to
    (B,C,Z,Y,X).  ``scratch``/``out`` let bench.py reuse buffers."""
    depth, feat = _f32(depth), _f32(feat)
    c = feat.shape[-1]
    B, Z, Y, X, C = bev_feat_shape
    assert C == c
    zyx = Z * Y * X
    if scratch is None:
        scratch = np.empty(B * zyx * c, np.float32)
    if out is None:
        out = np.empty((B, c, Z, Y, X), np.float32)
    rd, rf, rb = _i32(ranks_depth), _i32(ranks_feat), _i32(ranks_bev)
    st, ln = _i32(interval_starts), _i32(interval_lengths)
    lib().oracle_bev_pool_v2_op(
        c, len(st), _p(depth, _f32p), _p(feat, _f32p), _p(rd, _i32p),
        _p(rf, _i32p), _p(rb, _i32p), _p(st, _i32p), _p(ln, _i32p), B,
        ctypes.c_int64(zyx), _p(scratch, _f32p), _p(out, _f32p))
    return out


def bev_pool_v2_bwd(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """QuickCumsumCuda.backward (bev_pool.py:42-81) + bev_pool_grad_kernel
    (bev_pool_cuda.cu:64-118).  out_grad is (B,Z,Y,X,C)."""
    out_grad, depth, feat = _f32(out_grad), _f32(depth), _f32(feat)
    c = feat.shape[-1]
    rd, rf, rb = _i32(ranks_depth), _i32(ranks_feat), _i32(ranks_bev)
    order = np.argsort(rf, kind="stable")  # bev_pool.py:45
    rf, rd, rb = rf[order], rd[order], rb[order]
    kept = np.ones(len(rb), bool)  # :48-50
    kept[1:] = rf[1:] != rf[:-1]
    st = np.nonzero(kept)[0].astype(np.int32)  # :51
    ln = np.zeros_like(st)  # :52-55
    ln[:-1] = st[1:] - st[:-1]
    ln[-1] = len(rb) - st[-1]
    dg = np.zeros_like(depth)
    fg = np.zeros_like(feat)
    lib().oracle_bev_pool_v2_bwd(
        c, len(st), _p(out_grad, _f32p), _p(depth, _f32p), _p(feat, _f32p),
        _p(rd, _i32p), _p(rf, _i32p), _p(rb, _i32p), _p(st, _i32p),
        _p(ln, _i32p), _p(dg, _f32p), _p(fg, _f32p))
    return dg, fg


def msda_fwd(value, spatial_shapes, level_start_index, sampling_locations,
             attention_weights):
    """ext_module.ms_deform_attn_forward
    (multi_scale_deformable_attn_function.py:127-133)."""
    value = _f32(value)
    loc, attw = _f32(sampling_locations), _f32(attention_weights)
    ss = np.ascontiguousarray(spatial_shapes, np.int64)
    ls = np.ascontiguousarray(level_start_index, np.int64)
    bs, n_value, heads, ch = value.shape
    _, nq, _, levels, points, _ = loc.shape
    out = np.empty((bs, nq, heads * ch), np.float32)
    lib().oracle_msda_fwd(_p(value, _f32p), _p(ss, _i64p), _p(ls, _i64p),
                          _p(loc, _f32p), _p(attw, _f32p), bs, n_value, heads,
                          ch, levels, nq, points, _p(out, _f32p))
    return out


def msda_bwd(value, spatial_shapes, level_start_index, sampling_locations,
             attention_weights, grad_output):
    """ext_module.ms_deform_attn_backward
    (multi_scale_deformable_attn_function.py:159-169)."""
    value = _f32(value)
    loc, attw = _f32(sampling_locations), _f32(attention_weights)
    go = _f32(grad_output)
    ss = np.ascontiguousarray(spatial_shapes, np.int64)
    ls = np.ascontiguousarray(level_start_index, np.int64)
    bs, n_value, heads, ch = value.shape
    _, nq, _, levels, points, _ = loc.shape
    gv, gl, ga = np.zeros_like(value), np.zeros_like(loc), np.zeros_like(attw)
    lib().oracle_msda_bwd(_p(value, _f32p), _p(ss, _i64p), _p(ls, _i64p),
                          _p(loc, _f32p), _p(attw, _f32p), _p(go, _f32p), bs,
                          n_value, heads, ch, levels, nq, points,
                          _p(gv, _f32p), _p(gl, _f32p), _p(ga, _f32p))
    return gv, gl, ga
