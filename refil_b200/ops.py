"""Tensor-level wrappers over the C ABI (include/refil_b200.h).  PyTorch is used for device memory and streams only;
every function here enqueues hand-written sm_100a kernels on the current CUDA stream.  No CPU fallback: a
non-CUDA tensor raises."""
import torch

from . import _lib

ATTN_PART_WITHIN, ATTN_PART_INTERACT, ATTN_ACTIVE0, ATTN_DEFAULT = 1, 2, 4, 8
MIX_KIND = {"flex_qmix": 0, "lin_flex_qmix": 1, "vdn": 2}

_launches = 0  # number of kernels-launching C-ABI calls issued (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


def add_launches(n):
    """Account kernel launches that did not go through _call (CUDA-graph replays of a captured step)."""
    global _launches
    _launches += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t, dtype=None):
    """device pointer of a contiguous CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.RefilError("refil_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise _lib.RefilError("tensor of shape %s is not contiguous" % (tuple(t.shape),))
    if dtype is not None and t.dtype != dtype:
        raise _lib.RefilError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


_timing = None  # when enabled: {kernel name: [cuda event pairs]}


def set_timing(enabled):
    """Per-entry-point CUDA-event timing (bench.py's per-kernel breakdown; never on during the timed region)."""
    global _timing
    _timing = {} if enabled else None
    _work.clear()
    _shape_timing.clear()


_shape_timing = {}    # when timing: {"name[MxNxK]": [cuda event pairs]} for the dense-layer kernels


def shape_timing_summary():
    """-> {"name[MxNxK]": (launches, total_ms)}; synchronises."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in _shape_timing.items()}


_work = {}    # when timing: {kernel name: [algorithmic flops, algorithmic bytes]} accumulated per call


def _account(name, flops, nbytes):
    if _timing is not None:
        w = _work.setdefault(name, [0.0, 0.0])
        w[0] += flops
        w[1] += nbytes


def timing_summary():
    """-> {name: (launches, total_ms, algorithmic flops, algorithmic bytes)}; synchronises."""
    torch.cuda.synchronize()
    out = {}
    for name, evs in (_timing or {}).items():
        w = _work.get(name, [0.0, 0.0])
        out[name] = (len(evs), sum(a.elapsed_time(b) for a, b in evs), w[0], w[1])
    return out


def _call(name, *args, shape=None, as_name=None):
    """as_name: the timing bucket (a grouped launch is booked under its single-problem entry point, where its work is accounted)."""
    global _launches
    _launches += 1
    if _timing is None:
        _lib.call(name, *args, _stream())
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.call(name, *args, _stream())
    b.record()
    _timing.setdefault(as_name or name, []).append((a, b))
    if shape is not None:
        _shape_timing.setdefault("%s[%s]" % (name, "x".join(str(int(x)) for x in shape)), []).append((a, b))


def _pv(t, dtype=None):
    """device pointer of a 2-D CUDA tensor with contiguous rows -- a matrix or a column slice of a wider one (row stride = stride(0))"""
    if t.is_contiguous():
        return _p(t, dtype)
    if not t.is_cuda:
        raise _lib.RefilError("refil_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU fallback" % t.device)
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.RefilError("tensor of shape %s / strides %s is not a row-strided matrix" % (tuple(t.shape), tuple(t.stride())))
    if dtype is not None and t.dtype != dtype:
        raise _lib.RefilError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


F32, U8, I32, I64, F64 = torch.float32, torch.uint8, torch.int32, torch.int64, torch.float64


def _rm(row_mask):
    """row mask spec (entity_mask [N, ne] u8, na, rows_per_copy) or None"""
    if row_mask is None:
        return None, 1, 1, 1
    em, na, mper = row_mask
    return _p(em, U8), na, em.shape[-1], mper


# ---------------------------------------------------------------------------------------------------- dense
USE_TENSOR_CORES = True     # tcgen05 3xTF32 path for eligible shapes; False = fp32 FFMA kernel everywhere
TC_MIN_ROWS = 256


def _tc_ok(M, N, K):
    return USE_TENSOR_CORES and M >= TC_MIN_ROWS and _lib.load().refil_tc_gemm_supported(M, N, K) != 0


class min_tiles:
    """with ops.min_tiles(n): dense launches inside give every CTA at least n m-tiles (see refil_tc_set_min_tiles)."""

    def __init__(self, n):
        self.n = int(n or 0)

    def __enter__(self):
        self.prev = _lib.load().refil_tc_set_min_tiles(self.n) if self.n else None

    def __exit__(self, *exc):
        if self.n:
            _lib.load().refil_tc_set_min_tiles(self.prev)
        return False


def tc_head_ok(rows, n_out, width):
    """Can the backward of an [n_out <= 32, width] output head run on the tensor cores with its gradient padded to 32 columns?"""
    return _tc_ok(rows, width, 32) and _lib.load().refil_tc_wgrad_supported(rows, n_out, width) != 0


def tc_gemm_tn(A, B, sbn, sbk, out, N, K, bias=None, relu=False, relu_y=None, a_row_mask=None, c_row_mask=None, k_valid=0):
    M = A.shape[0]
    aem, ana, ane, amper = _rm(a_row_mask)
    cem, cna, cne, cmper = _rm(c_row_mask)
    _account("tc_gemm_tn", 2.0 * M * N * K, 4.0 * (M * K * (2 if relu_y is not None else 1) + M * N + N * K))
    _call("tc_gemm_tn", _pv(A, F32), A.stride(0), _p(relu_y, F32), A.shape[1], aem, ana, ane, amper, _pv(B, F32), sbn, sbk,
          int(k_valid), _p(bias, F32), int(relu), cem, cna, cne, cmper, _pv(out, F32), out.stride(0), M, N, K, shape=(M, N, K))
    return out


QKV_SPLIT_MIN_ROWS = 65536   # entity rows from which in_trans is split into K|V (all rows) and Q (agent rows only)


def qkv_split_ok(n_units, ne, na, d):
    """Split in_trans into a K|V product over all entity rows and a Q product over the agent rows alone?  The reference computes
    queries for every entity and keeps the agents' (modules/layers/attention.py:46-48); the other 2/3 of the Q third are never read.
    REFIL_QKV_SPLIT=0 / 1 forces the choice (A/B runs)."""
    import os
    env = os.environ.get("REFIL_QKV_SPLIT")
    if env == "0" or not USE_TENSOR_CORES or na >= ne or 32 % na != 0 or os.environ.get("REFIL_TC_MODE") == "ss":
        return False
    if not (_tc_ok(n_units * ne, 2 * d, d) and _tc_ok(n_units * na, d, d) and _lib.load().refil_tc_gemm_k_slices(d, 2 * d) >= 1):
        return False
    return env == "1" or n_units * ne >= QKV_SPLIT_MIN_ROWS


def tc_gemm_tn_rows(A, W, out, n_groups, group_rows, group_stride, accumulate=False, transpose_w=False):
    """Row-grouped product on the first `group_rows` of every `group_stride` rows of A and out (refil_tc_gemm_tn_rows):
    out (+)= A W^T, or A W with transpose_w (backward-data).  A / out may be column slices of wider matrices."""
    if transpose_w:
        K, N = W.shape
        sbn, sbk = 1, W.stride(0)
    else:
        N, K = W.shape
        sbn, sbk = W.stride(0), 1
    M = n_groups * group_rows
    _account("tc_gemm_tn", 2.0 * M * N * K, 4.0 * (M * K + M * N * (2 if accumulate else 1) + N * K))
    _call("tc_gemm_tn_rows", _pv(A, F32), A.stride(0), _pv(W, F32), sbn, sbk, _pv(out, F32), out.stride(0), n_groups, group_rows,
          group_stride, int(accumulate), N, K, shape=(M, N, K), as_name="tc_gemm_tn")
    return out


def in_trans_fwd_split(x1, W, qkv, n_units, ne, na):
    """qkv[:, d:] = x1 W[d:]^T for every entity row, qkv[agent rows, :d] = x1[agent rows] W[:d]^T, as ONE launch of two problems;
    the Q columns of the other rows are left unwritten (nothing reads them: the attention kernels take queries from the agent
    rows only)."""
    d = x1.shape[1]
    M = n_units * ne
    arr = (_GemmDesc * 2)()
    kv, q = arr[0], arr[1]
    for dsc in (kv, q):
        dsc.A, dsc.lda, dsc.relu_y, dsc.ldy = _p(x1, F32), d, None, d
        dsc.a_mask, dsc.a_na, dsc.a_ne, dsc.a_mper = None, 1, 1, 1
        dsc.sbn, dsc.sbk, dsc.b_k_valid, dsc.bias, dsc.relu = W.stride(0), 1, 0, None, 0
        dsc.c_mask, dsc.c_na, dsc.c_ne, dsc.c_mper = None, 1, 1, 1
        dsc.ldc = qkv.stride(0)
    kv.B, kv.C, kv.M, kv.n_cols = _pv(W[d:], F32), _pv(qkv[:, d:], F32), M, 2 * d
    q.B, q.C, q.M, q.n_cols = _pv(W[:d], F32), _p(qkv, F32), n_units * na, d
    q.row_group, q.row_group_stride = na, ne
    _account("tc_gemm_tn", 2.0 * (M * 2 * d + n_units * na * d) * d, 4.0 * (M * d + M * 2 * d + n_units * na * d + 3 * d * d))
    _call("tc_gemm_tn_group", _ct.addressof(arr), 2, 2 * d, d, shape=(2, M, 2 * d, d), as_name="tc_gemm_tn")
    return qkv
