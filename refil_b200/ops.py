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


def in_trans_bwd_data_split(dqkv, W, dx1, n_units, ne, na):
    """dx1 = dqkv[:, d:] W[d:] on every entity row + dqkv[agent rows, :d] W[:d] accumulated into the agent rows, as ONE launch of two
    problems (the Q columns of dqkv are zero outside the agent rows: the backward of the slice `query[:n_queries]`)."""
    d = dx1.shape[1]
    M = n_units * ne
    arr = (_GemmDesc * 2)()
    kv, q = arr[0], arr[1]
    for dsc in (kv, q):
        dsc.lda, dsc.relu_y, dsc.ldy = dqkv.stride(0), None, dqkv.stride(0)
        dsc.a_mask, dsc.a_na, dsc.a_ne, dsc.a_mper = None, 1, 1, 1
        dsc.sbn, dsc.sbk, dsc.b_k_valid, dsc.bias, dsc.relu = 1, W.stride(0), 0, None, 0     # "B"[n = k_in, i] = W[i, n]
        dsc.c_mask, dsc.c_na, dsc.c_ne, dsc.c_mper = None, 1, 1, 1
        dsc.C, dsc.ldc = _p(dx1, F32), d
    kv.A, kv.B, kv.M, kv.k_len = _pv(dqkv[:, d:], F32), _pv(W[d:], F32), M, 2 * d
    q.A, q.B, q.M, q.k_len = _p(dqkv, F32), _pv(W[:d], F32), n_units * na, d
    q.row_group, q.row_group_stride, q.accumulate = na, ne, 1
    _account("tc_gemm_tn", 2.0 * (M * 2 * d + n_units * na * d) * d, 4.0 * (M * 2 * d + n_units * na * d + M * d + 3 * d * d))
    _call("tc_gemm_tn_group", _ct.addressof(arr), 2, d, 2 * d, shape=(2, M, d, 2 * d), as_name="tc_gemm_tn")
    return dx1


def tc_wgrad_y_tma():
    """Is the TMA path of the weight-gradient kernel's Y operand on (row-grouped problems need it)?  REFIL_TCW_Y=ldg turns it off."""
    import os
    return os.environ.get("REFIL_TCW_Y", "")[:1] != "l" and os.environ.get("REFIL_TC_MODE") != "ss"


def in_trans_bwd_weight_split(dqkv, x1, dW, n_units, ne, na):
    """dW[d:] += dqkv[:, d:]^T x1 over every entity row, dW[:d] += dqkv[agent rows, :d]^T x1[agent rows], as ONE launch of two
    problems (the Q columns of dqkv are zero outside the agent rows)."""
    d = x1.shape[1]
    M = n_units * ne
    arr = (_WgradDesc * 2)()
    kv, q = arr[0], arr[1]
    for dsc in (kv, q):
        dsc.ldx, dsc.relu_y, dsc.ldy = dqkv.stride(0), None, dqkv.stride(0)
        dsc.x_mask, dsc.na, dsc.ne, dsc.mper = None, 1, 1, 1
        dsc.Y, dsc.ldyy, dsc.y_shift, dsc.y_period = _p(x1, F32), d, 0, 1
        dsc.lddw, dsc.q_valid, dsc.db = d, 0, None
    kv.X, kv.dW, kv.M, kv.p_cols = _pv(dqkv[:, d:], F32), _pv(dW[d:], F32), M, 2 * d
    q.X, q.dW, q.M, q.p_cols = _p(dqkv, F32), _pv(dW[:d], F32), n_units * na, d
    q.row_group, q.row_group_stride = na, ne
    _account("tc_gemm_wgrad", 2.0 * (M * 2 * d + n_units * na * d) * d, 4.0 * (M * 2 * d + M * d + 2 * n_units * na * d + 3 * d * d))
    _call("tc_gemm_wgrad_group", _ct.addressof(arr), 2, 2 * d, d, shape=(2, M, 2 * d, d), as_name="tc_gemm_wgrad")


# ---- grouped launches: the same layer of several networks in ONE kernel launch (include/refil_b200.h: RefilGemmDesc) ----------
import ctypes as _ct


class _GemmDesc(_ct.Structure):
    _fields_ = [("A", _ct.c_void_p), ("lda", _ct.c_longlong), ("relu_y", _ct.c_void_p), ("ldy", _ct.c_longlong),
                ("a_mask", _ct.c_void_p), ("a_na", _ct.c_int), ("a_ne", _ct.c_int), ("a_mper", _ct.c_int),
                ("B", _ct.c_void_p), ("sbn", _ct.c_longlong), ("sbk", _ct.c_longlong), ("b_k_valid", _ct.c_int),
                ("bias", _ct.c_void_p), ("relu", _ct.c_int),
                ("c_mask", _ct.c_void_p), ("c_na", _ct.c_int), ("c_ne", _ct.c_int), ("c_mper", _ct.c_int),
                ("C", _ct.c_void_p), ("ldc", _ct.c_longlong), ("M", _ct.c_int),
                ("row_group", _ct.c_int), ("row_group_stride", _ct.c_int), ("accumulate", _ct.c_int), ("n_cols", _ct.c_int),
                ("k_len", _ct.c_int)]


class _WgradDesc(_ct.Structure):
    _fields_ = [("X", _ct.c_void_p), ("ldx", _ct.c_longlong), ("relu_y", _ct.c_void_p), ("ldy", _ct.c_longlong),
                ("x_mask", _ct.c_void_p), ("na", _ct.c_int), ("ne", _ct.c_int), ("mper", _ct.c_int),
                ("Y", _ct.c_void_p), ("ldyy", _ct.c_longlong), ("y_shift", _ct.c_int), ("y_period", _ct.c_int),
                ("dW", _ct.c_void_p), ("lddw", _ct.c_longlong), ("q_valid", _ct.c_int),
                ("db", _ct.c_void_p), ("M", _ct.c_int),
                ("p_cols", _ct.c_int), ("row_group", _ct.c_int), ("row_group_stride", _ct.c_int)]


MAX_GROUP = 8


def _group_ok(rows):
    return USE_TENSOR_CORES and len(rows) > 1 and min(rows) >= TC_MIN_ROWS


def linear_fwd_group(items):
    """items: list of (A, W, bias, out, relu, row_mask) with one (N, K) geometry -> as few launches as possible (<= 8 problems
    each); falls back to per-item linear_fwd when the shape is not tensor-core eligible."""
    A0, W0 = items[0][0], items[0][1]
    K, N = A0.shape[1], W0.shape[0]
    same = all(it[0].shape[1] == K and it[1].shape == W0.shape and bool(it[4]) == bool(items[0][4]) for it in items)
    sliced = _lib.load().refil_tc_gemm_k_slices(N, K) != 1
    if not (same and _group_ok([it[0].shape[0] for it in items]) and all(_tc_ok(it[0].shape[0], N, K) for it in items)) or \
            (sliced and any(it[2] is not None or it[4] or it[5] is not None for it in items)):
        for A, W, bias, out, relu, rm in items:
            linear_fwd(A, W, bias, out, relu=relu, row_mask=rm)
        return
    Kw = W0.shape[1]
    for lo in range(0, len(items), MAX_GROUP):
        part = items[lo:lo + MAX_GROUP]
        arr = (_GemmDesc * len(part))()
        for d, (A, W, bias, out, relu, rm) in zip(arr, part):
            cem, cna, cne, cmper = _rm(rm)
            M = A.shape[0]
            d.A, d.lda, d.relu_y, d.ldy = _p(A, F32), K, None, K
            d.a_mask, d.a_na, d.a_ne, d.a_mper = None, 1, 1, 1
            d.B, d.sbn, d.sbk, d.b_k_valid = _p(W, F32), Kw, 1, (Kw if Kw != K else 0)
            d.bias, d.relu = _p(bias, F32), int(relu)
            d.c_mask, d.c_na, d.c_ne, d.c_mper = cem, cna, cne, cmper
            d.C, d.ldc, d.M = _p(out, F32), out.shape[1], M
            _account("tc_gemm_tn", 2.0 * M * N * K, 4.0 * (M * K + M * N + N * K))
        _call("tc_gemm_tn_group", _ct.addressof(arr), len(part), N, K, shape=(len(part), part[0][0].shape[0], N, K),
              as_name="tc_gemm_tn")


def linear_bwd_data_group(items):
    """items: list of (dC, W, dA, relu_y, row_mask): dA = g(dC) W, grouped like linear_fwd_group."""
    dC0, W0 = items[0][0], items[0][1]
    N, K = dC0.shape[1], W0.shape[1]
    same = all(it[0].shape[1] == N and it[1].shape == W0.shape and (it[3] is None) == (items[0][3] is None) for it in items)
    if not (same and _group_ok([it[0].shape[0] for it in items]) and all(_tc_ok(it[0].shape[0], K, N) for it in items)):
        for dC, W, dA, relu_y, rm in items:
            linear_bwd_data(dC, W, dA, relu_y=relu_y, row_mask=rm)
        return
    Nw = W0.shape[0]
    for lo in range(0, len(items), MAX_GROUP):
        part = items[lo:lo + MAX_GROUP]
        arr = (_GemmDesc * len(part))()
        for d, (dC, W, dA, relu_y, rm) in zip(arr, part):
            aem, ana, ane, amper = _rm(rm)
            M = dC.shape[0]
            d.A, d.lda, d.relu_y, d.ldy = _p(dC, F32), N, _p(relu_y, F32), N
            d.a_mask, d.a_na, d.a_ne, d.a_mper = aem, ana, ane, amper
            d.B, d.sbn, d.sbk, d.b_k_valid = _p(W, F32), 1, K, (Nw if Nw != N else 0)     # "B"[j = k, i = n] = W[n * K + k]
            d.bias, d.relu = None, 0
            d.c_mask, d.c_na, d.c_ne, d.c_mper = None, 1, 1, 1
            d.C, d.ldc, d.M = _p(dA, F32), K, M
            _account("tc_gemm_tn", 2.0 * M * N * K, 4.0 * (M * N * (2 if relu_y is not None else 1) + M * K + N * K))
        _call("tc_gemm_tn_group", _ct.addressof(arr), len(part), K, N, shape=(len(part), part[0][0].shape[0], K, N),
              as_name="tc_gemm_tn")


def linear_bwd_weight_group(items):
    """items: list of (dC, A, dW, db, relu_y, row_mask): dW += g(dC)^T A, db += colsum g(dC), grouped."""
    dC0, A0, dW0 = items[0][0], items[0][1], items[0][2]
    N, K = dC0.shape[1], A0.shape[1]
    P, Qv = dW0.shape
    same = all(it[0].shape[1] == N and it[1].shape[1] == K and it[2].shape == dW0.shape and
               (it[3] is None) == (items[0][3] is None) and (it[4] is None) == (items[0][4] is None) for it in items)
    lib = _lib.load()
    if not (same and _group_ok([it[0].shape[0] for it in items]) and N % 4 == 0 and
            all(lib.refil_tc_wgrad_supported(it[0].shape[0], P, K) != 0 for it in items)):
        for dC, A, dW, db, relu_y, rm in items:
            linear_bwd_weight(dC, A, dW, db, relu_y=relu_y, row_mask=rm)
        return
    for lo in range(0, len(items), MAX_GROUP):
        part = items[lo:lo + MAX_GROUP]
        arr = (_WgradDesc * len(part))()
        for d, (dC, A, dW, db, relu_y, rm) in zip(arr, part):
            em, na, ne, mper = _rm(rm)
            M = dC.shape[0]
            d.X, d.ldx, d.relu_y, d.ldy = _p(dC, F32), N, _p(relu_y, F32), N
            d.x_mask, d.na, d.ne, d.mper = em, na, ne, mper
            d.Y, d.ldyy, d.y_shift, d.y_period = _p(A, F32), K, 0, 1
            d.dW, d.lddw, d.q_valid, d.db, d.M = _p(dW, F32), Qv, (Qv if Qv != K else 0), _p(db, F32), M
            _account("tc_gemm_wgrad", 2.0 * M * N * K, 4.0 * (M * N * (2 if relu_y is not None else 1) + M * K + N * K))
        _call("tc_gemm_wgrad_group", _ct.addressof(arr), len(part), P, K, shape=(len(part), part[0][0].shape[0], N, K),
              as_name="tc_gemm_wgrad")


def linear_fwd(A, W, bias, out, relu=False, row_mask=None):
    """out = [rowmask][relu](A W^T + bias).  A may be wider than W (zero-padded input columns, e.g. the packed fc1 input of 64
    columns against the 53-column weight): the weight is read in place with the tail of the reduction taken as zero."""
    M, K = A.shape
    N, Kw = W.shape
    if _tc_ok(M, N, K) and (_lib.load().refil_tc_gemm_k_slices(N, K) == 1 or (bias is None and not relu and row_mask is None)):
        return tc_gemm_tn(A, W, Kw, 1, out, N, K, bias=bias, relu=relu, c_row_mask=row_mask, k_valid=Kw if Kw != K else 0)
    if Kw != K:
        raise _lib.RefilError("linear_fwd: input has %d columns, weight %d (only the tensor-core path takes a padded input)" % (K, Kw))
    em, na, ne, mper = _rm(row_mask)
    _call("linear_fwd", _p(A, F32), K, _p(W, F32), W.shape[1], _p(bias, F32), _p(out, F32), N, M, N, K, int(relu),
          em, na, ne, mper)
    return out


def embed_fwd(entities, last_action, n_actions, W, bias, out, relu=True):
    M = entities.numel() // entities.shape[-1]
    ed = entities.shape[-1]
    _call("embed_fwd", _p(entities, F32), ed, _p(last_action, I32), n_actions, _p(W, F32), _p(bias, F32), _p(out, F32),
          M, W.shape[0], int(relu))
    return out


def linear_bwd_data(dC, W, dA, relu_y=None, row_mask=None):
    """dA = g(dC) W.  dC may be wider than W has rows (zero-padded output columns, e.g. the 14-action head in 32 columns)."""
    M, N = dC.shape
    Nw, K = W.shape
    if _tc_ok(M, K, N):      # dA[M,K] = g(dC)[M,N] W[N,K]: "B"[j=k, i=n] = W[n*K + k]
        return tc_gemm_tn(dC, W, 1, K, dA, K, N, relu_y=relu_y, a_row_mask=row_mask, k_valid=Nw if Nw != N else 0)
    if Nw != N:
        raise _lib.RefilError("linear_bwd_data: gradient has %d columns, weight %d rows" % (N, Nw))
    em, na, ne, mper = _rm(row_mask)
    _call("linear_bwd_data", _p(dC, F32), N, _p(relu_y, F32), N, em, na, ne, mper, _p(W, F32), K, _p(dA, F32), K, M, N, K)
    return dA


def linear_bwd_weight(dC, A, dW, db, relu_y=None, row_mask=None):
    """dW[P, Q] += g(dC)^T A, db += colsum g(dC).  dW may be narrower than A (zero-padded input columns: fc1) and have fewer
    rows than dC has columns (zero-padded output columns: the action head); it is accumulated in place."""
    M, N = dC.shape
    K = A.shape[1]
    P, Qv = dW.shape
    em, na, ne, mper = _rm(row_mask)
    if USE_TENSOR_CORES and M >= TC_MIN_ROWS and N % 4 == 0 and _lib.load().refil_tc_wgrad_supported(M, P, K) != 0:
        _account("tc_gemm_wgrad", 2.0 * M * N * K, 4.0 * (M * N * (2 if relu_y is not None else 1) + M * K + N * K))
        _call("tc_gemm_wgrad", _p(dC, F32), N, _p(relu_y, F32), N, em, na, ne, mper, _p(A, F32), K, 0, 1, _p(dW, F32), Qv,
              Qv if Qv != K else 0, _p(db, F32), M, P, K, shape=(M, N, K))
        return
    if (P, Qv) != (N, K):
        raise _lib.RefilError("linear_bwd_weight: padded operands need the tensor-core path")
    _call("linear_bwd_weight", _p(dC, F32), N, _p(relu_y, F32), N, em, na, ne, mper, _p(A, F32), K, _p(dW, F32), K,
          _p(db, F32), M, N, K)


def embed_bwd_weight(dC, relu_y, entities, last_action, n_actions, dW, db):
    M, N = dC.shape
    _call("embed_bwd_weight", _p(dC, F32), N, _p(relu_y, F32), N, _p(entities, F32), entities.shape[-1],
          _p(last_action, I32), n_actions, _p(dW, F32), _p(db, F32), M, N)


def gru_bwd_weight_hh(dGH, HS, n_agents, T, dWhh, dbhh):
    M, r = HS.shape
    if USE_TENSOR_CORES and M >= TC_MIN_ROWS and _lib.load().refil_tc_wgrad_supported(M, 3 * r, r) != 0:
        _call("tc_gemm_wgrad", _p(dGH, F32), 3 * r, None, 3 * r, None, 1, 1, 1, _p(HS, F32), r, n_agents, T,
              _p(dWhh, F32), r, 0, _p(dbhh, F32), M, 3 * r, r)
        return
    _call("gru_bwd_weight_hh", _p(dGH, F32), _p(HS, F32), n_agents, T, _p(dWhh, F32), _p(dbhh, F32), M, r)


def pack_inputs(entities, last_action, n_actions, out):
    rows = entities.numel() // entities.shape[-1]
    _call("pack_inputs", _p(entities, F32), entities.shape[-1], _p(last_action, I32), n_actions, _p(out, F32), rows,
          out.shape[-1])
    return out


def last_action_index(actions, out, n_entities):
    B, T, na = actions.shape[0], actions.shape[1], actions.shape[2]
    _call("last_action_index", _p(actions, I64), _p(out, I32), B, T, na, n_entities)
    return out


# ---------------------------------------------------------------------------------------------------- attention
def _copies(copies):
    m = [None, None, None]
    s = [0, 0, 0]
    mode = [0, 0, 0]
    for i, (mask, stride, md) in enumerate(copies):
        m[i] = _p(mask, U8)
        s[i] = int(stride)
        mode[i] = int(md)
    return m + s + mode


def masked_attn_fwd(qkv, out, copies, group_bits, entity_mask, N, T, ne, nq, d, H):
    # algorithmic bytes per (b, t, copy) unit = 4d(2ne + 2nq) + nq*ne, flops = 4*nq*ne*d
    _account("masked_attn_fwd", 4.0 * nq * ne * d * N * len(copies), (4.0 * d * (2 * ne + 2 * nq) + nq * ne) * N * len(copies))
    _call("masked_attn_fwd", _p(qkv, F32), _p(out, F32), *_copies(copies), _p(group_bits, U8), _p(entity_mask, U8),
          N, T, ne, nq, d, H, len(copies))
    return out


def masked_attn_bwd(qkv, dout, dqkv, copies, group_bits, entity_mask, N, T, ne, nq, d, H):
    _account("masked_attn_bwd", 10.0 * nq * ne * d * N * len(copies),
             (4.0 * d * (2 * ne + 2 * nq) * 2 + nq * ne) * N * len(copies))
    _call("masked_attn_bwd", _p(qkv, F32), _p(dout, F32), _p(dqkv, F32), *_copies(copies), _p(group_bits, U8),
          _p(entity_mask, U8), N, T, ne, nq, d, H, len(copies))
    return dqkv


POOL_TYPES = {"mean": 0, "max": 1}


def entity_pool_fwd(E, out, copies, group_bits, entity_mask, N, T, ne, nq, d, pool_type):
    """EntityPoolingLayer between in_trans and out_trans (attention.py:111-124): E [N*ne, d] -> out [C*N*nq, d]."""
    _account("entity_pool_fwd", 0.0, 4.0 * d * (ne + nq * len(copies)) * N)
    _call("entity_pool_fwd", _p(E, F32), _p(out, F32), *_copies(copies), _p(group_bits, U8), _p(entity_mask, U8),
          N, T, ne, nq, d, len(copies), POOL_TYPES[pool_type])
    return out


def entity_pool_bwd(E, dout, dE, copies, group_bits, entity_mask, N, T, ne, nq, d, pool_type):
    _account("entity_pool_bwd", 0.0, 4.0 * d * (2 * ne + nq * len(copies)) * N)
    _call("entity_pool_bwd", _p(E, F32), _p(dout, F32), _p(dE, F32), *_copies(copies), _p(group_bits, U8),
          _p(entity_mask, U8), N, T, ne, nq, d, len(copies), POOL_TYPES[pool_type])
    return dE


class _AttnDesc(_ct.Structure):
    _fields_ = [("qkv", _ct.c_void_p), ("out", _ct.c_void_p), ("dout", _ct.c_void_p), ("dqkv", _ct.c_void_p),
                ("mask0", _ct.c_void_p), ("mask1", _ct.c_void_p), ("mask2", _ct.c_void_p),
                ("s0", _ct.c_longlong), ("s1", _ct.c_longlong), ("s2", _ct.c_longlong),
                ("m0", _ct.c_int), ("m1", _ct.c_int), ("m2", _ct.c_int),
                ("group_bits", _ct.c_void_p), ("entity_mask", _ct.c_void_p), ("n_copies", _ct.c_int)]


def _attn_descs(items, fwd):
    arr = (_AttnDesc * len(items))()
    for d, it in zip(arr, items):
        if fwd:
            qkv, out, copies, group_bits, entity_mask = it
            d.out, d.dout, d.dqkv = _p(out, F32), None, None
        else:
            qkv, dout, dqkv, copies, group_bits, entity_mask = it
            d.out, d.dout, d.dqkv = None, _p(dout, F32), _p(dqkv, F32)
        c = _copies(copies)
        d.qkv = _p(qkv, F32)
        d.mask0, d.mask1, d.mask2, d.s0, d.s1, d.s2, d.m0, d.m1, d.m2 = c
        d.group_bits, d.entity_mask, d.n_copies = _p(group_bits, U8), _p(entity_mask, U8), len(copies)
    return arr


def masked_attn_fwd_group(items, N, T, ne, nq, d, H):
    """items: list of (qkv, out, copies, group_bits, entity_mask) of one geometry -> <= 8 problems per launch."""
    for lo in range(0, len(items), MAX_GROUP):
        part = items[lo:lo + MAX_GROUP]
        for it in part:
            C = len(it[2])
            _account("masked_attn_fwd", 4.0 * nq * ne * d * N * C, (4.0 * d * (2 * ne + 2 * nq) + nq * ne) * N * C)
        arr = _attn_descs(part, True)
        _call("masked_attn_fwd_group", _ct.addressof(arr), len(part), N, T, ne, nq, d, H, as_name="masked_attn_fwd")


def masked_attn_bwd_group(items, N, T, ne, nq, d, H):
    """items: list of (qkv, dout, dqkv, copies, group_bits, entity_mask)."""
    for lo in range(0, len(items), MAX_GROUP):
        part = items[lo:lo + MAX_GROUP]
        for it in part:
            C = len(it[3])
            _account("masked_attn_bwd", 10.0 * nq * ne * d * N * C, (4.0 * d * (2 * ne + 2 * nq) * 2 + nq * ne) * N * C)
        arr = _attn_descs(part, False)
        _call("masked_attn_bwd_group", _ct.addressof(arr), len(part), N, T, ne, nq, d, H, as_name="masked_attn_bwd")


# ---------------------------------------------------------------------------------------------------- GRU
def gru_scan_fwd(GI, Whh, bhh, h0, HS, gates, n_seq, T, na):
    r = Whh.shape[1]
    _call("gru_scan_fwd", _p(GI, F32), _p(Whh, F32), _p(bhh, F32), _p(h0, F32), _p(HS, F32), _p(gates, F32), n_seq, T,
          na, r)
    return HS


def gru_scan_bwd(dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, na):
    r = Whh.shape[1]
    _call("gru_scan_bwd", _p(dHS, F32), _p(gates, F32), _p(HS, F32), _p(h0, F32), _p(Whh, F32), _p(dGI, F32),
          _p(dGH, F32), n_seq, T, na, r)


# ---------------------------------------------------------------------------------------------------- mixers / TD
def mixer_fwd(kind, W1, B1, WF, V, q, qW, qI, qtot, qtot_im, N, na, me, w1_copies, imagine, softmax_w, tanh_nl,
              ingroup=None):
    _call("mixer_fwd", kind, _p(W1, F32), _p(B1, F32), _p(WF, F32), _p(V, F32), _p(q, F32), _p(qW, F32), _p(qI, F32),
          _p(qtot, F32), _p(qtot_im, F32), _p(ingroup, F32), N, na, me, w1_copies, int(imagine), int(softmax_w),
          int(tanh_nl))


def mixer_bwd(kind, W1, B1, WF, V, q, qW, qI, g_plain, g_im, dW1, dB1, dWF, dV, dq, dqW, dqI, N, na, me, w1_copies,
              imagine, softmax_w, tanh_nl):
    _call("mixer_bwd", kind, _p(W1, F32), _p(B1, F32), _p(WF, F32), _p(V, F32), _p(q, F32), _p(qW, F32), _p(qI, F32),
          _p(g_plain, F32), _p(g_im, F32), _p(dW1, F32), _p(dB1, F32), _p(dWF, F32), _p(dV, F32), _p(dq, F32),
          _p(dqW, F32), _p(dqI, F32), N, na, me, w1_copies, int(imagine), int(softmax_w), int(tanh_nl))


def gather_chosen(Q, actions, chosen, copies, rows_per_copy, A):
    _call("gather_chosen", _p(Q, F32), _p(actions, I64), _p(chosen, F32), copies, rows_per_copy, A)
    return chosen


def scatter_dq(dchosen, actions, dQ, copies, rows_per_copy, A, T, na):
    _call("scatter_dq", _p(dchosen, F32), _p(actions, I64), _p(dQ, F32), copies, rows_per_copy, A, T, na)
    return dQ


def target_max(q_online, q_target, avail, tgt, cur_max, rows, A, double_q):
    _call("target_max", _p(q_online, F32), _p(q_target, F32), _p(avail, I32), _p(tgt, F32), _p(cur_max, I64), rows, A,
          int(double_q))
    return tgt


def td_loss(qtot, qtot_im, tgt_tot, reward, terminated, filled, g_plain, g_im, targets_out, stats, B, T, gamma, lmbda):
    _call("td_loss", _p(qtot, F32), _p(qtot_im, F32), _p(tgt_tot, F32), _p(reward, F32), _p(terminated, U8),
          _p(filled, I64), _p(g_plain, F32), _p(g_im, F32), _p(targets_out, F32), _p(stats, F64), B, T, float(gamma),
          float(lmbda))


def grad_sumsq(grads, n_params, out):
    _call("grad_sumsq", _p(grads, F32), n_params, _p(out, F64))


def pack_stats(stats, tail, n):
    _call("pack_stats", _p(stats, F64), _p(tail, F32), n)


def clip_rmsprop_step(params, grads, square_avg, n_params, mask_sum, sumsq, grad_norm_out, clip, lr, alpha, eps, wd):
    _call("clip_rmsprop_step", _p(params, F32), _p(grads, F32), _p(square_avg, F32), n_params, _p(mask_sum, F32),
          _p(sumsq, F64), _p(grad_norm_out, F32), float(clip), float(lr), float(alpha), float(eps), float(wd))


# ---------------------------------------------------------------------------------------------------- acting / env
def ff_agent_act_supported(ne, na, ein, d, H, A):
    return _lib.load().refil_ff_agent_act_supported(ne, na, ein, d, H, A) != 0


def ff_agent_act(entities, actions, n_actions, obs_mask, entity_mask, params, q, t, select=None):
    """Fused acting forward of the FF entity-attention agent on timestep t of the EpisodeBatch tensors (read in place).
    params: dict with fc1 / attn.in_trans / attn.out_trans / fc2 weights.  q [E, na, A] is written.
    select: optional dict(avail [E,T,na,A] i32, actions_out [E,T,na,1] i64, u_pick, u_act, est_flags, epsilon, eps_dev) -> the
    epsilon-greedy choice is made in the same launch and written to actions_out[:, t]."""
    E, T, ne, ed = entities.shape
    na, A = q.shape[1], q.shape[2]
    sel = select or {}
    _call("ff_agent_act", _p(entities, F32), ed, _p(actions, I64), n_actions, _p(obs_mask, U8), obs_mask.shape[2],
          _p(entity_mask, U8), _p(params["fc1.weight"], F32), _p(params["fc1.bias"], F32), _p(params["attn.in_trans.weight"], F32),
          _p(params["attn.out_trans.weight"], F32), _p(params["attn.out_trans.bias"], F32), _p(params["fc2.weight"], F32),
          _p(params["fc2.bias"], F32), _p(q, F32), E, T, int(t), ne, na, A, _p(sel.get("avail"), I32), _p(sel.get("u_pick"), F32),
          _p(sel.get("u_act"), F32), _p(sel.get("est_flags"), I32), float(sel.get("epsilon", 0.0)), _p(sel.get("eps_dev"), F32),
          _p(sel.get("actions_out"), I64))
    return q


def select_actions(q, avail, u_pick, u_act, est_flags, epsilon, actions_out, B, na, A, eps_dev=None):
    """q [B, na, A] contiguous; avail / actions_out may be time slices of EpisodeBatch tensors (row stride taken from them)."""
    if avail.stride(-1) != 1 or avail.stride(-2) != A or actions_out.stride(-1) != 1:
        raise _lib.RefilError("select_actions: avail / actions_out must be contiguous in their last two dimensions")
    if not (avail.is_cuda and actions_out.is_cuda and avail.dtype == I32 and actions_out.dtype == I64):
        raise _lib.RefilError("select_actions: avail int32 / actions int64 CUDA tensors expected")
    _call("select_actions", _p(q, F32), na * A, avail.data_ptr(), avail.stride(0), _p(u_pick, F32), _p(u_act, F32),
          _p(est_flags, I32), float(epsilon), _p(eps_dev, F32), actions_out.data_ptr(), actions_out.stride(0), B, na, A)
    return actions_out
