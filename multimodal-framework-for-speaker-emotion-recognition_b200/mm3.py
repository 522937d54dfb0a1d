"""fp32-accurate dense products on the tcgen05 tensor cores for the TIME-PARALLEL parts of the path
(input projections, encoder projections, heads and the hoisted weight-gradient products — real dense
GEMMs over all T*N rows), through our own kernel ``lsthm_gemm3``
(csrc/gemm3_kernels.cuh; C ABI in include/lsthm_b200.h).

A single-pass TF32/bf16 product breaks the fp32 parity bar (TF32 on the LSTHM products
alone gives gradient errors of 1.4e-3), so the kernel splits every fp32 operand into two bf16 terms
while staging it and issues three UMMAs per k-step (a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, fp32 accumulate in
TMEM): ~2^-16 relative error per term, measured <2e-5 end to end against fp64.

(An earlier attempt to get the same effect from three library bf16 GEMMs over torch-side split/concat
copies was SLOWER than the fp32 SIMT SGEMM it replaced — 70.7 vs 52.6 ms/step — and was dropped.)

There is ONE backend: every fp32 CUDA product goes to ``lsthm_gemm3``.  Output widths that are not a multiple of 4 (the
6/7-class head) are zero-padded to the next multiple and sliced; an inner dimension that is not a multiple of 4 is an
error.  The plain torch expressions below are reached only for CPU tensors and fp64 (the parity tests' truth runs), where
no CUDA kernel applies.
"""
from __future__ import annotations


import torch
import torch.nn.functional as F

from . import _lib

launches = {"gemm3": 0}
# bench.py sets this to a list to collect (start_event, end_event, 2*M*N*K) per lsthm_gemm3 launch
events = None


def _gemm(mode, a, b, bias=None):
    launches["gemm3"] += 1
    if events is None:
        torch.cuda.nvtx.range_push("lsthm/K5_gemm3")         # K5 = the time-parallel GEMMs
        try:
            return _lib.gemm3(mode, a, b, bias)
        finally:
            torch.cuda.nvtx.range_pop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = _lib.gemm3(mode, a, b, bias)
    e1.record()
    k = a.shape[0] if mode == _lib.GEMM_TN else a.shape[1]
    events.append((e0, e1, 2.0 * c.shape[0] * c.shape[1] * k))
    return c


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D view with unit inner stride and 16-byte aligned rows (copy only if the layout forces it)."""
    if t.stride(1) != 1 or t.stride(0) % 4:
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _ok(*ts: torch.Tensor) -> bool:
    """True when the operands are what the CUDA kernel computes on (fp32 on a CUDA device).  CPU tensors and fp64 —
    the tests' truth runs — take the equivalent torch expression; nothing else does."""
    return all(t.is_cuda and t.dtype == torch.float32 for t in ts)


def _need4(name: str, *dims: int) -> None:
    if any(d % 4 for d in dims):
        raise RuntimeError(f"{name}: lsthm_gemm3 needs matrix widths that are multiples of 4, got {dims} (pad the operand)")


def mm_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a[K,M]^T @ b[K,N]  (weight-gradient product: reduction over the T*N rows)."""
    if not _ok(a, b):
        return a.t() @ b
    _need4("mm_tn", a.shape[1], b.shape[1])
    return _gemm(_lib.GEMM_TN, _rows(a), _rows(b))


def mm_nn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a[M,K] @ b[K,N]."""
    if not _ok(a, b):
        return a @ b
    _need4("mm_nn", a.shape[1], b.shape[1])
    return _gemm(_lib.GEMM_NN, _rows(a), _rows(b))


def mm_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a[M,K] @ b[N,K]^T."""
    if not _ok(a, b):
        return a @ b.t()
    _need4("mm_nt", a.shape[1], b.shape[0])
    return _gemm(_lib.GEMM_NT, _rows(a), _rows(b))


def colsum(a: torch.Tensor) -> torch.Tensor:
    """a.sum(0) for a 2-D matrix — the bias gradient of a Linear over all T*N rows — on our deterministic kernel."""
    if a.is_cuda and a.dtype == torch.float32 and a.dim() == 2 and a.shape[0] > 0:
        _need4("colsum", a.shape[1])
        launches["gemm3"] += 2
        return _lib.colsum(_rows(a))
    return a.sum(0)


class _LinearTC(torch.autograd.Function):
    """y = x W^T (+ b) (optionally followed by ReLU in the GEMM epilogue) on 2-D row matrices."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        ctx.has_bias, ctx.relu = bias is not None, relu
        y = _gemm(_lib.GEMM_NT_RELU if relu else _lib.GEMM_NT, _rows(x), _rows(weight), None if bias is None else bias.contiguous())
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        if ctx.relu:
            dy = torch.ops.aten.threshold_backward(dy, y, 0.0)
        dy = _rows(dy)
        dx = mm_nn(dy, weight) if ctx.needs_input_grad[0] else None
        dw = mm_tn(dy, x) if ctx.needs_input_grad[1] else None
        db = colsum(dy) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None


def linear_into(x: torch.Tensor, weight: torch.Tensor, bias, out: torch.Tensor) -> None:
    """out[:] = x @ weight^T + bias for 2-D operands, written straight into ``out`` (may be a column block of a wider
    matrix); no autograd (used inside custom backward/forward bodies)."""
    if _ok(x, weight, out):
        _need4("linear_into", x.shape[1], weight.shape[0])
        if out.stride(1) != 1 or out.stride(0) % 4 or out.data_ptr() % 16:
            raise RuntimeError("linear_into: the output view needs unit inner stride and 16-byte aligned rows")
        launches["gemm3"] += 1
        _lib.gemm3(_lib.GEMM_NT, _rows(x), _rows(weight), None if bias is None else bias.contiguous(), out=out)
    else:
        out.copy_(F.linear(x, weight, bias))


class _LinearCat(torch.autograd.Function):
    """[x_1 W_1^T + b_1 | x_2 W_2^T + b_2 | ...] written straight into the column blocks of ONE output matrix (no torch.cat /
    torch.stack copy of the T*N-row results).  Inputs: n row matrices, then n weights, then n biases."""

    @staticmethod
    def forward(ctx, n, *args):
        xs, ws, bs = args[:n], args[n:2 * n], args[2 * n:3 * n]
        widths = [w.shape[0] for w in ws]
        out = torch.empty(xs[0].shape[0], sum(widths), device=xs[0].device, dtype=torch.float32)
        o = 0
        for x, w, b, wd in zip(xs, ws, bs, widths):
            launches["gemm3"] += 1
            _lib.gemm3(_lib.GEMM_NT, _rows(x), _rows(w), b.contiguous(), out=out[:, o:o + wd])
            o += wd
        ctx.n, ctx.widths = n, widths
        ctx.save_for_backward(*xs, *ws)
        return out

    @staticmethod
    def backward(ctx, dy):
        n = ctx.n
        xs, ws = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        dy = _rows(dy)
        gx, gw, gb, o = [], [], [], 0
        for i, (x, w, wd) in enumerate(zip(xs, ws, ctx.widths)):
            blk = dy[:, o:o + wd]
            gx.append(mm_nn(blk, w) if ctx.needs_input_grad[1 + i] else None)
            gw.append(mm_tn(blk, x) if ctx.needs_input_grad[1 + n + i] else None)
            gb.append(colsum(blk) if ctx.needs_input_grad[1 + 2 * n + i] else None)
            o += wd
        return (None, *gx, *gw, *gb)


class _LinearBlock(torch.autograd.Function):
    """One column block of ``linear_cat`` as its own autograd node, for callers that build the blocks on different streams
    (mab_net._EncoderBranches): forward writes ``x W^T + b`` into ``out`` — a column block of a wider matrix the caller owns —
    and returns a zero-size handle; backward takes the block's gradient (a strided column block of the wide gradient) from
    ``box["dy"]``, which the caller sets before it differentiates the handle."""

    @staticmethod
    def forward(ctx, x, weight, bias, out, box):
        _need4("linear_block", x.shape[1], weight.shape[0])
        launches["gemm3"] += 1
        _lib.gemm3(_lib.GEMM_NT, _rows(x), _rows(weight), bias.contiguous(), out=out)
        ctx.save_for_backward(x, weight)
        ctx.box = box
        return x.new_empty(0)

    @staticmethod
    def backward(ctx, _):
        x, weight = ctx.saved_tensors
        blk = ctx.box.pop("dy")
        dx = mm_nn(blk, weight) if ctx.needs_input_grad[0] else None
        dw = mm_tn(blk, x) if ctx.needs_input_grad[1] else None
        db = colsum(blk) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


def linear_cat(xs, weights, biases) -> torch.Tensor:
    """cat([F.linear(x_i, W_i, b_i)], dim=-1) for inputs that share their leading dimensions ([..., K_i] each)."""
    lead = xs[0].shape[:-1]
    if not all(_ok(x, w) for x, w in zip(xs, weights)):
        return torch.cat([F.linear(x, w, b) for x, w, b in zip(xs, weights, biases)], dim=-1)
    for x, w in zip(xs, weights):
        _need4("linear_cat", x.shape[-1], w.shape[0])
    rows = [x.reshape(-1, x.shape[-1]) for x in xs]
    y = _LinearCat.apply(len(xs), *rows, *weights, *biases)
    return y.view(*lead, y.shape[-1])


def rows_view(x: torch.Tensor):
    """A 3-D activation [A, B, K] as a 2-D row matrix WITHOUT a copy when its storage allows it.

    Returns (rows[R, K], swapped): ``swapped`` is True when the rows are in [B, A] order, i.e. ``x`` is a permuted
    view of time-major storage such as ``x_time_major.permute(1, 0, 2)`` (the reference hands its encoders exactly
    that, model/HybridRNN_ATV.py:90-92).  Row-wise ops don't care about the order; whoever needs it (the attention
    kernel) is told through its row strides.  None if no copy-free row view exists.
    """
    A, B, K = x.shape
    if x.stride(2) != 1:
        return None
    s0, s1 = x.stride(0), x.stride(1)
    if s0 == B * s1 and s1 % 4 == 0 and s1 >= K:
        return x.as_strided((A * B, K), (s1, 1)), False
    if s1 == A * s0 and s0 % 4 == 0 and s0 >= K:
        return x.as_strided((B * A, K), (s0, 1)), True
    return None


def linear3(x: torch.Tensor, weight: torch.Tensor, bias=None, relu: bool = False) -> torch.Tensor:
    """Drop-in for F.linear (+ optional ReLU) on the time-parallel projections (x [..., K], weight [N, K]).
    3-D inputs that are permuted views of time-major storage are processed in place (no permute copy) and the
    result is returned as the same kind of view."""
    K, N = x.shape[-1], weight.shape[0]
    if not _ok(x, weight):
        y = F.linear(x, weight, bias)
        return F.relu(y) if relu else y
    _need4("linear3 (inner dimension)", K)
    if N % 4:
        # e.g. the 6/7-class head (nn_out.3): zero rows up to the next multiple of 4, sliced away again (autograd slices the
        # padded weight's gradient back through F.pad)
        pad = (-N) % 4
        y = linear3(x, F.pad(weight, (0, 0, 0, pad)), None if bias is None else F.pad(bias, (0, pad)), relu)
        return y[..., :N]
    if x.dim() == 3 and x.data_ptr() % 16 == 0:
        rv = rows_view(x)
        if rv is not None:
            rows, swapped = rv
            y = _LinearTC.apply(rows, weight, bias, relu)
            A, B = x.shape[0], x.shape[1]
            return y.view(B, A, N).transpose(0, 1) if swapped else y.view(A, B, N)
    lead = x.shape[:-1]
    y = _LinearTC.apply(x.reshape(-1, K), weight, bias, relu)
    return y.view(*lead, N)
