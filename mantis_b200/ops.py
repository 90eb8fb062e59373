"""torch-facing wrappers of the C-ABI kernels (mantis_b200/csrc) + autograd Functions.

PyTorch is plumbing here (device memory, streams, autograd graph); every arithmetic op of the hot path is one of
the hand-written CUDA kernels.  There is no CPU path: tensors must live on an sm_100a device.
"""
import ctypes
import math
import os

import torch

from . import _lib
from ._lib import check

F32, BF16 = 0, 1
ACT_KINDS = {"gelu": 0, "gelu_pytorch_tanh": 1, "gelu_new": 1, "gelu_tanh": 1, "quick_gelu": 2}
# epilogue codes of the tcgen05 GEMM: 0 none, 1 erf, 2 tanh, 3 quick
_GEMM_ACT = {None: 0, "gelu": 1, "gelu_pytorch_tanh": 2, "gelu_new": 2, "gelu_tanh": 2, "quick_gelu": 3}

# Minimum M*N*K for the tensor-core GEMM (below this the SIMT kernel is as fast and supports any alignment).
FAST_GEMM_MIN_WORK = int(os.environ.get("MB200_FAST_GEMM_MIN_WORK", str(128 * 128 * 64)))
FORCE_GENERIC = os.environ.get("MB200_FORCE_GENERIC", "0") == "1"
GEMM_2CTA = os.environ.get("MB200_GEMM_2CTA", "1") == "1"      # CTA-pair tiles for large problems

launch_count = 0   # kernels (C-ABI calls) issued; bench.py reports it as gpu_launches


def _L():
    return _lib.lib()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"mantis_b200 kernels support float32 and bfloat16, got {t.dtype}")


def _p(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MantisB200Error("mantis_b200 ops need CUDA tensors (there is no CPU fallback)")


def _call(name, *args):
    global launch_count
    launch_count += 1
    rc = getattr(_L(), name)(*args)
    check(rc, name)


# ---------------------------------------------------------------------------------------------- GEMM / linear
def _fast_ok(a, b, M, N, K):
    if FORCE_GENERIC or a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        return False
    if M * N * K < FAST_GEMM_MIN_WORK or K < 8:
        return False
    return (a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)


def gemm(a, b, trans_a=False, trans_b=True, bias=None, act=None, addend=None, out=None):
    """out[M,N] = act(op(a) @ op(b) + bias) + addend.   a: [M,K] (or [K,M] if trans_a);  b: [N,K] if trans_b
    (nn.Linear weight layout) else [K,N].  2-D tensors with unit inner stride."""
    _need_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1, "gemm operands must be 2-D, inner-contiguous"
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[0], b.shape[1]) if trans_b else (b.shape[1], b.shape[0])
    assert K == Kb, f"gemm K mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    assert out.stride(1) == 1
    if M == 0 or N == 0:
        return out
    if bias is not None:
        bias = bias.contiguous()
    if addend is not None:
        assert addend.shape == out.shape and addend.stride(1) == 1
    if (M <= 16 and not trans_a and trans_b and act is None and not FORCE_GENERIC and a.dtype == torch.bfloat16
            and b.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and K % 8 == 0 and N * K >= 65536
            and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        _call("mb200_skinny_gemm_bf16", _p(a), _p(b), _p(out), _p(bias), _p(addend), M, N, K, a.stride(0), b.stride(0),
              out.stride(0), addend.stride(0) if addend is not None else 0, _st())
        return out
    if out.dtype == torch.float32 and a.dtype == torch.bfloat16:
        # fp32 destination of a bf16 product: the wgrad into the fp32 main-gradient buffer (C32 (+)= op(a) op(b))
        assert bias is None and act is None and (addend is None or addend.data_ptr() == out.data_ptr()), \
            "the fp32-output GEMM only accumulates into its own destination"
        if _fast_ok(a, b, M, N, K) and out.stride(0) % 4 == 0 and out.data_ptr() % 16 == 0:
            _call("mb200_gemm_bf16_acc32", _p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                  int(trans_a), int(trans_b), int(addend is not None), _st())
            return out
        _call("mb200_gemm_generic", _p(a), _p(b), _p(out), None, M, N, K, a.stride(0), b.stride(0), out.stride(0),
              int(trans_a), int(trans_b), 1.0, 1.0 if addend is not None else 0.0, 1, 0, 0, 0, _dt(a), _dt(out), _st())
        return out
    if _fast_ok(a, b, M, N, K) and out.dtype == torch.bfloat16:
        fn = "mb200_gemm_bf16_2cta" if (GEMM_2CTA and M >= 512 and N >= 512) else "mb200_gemm_bf16"
        _call(fn, _p(a), _p(b), _p(out), _p(bias), _p(addend), M, N, K, a.stride(0), b.stride(0),
              out.stride(0), addend.stride(0) if addend is not None else 0, int(trans_a), int(trans_b),
              _GEMM_ACT[act], _st())
        return out
    # generic SIMT path
    beta = 0.0
    if addend is not None and act is None:
        if addend.data_ptr() != out.data_ptr():
            out.copy_(addend)
        beta = 1.0
    _call("mb200_gemm_generic", _p(a), _p(b), _p(out), _p(bias), M, N, K, a.stride(0), b.stride(0), out.stride(0),
          int(trans_a), int(trans_b), 1.0, beta, 1, 0, 0, 0, _dt(a), _dt(out), _st())
    if act is not None:
        n = out.numel()
        assert out.is_contiguous() and n % 8 == 0
        _call("mb200_act_fwd", _p(out), _p(out), n, ACT_KINDS[act], _dt(out), _st())
        if addend is not None:
            _call("mb200_add", _p(out), _p(addend.contiguous()), _p(out), n, _dt(out), _st())
    return out


def colsum(x2d, out=None, accumulate=False):
    """column sums of [n, N] -> [N] (bias gradients), fp32 accumulation"""
    n, N = x2d.shape
    assert x2d.stride(1) == 1
    parts = _L().mb200_colsum_parts(n)
    part = torch.empty((parts, N), dtype=torch.float32, device=x2d.device)
    if out is None:
        out = torch.empty((N,), dtype=x2d.dtype, device=x2d.device)
    _call("mb200_colsum", _p(x2d), _p(part), _p(out), int(accumulate), n, N, x2d.stride(0), _dt(x2d), _st())
    return out


WGRAD_SIDE_STREAM = os.environ.get("MB200_WGRAD_STREAM", "1") == "1"     # measured: 1626 -> 1600 ms/step (B200, config 2)
_side_streams = {}


def wgrad_stream(device):
    """the side stream the weight-gradient GEMMs run on when MB200_WGRAD_STREAM=1 (None otherwise)"""
    if not WGRAD_SIDE_STREAM:
        return None
    key = torch.device(device).index or 0
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def join_wgrad_stream(device):
    """make the current stream wait for every weight-gradient GEMM issued so far (before the gradients are reduced / consumed)"""
    st = wgrad_stream(device)
    if st is not None:
        torch.cuda.current_stream(device).wait_stream(st)


def _wgrad_into_main(g2, x2, w):
    """main_grad += g2^T @ x2 for a weight whose gradient lives in a B200Trainer flat buffer (`w._b200_main_grad`, an
    [out, in] view; fp32 by default): accumulation happens in the wgrad epilogue, in the buffer's own precision.  The trainer
    never zero-fills the buffer: it marks every parameter "fresh" after an optimizer step and the first gradient written
    afterwards overwrites (B200Trainer._mark_fresh / _fold_grad)."""
    mg = w._b200_main_grad
    fresh = getattr(w, "_b200_grad_fresh", False)      # first product since the optimizer step: write, do not accumulate
    side = wgrad_stream(mg.device)
    if side is None:
        gemm(g2, x2, trans_a=True, trans_b=False, addend=None if fresh else mg, out=mg)
    else:
        # The weight gradient is off the critical path of backward (nothing reads it before the optimizer): issued on a side
        # stream, its persistent CTAs take the SMs that the ragged last wave of the dgrad GEMMs leaves idle, and vice versa.
        side.wait_stream(torch.cuda.current_stream(mg.device))
        with torch.cuda.stream(side):
            gemm(g2, x2, trans_a=True, trans_b=False, addend=None if fresh else mg, out=mg)
        g2.record_stream(side); x2.record_stream(side)   # autograd frees them on the main stream; the allocator must wait
    w._b200_grad_fresh = False


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, residual):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        res2 = residual.reshape(-1, weight.shape[0]) if residual is not None else None
        need_pre = act is not None and (x.requires_grad or weight.requires_grad)
        if need_pre:
            pre = gemm(x2, weight, bias=bias)
            y = activation_fwd(pre, act)
            if res2 is not None:
                y = add(y, res2)
        else:
            pre = None
            y = gemm(x2, weight, bias=bias, act=act, addend=res2)
        ctx.save_for_backward(x2, weight, pre)
        ctx.weight_ref = weight if getattr(weight, "_b200_main_grad", None) is not None else None
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.in_shape = shp
        return y.reshape(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight, pre = ctx.saved_tensors
        g2 = gy.reshape(-1, weight.shape[0])
        if g2.stride(1) != 1:
            g2 = g2.contiguous()
        gres = gy if ctx.has_res else None
        if ctx.act is not None:
            g2 = activation_bwd(pre, g2, ctx.act)
        gx = gw = gb = None
        if ctx.needs_input_grad[1]:                  # first: it may run on the side stream, next to the dgrad below
            wref = ctx.weight_ref
            if wref is not None:
                # gradient accumulation fused into the wgrad epilogue: main_grad += dy^T @ x (no temporary, no extra pass;
                # main_grad is the trainer's flat buffer, fp32 by default)
                _wgrad_into_main(g2, x2, wref)
            else:
                gw = gemm(g2, x2, trans_a=True, trans_b=False)                             # dW = dy^T @ x
        if ctx.needs_input_grad[0]:
            gx = gemm(g2, weight, trans_a=False, trans_b=False).reshape(ctx.in_shape)      # dx = dy @ W
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(g2)
        return gx, gw, gb, None, gres


class _MultiLinearFn(torch.autograd.Function):
    """(x W1^T, x W2^T[, x W3^T]) for bias-free projections that share their input (q/k/v, gate/up).  Same GEMMs as separate
    linears in forward; in backward the input gradient is accumulated inside the dgrad epilogues (dx = g1 W1; dx += g2 W2;
    ...) instead of autograd summing three tensors with extra elementwise passes."""

    @staticmethod
    def forward(ctx, x, *weights):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, *weights)
        ctx.wrefs = [w if getattr(w, "_b200_main_grad", None) is not None else None for w in weights]
        ctx.in_shape = shp
        return tuple(gemm(x2, w).reshape(*shp[:-1], w.shape[0]) for w in weights)

    @staticmethod
    def backward(ctx, *gys):
        x2, *weights = ctx.saved_tensors
        gx = None
        gws = []
        for i, (w, gy) in enumerate(zip(weights, gys)):
            g2 = gy.reshape(-1, w.shape[0])
            if g2.stride(1) != 1:
                g2 = g2.contiguous()
            gw = None
            if ctx.needs_input_grad[1 + i]:
                wref = ctx.wrefs[i]
                if wref is not None:
                    _wgrad_into_main(g2, x2, wref)
                else:
                    gw = gemm(g2, x2, trans_a=True, trans_b=False)
            gws.append(gw)
            if ctx.needs_input_grad[0]:
                if gx is None:
                    gx = gemm(g2, w, trans_a=False, trans_b=False)
                else:
                    gemm(g2, w, trans_a=False, trans_b=False, addend=gx, out=gx)
        return (gx.reshape(ctx.in_shape) if gx is not None else None, *gws)


def multi_linear(x, *weights):
    return _MultiLinearFn.apply(x, *weights)


def linear(x, weight, bias=None, act=None, residual=None):
    """y = act(x @ weight^T + bias) + residual  (weight in nn.Linear [out, in] layout)"""
    return _LinearFn.apply(x, weight, bias, act, residual)


# ---------------------------------------------------------------------------------------------- elementwise
def activation_fwd(x, kind):
    x = x.contiguous()
    y = torch.empty_like(x)
    _call("mb200_act_fwd", _p(x), _p(y), x.numel(), ACT_KINDS[kind], _dt(x), _st())
    return y


def activation_bwd(x, dy, kind):
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    _call("mb200_act_bwd", _p(x), _p(dy), _p(dx), x.numel(), ACT_KINDS[kind], _dt(x), _st())
    return dx


class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        ctx.save_for_backward(x)
        ctx.kind = kind
        return activation_fwd(x, kind)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return activation_bwd(x.contiguous(), dy, ctx.kind), None


def activation(x, kind):
    return _ActFn.apply(x, kind)


def add(a, b):
    a = a.contiguous(); b = b.contiguous()
    assert a.shape == b.shape and a.numel() % 8 == 0
    y = torch.empty_like(a)
    _call("mb200_add", _p(a), _p(b), _p(y), a.numel(), _dt(a), _st())
    return y


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def residual_add(a, b):
    return _AddFn.apply(a, b)


class _SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        gate = gate.contiguous(); up = up.contiguous()
        out = torch.empty_like(gate)
        _call("mb200_swiglu_fwd", _p(gate), _p(up), _p(out), gate.numel(), _dt(gate), _st())
        ctx.save_for_backward(gate, up)
        return out

    @staticmethod
    def backward(ctx, dout):
        gate, up = ctx.saved_tensors
        dout = dout.contiguous()
        dg = torch.empty_like(gate); du = torch.empty_like(up)
        _call("mb200_swiglu_bwd", _p(gate), _p(up), _p(dout), _p(dg), _p(du), gate.numel(), _dt(gate), _st())
        return dg, du


def swiglu(gate, up):
    """silu(gate) * up"""
    return _SwigluFn.apply(gate, up)


FUSED_SWIGLU_MLP = os.environ.get("MB200_FUSED_SWIGLU_MLP", "1") == "1"


def swiglu_mlp_ok(x, wg, wu, wd):
    """can down(silu(x Wg^T) * (x Wu^T)) take the fused-epilogue path?  (bf16, TMA-aligned, large enough for the tensor-core GEMM)"""
    if not FUSED_SWIGLU_MLP or FORCE_GENERIC or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    D, I = x.shape[-1], wg.shape[0]
    M = x.numel() // D
    if any(w.dtype != torch.bfloat16 or w.dim() != 2 or not w.is_contiguous() or w.data_ptr() % 16 for w in (wg, wu, wd)):
        return False
    return (M * I * D >= FAST_GEMM_MIN_WORK and D % 8 == 0 and I % 8 == 0 and M > 16
            and wg.shape == (I, D) and wu.shape == (I, D) and wd.shape == (D, I))


class _SwigluMLPFn(torch.autograd.Function):
    """y = (silu(x Wg^T) * (x Wu^T)) Wd^T (+ residual), the LLaMA / Mistral MLP (llama/modeling_llama.py:171-184), as FOUR tensor-core
    launches forward (gate GEMM, up GEMM with the SwiGLU in its epilogue, down GEMM with the residual in its epilogue) and the
    SwiGLU backward inside the down-projection's dgrad epilogue: the two elementwise kernels of the unfused path (3 and 5 passes
    over [tokens, intermediate]) never run, and d(act) never exists in memory."""

    @staticmethod
    def forward(ctx, x, wg, wu, wd, residual):
        shp = x.shape
        D, I = shp[-1], wg.shape[0]
        x2 = x.reshape(-1, D)
        if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        M = x2.shape[0]
        g = gemm(x2, wg)
        u = torch.empty((M, I), dtype=x.dtype, device=x.device)
        a = torch.empty((M, I), dtype=x.dtype, device=x.device)
        _call("mb200_gemm_bf16_swiglu_fwd", _p(x2), _p(wu), _p(g), _p(u), _p(a), M, I, D, x2.stride(0), wu.stride(0), I, _st())
        res2 = residual.reshape(-1, wd.shape[0]) if residual is not None else None
        y = gemm(a, wd, addend=res2)
        ctx.save_for_backward(x2, wg, wu, wd, g, u, a)
        ctx.wrefs = [w if getattr(w, "_b200_main_grad", None) is not None else None for w in (wg, wu, wd)]
        ctx.in_shape, ctx.has_res = shp, residual is not None
        return y.reshape(*shp[:-1], wd.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, wg, wu, wd, g, u, a = ctx.saved_tensors
        Dout, I = wd.shape
        g2 = gy.reshape(-1, Dout)
        if g2.stride(1) != 1 or g2.stride(0) % 8 or g2.data_ptr() % 16:
            g2 = g2.contiguous()
        M = g2.shape[0]
        def wgrad(i, dy_, x_):
            if not ctx.needs_input_grad[1 + i]:
                return None
            if ctx.wrefs[i] is not None:
                _wgrad_into_main(dy_, x_, ctx.wrefs[i])
                return None
            return gemm(dy_, x_, trans_a=True, trans_b=False)

        gwd = wgrad(2, g2, a)                          # needs nothing computed here: may overlap the fused dgrad below
        dg = torch.empty_like(g); du = torch.empty_like(u)
        _call("mb200_gemm_bf16_swiglu_bwd", _p(g2), _p(wd), _p(g), _p(u), _p(dg), _p(du), M, I, Dout, g2.stride(0), wd.stride(0),
              I, _st())
        grads_w = [wgrad(0, dg, x2), wgrad(1, du, x2), gwd]
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm(dg, wg, trans_a=False, trans_b=False)
            gemm(du, wu, trans_a=False, trans_b=False, addend=gx, out=gx)
            gx = gx.reshape(ctx.in_shape)
        return gx, grads_w[0], grads_w[1], grads_w[2], (gy if ctx.has_res else None)


def swiglu_mlp(x, wg, wu, wd, residual=None):
    return _SwigluMLPFn.apply(x, wg, wu, wd, residual)


# ---------------------------------------------------------------------------------------------- norms
class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        n, D = x2.shape
        y = torch.empty_like(x2)
        rstd = torch.empty((n,), dtype=torch.float32, device=x.device)
        w = w.contiguous()
        _call("mb200_rmsnorm_fwd", _p(x2), _p(w), _p(y), _p(rstd), n, D, float(eps), _dt(x2), _st())
        ctx.save_for_backward(x2, w, rstd)
        ctx.shp = shp
        return y.reshape(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, rstd = ctx.saved_tensors
        n, D = x2.shape
        dy2 = dy.reshape(-1, D).contiguous()
        dx = torch.empty_like(x2)
        parts = _L().mb200_norm_bwd_parts(n)
        need_w = ctx.needs_input_grad[1]
        dw_part = torch.empty((parts, D), dtype=torch.float32, device=x2.device) if need_w else None
        dw = torch.empty_like(w) if need_w else None
        _call("mb200_rmsnorm_bwd", _p(x2), _p(w), _p(dy2), _p(rstd), _p(dx), _p(dw_part), _p(dw), 0, 0, n, D,
              _dt(x2), _st())
        return dx.reshape(ctx.shp), dw, None


def rms_norm(x, weight, eps):
    return _RMSNormFn.apply(x, weight, eps)


class _RMSNormResFn(torch.autograd.Function):
    """(rms_norm(x), x): the pre-norm residual pattern `h = x + sublayer(norm(x))` (llama/modeling_llama.py decoder layer).
    The second output is x itself, handed to the residual branch; in backward the gradient arriving through that branch is
    added inside the norm's dx pass (mb200_rmsnorm_bwd_res) instead of by a separate elementwise kernel of autograd."""

    @staticmethod
    def forward(ctx, x, w, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        n, D = x2.shape
        y = torch.empty_like(x2)
        rstd = torch.empty((n,), dtype=torch.float32, device=x.device)
        w = w.contiguous()
        _call("mb200_rmsnorm_fwd", _p(x2), _p(w), _p(y), _p(rstd), n, D, float(eps), _dt(x2), _st())
        ctx.save_for_backward(x2, w, rstd)
        ctx.shp = shp
        return y.reshape(shp), x.view(shp)

    @staticmethod
    def backward(ctx, dy, dres):
        x2, w, rstd = ctx.saved_tensors
        n, D = x2.shape
        dy2 = dy.reshape(-1, D).contiguous()
        dr2 = dres.reshape(-1, D).contiguous() if dres is not None else None
        dx = torch.empty_like(x2)
        parts = _L().mb200_norm_bwd_parts(n)
        need_w = ctx.needs_input_grad[1]
        dw_part = torch.empty((parts, D), dtype=torch.float32, device=x2.device) if need_w else None
        dw = torch.empty_like(w) if need_w else None
        _call("mb200_rmsnorm_bwd_res", _p(x2), _p(w), _p(dy2), _p(rstd), _p(dr2), _p(dx), _p(dw_part), _p(dw), 0, n, D,
              _dt(x2), _st())
        return dx.reshape(ctx.shp), dw, None


def rms_norm_res(x, weight, eps):
    """-> (rms_norm(x), x) with the residual branch's gradient folded into the norm's backward"""
    return _RMSNormResFn.apply(x, weight, eps)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        n, D = x2.shape
        y = torch.empty_like(x2)
        need = x.requires_grad or w.requires_grad
        mean = torch.empty((n,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((n,), dtype=torch.float32, device=x.device)
        w = w.contiguous(); b = b.contiguous() if b is not None else None
        _call("mb200_layernorm_fwd", _p(x2), _p(w), _p(b), _p(y), _p(mean), _p(rstd), n, D, float(eps), _dt(x2), _st())
        if need:
            ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shp = shp
        ctx.has_b = b is not None
        return y.reshape(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        n, D = x2.shape
        dy2 = dy.reshape(-1, D).contiguous()
        dx = torch.empty_like(x2)
        parts = _L().mb200_norm_bwd_parts(n)
        dw_part = torch.empty((parts, D), dtype=torch.float32, device=x2.device)
        db_part = torch.empty((parts, D), dtype=torch.float32, device=x2.device)
        dw = torch.empty_like(w)
        db = torch.empty_like(w) if ctx.has_b else None
        _call("mb200_layernorm_bwd", _p(x2), _p(w), _p(dy2), _p(mean), _p(rstd), _p(dx), _p(dw_part), _p(db_part),
              _p(dw), _p(db), 0, n, D, _dt(x2), _st())
        return dx.reshape(ctx.shp), dw, db, None


def layer_norm(x, weight, bias, eps):
    return _LayerNormFn.apply(x, weight, bias, eps)


# ---------------------------------------------------------------------------------------------- embedding
class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table):
        ids_c = ids.contiguous().to(torch.int64)
        V, D = table.shape
        out = torch.empty((*ids.shape, D), dtype=table.dtype, device=table.device)
        _call("mb200_embedding_fwd", _p(ids_c), _p(table), _p(out), ids_c.numel(), D, V, _dt(table), _st())
        ctx.save_for_backward(ids_c)
        ctx.V, ctx.D = V, D
        return out

    @staticmethod
    def backward(ctx, gout):
        (ids_c,) = ctx.saved_tensors
        gout = gout.contiguous()
        gt = torch.zeros((ctx.V, ctx.D), dtype=gout.dtype, device=gout.device)
        _call("mb200_embedding_bwd", _p(ids_c), _p(gout), _p(gt), ids_c.numel(), ctx.D, ctx.V, _dt(gout), _st())
        return None, gt


def embedding(ids, table):
    _need_cuda(ids, table)
    assert table.is_contiguous()
    return _EmbeddingFn.apply(ids, table)


# ---------------------------------------------------------------------------------------------- RoPE
def rope_apply(x, pos, inv_freq, attn_scaling=1.0, backward=False):
    """x: [B, S, H, hd] view (hd contiguous, uniform token stride) ; pos: [B, S] int64 -> new contiguous tensor"""
    B, S, H, hd = x.shape
    assert x.stride(3) == 1 and x.stride(2) == hd and x.stride(0) == S * x.stride(1), "rope needs a [B,S,H,hd] view of a row-major buffer"
    pos = pos.contiguous().to(torch.int64)
    y = torch.empty((B, S, H, hd), dtype=x.dtype, device=x.device)
    _call("mb200_rope", _p(x), _p(y), _p(pos), _p(inv_freq), B * S, H, hd, x.stride(1), H * hd, float(attn_scaling),
          int(backward), _dt(x), _st())
    return y


def rope_table(pos, inv_freq, hd, attn_scaling, dtype):
    """[B*S, hd/2] float2 {cos, sin} table (rounded to `dtype` like the reference's cos/sin tensors)"""
    pos = pos.contiguous().to(torch.int64)
    n = pos.numel()
    tab = torch.empty((n, hd // 2, 2), dtype=torch.float32, device=pos.device)
    _call("mb200_rope_table", _p(pos), _p(inv_freq), _p(tab), n, hd, float(attn_scaling), BF16 if dtype == torch.bfloat16 else F32, _st())
    return tab


def _rope2(q, k, tab, backward):
    B, S, Hq, hd = q.shape
    Hk = k.shape[2]
    qo = torch.empty((B, S, Hq, hd), dtype=q.dtype, device=q.device)
    ko = torch.empty((B, S, Hk, hd), dtype=k.dtype, device=k.device)
    _call("mb200_rope2_bf16", _p(q), _p(k), _p(qo), _p(ko), _p(tab), B * S, Hq, Hk, hd, q.stride(1), k.stride(1),
          int(backward), _st())
    return qo, ko


def _rope2_ok(q, k):
    def ok(x):
        B, S, H, hd = x.shape
        return (x.dtype == torch.bfloat16 and hd % 16 == 0 and x.stride(3) == 1 and x.stride(2) == hd
                and x.stride(0) == S * x.stride(1) and x.stride(1) % 8 == 0 and x.data_ptr() % 16 == 0)
    return ok(q) and ok(k) and not FORCE_GENERIC


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, pos, inv_freq, attn_scaling, tab):
        ctx.fast = tab is not None and _rope2_ok(q, k)
        if ctx.fast:
            qo, ko = _rope2(q, k, tab, False)
            ctx.save_for_backward(tab)
        else:
            qo = rope_apply(q, pos, inv_freq, attn_scaling)
            ko = rope_apply(k, pos, inv_freq, attn_scaling)
            ctx.save_for_backward(pos, inv_freq)
        ctx.scaling = attn_scaling
        return qo, ko

    @staticmethod
    def backward(ctx, dq, dk):
        if ctx.fast:
            (tab,) = ctx.saved_tensors
            gq, gk = _rope2(dq.contiguous(), dk.contiguous(), tab, True)
            return gq, gk, None, None, None, None
        pos, inv_freq = ctx.saved_tensors
        return (rope_apply(dq.contiguous(), pos, inv_freq, ctx.scaling, backward=True),
                rope_apply(dk.contiguous(), pos, inv_freq, ctx.scaling, backward=True), None, None, None, None)


def rope(q, k, pos, inv_freq, attn_scaling=1.0, tab=None):
    """q: [B,S,H,hd], k: [B,S,Hkv,hd] -> rotated copies (rotate_half convention).  `tab` = rope_table(...) of this forward
    pass enables the one-launch vectorised path."""
    return _RopeFn.apply(q, k, pos, inv_freq, attn_scaling, tab)


# ---------------------------------------------------------------------------------------------- attention
def _strides12(q, k, v, o):
    arr = (ctypes.c_longlong * 12)(q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
                                   v.stride(0), v.stride(1), v.stride(2), o.stride(0), o.stride(1), o.stride(2))
    return arr


FAST_ATTN_MIN_SQ = int(os.environ.get("MB200_FAST_ATTN_MIN_SQ", "64"))
ATTN_FWD2 = os.environ.get("MB200_ATTN_FWD2", "1") == "1"      # two query tiles per CTA (ping-pong) forward kernel


def _attn_fast_ok(q, k, v, hd, Sq, Sk):
    if FORCE_GENERIC or q.dtype != torch.bfloat16 or hd != 128 or Sq < FAST_ATTN_MIN_SQ:
        return False
    for t in (q, k, v):
        if t.data_ptr() % 16 or any(s % 8 for s in t.stride()[:3]):
            return False
    return True


def _window_code(causal, window, Sk):
    """the kernels' `causal` argument: 0 none, 1 causal, W > 1 causal with a sliding window of W keys (only when it bites)"""
    if window is not None and causal and Sk > int(window) > 1:
        return int(window)
    return int(bool(causal))


def attention_fwd(q, k, v, causal, kmask, scale, return_kbits=False, out=None, window=None):
    """q [B,Sq,H,hd], k/v [B,Sk,Hkv,hd] (hd contiguous). Returns (o [B,Sq,H,hd] contiguous, lse [B,H,Sq] fp32).
    `out`: optional contiguous destination for o (a slice of a packed output, see attention_varlen).
    `window`: Mistral sliding window (key j visible to query i iff i - j < window); contexts longer than the window run on
    the generic kernel."""
    _need_cuda(q, k, v)
    B, Sq, H, hd = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    assert q.stride(3) == 1 and k.stride(3) == 1 and v.stride(3) == 1
    o = out if out is not None else torch.empty((B, Sq, H, hd), dtype=q.dtype, device=q.device)
    assert o.is_contiguous() and o.shape == (B, Sq, H, hd)
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    if kmask is not None:
        kmask = kmask.contiguous().to(torch.int64)
        assert kmask.shape == (B, Sk)
    st = _strides12(q, k, v, o)
    code = _window_code(causal, window, Sk)
    if code <= 1 and _attn_fast_ok(q, k, v, hd, Sq, Sk):
        kbits = None
        if ATTN_FWD2 and Sq > 128:
            if kmask is not None:
                kbits = kmask_bits(kmask).reshape(-1)
            _call("mb200_attn_fwd2_bf16", _p(q), _p(k), _p(v), _p(o), _p(lse), B, H, Hkv, Sq, Sk, hd, st, float(scale),
                  int(causal), _p(kbits), (Sk + 31) // 32 if kmask is not None else 0, _st())
            if return_kbits:
                return o, lse, kbits, True
            return o, lse
        if kmask is not None:
            kbits = torch.empty((B * ((Sk + 31) // 32),), dtype=torch.int32, device=q.device)
        _call("mb200_attn_fwd_bf16", _p(q), _p(k), _p(v), _p(o), _p(lse), B, H, Hkv, Sq, Sk, hd, st, float(scale),
              int(causal), _p(kmask), Sk if kmask is not None else 0, _p(kbits), _st())
        if return_kbits:
            return o, lse, kbits, True
        return o, lse
    _call("mb200_attn_generic_fwd", _p(q), _p(k), _p(v), _p(o), _p(lse), B, H, Hkv, Sq, Sk, hd, st, float(scale),
          code, _p(kmask), Sk if kmask is not None else 0, _dt(q), _st())
    if return_kbits:
        return o, lse, None, False
    return o, lse


def kmask_bits(kmask):
    """[B, Sk] int mask -> [B, ceil(Sk/32)] uint32 bitmask (as int32 tensor)"""
    km = kmask.contiguous().to(torch.int64)
    B, Sk = km.shape
    bits = torch.empty((B, (Sk + 31) // 32), dtype=torch.int32, device=km.device)
    _call("mb200_kmask_bits", _p(km), Sk, _p(bits), B, Sk, _st())
    return bits


def decode_attention(q, k_cache, v_cache, ctx, kmask, scale, kbits=None):
    """q [B,1,H,128] bf16; k_cache/v_cache [B,cap,Hkv,128] (first ctx tokens valid); kmask [B,ctx] or None
    (or its precomputed bitmask `kbits`, shared by all layers of a decode step)."""
    B, _, H, hd = q.shape
    Hkv = k_cache.shape[2]
    o = torch.empty((B, 1, H, hd), dtype=q.dtype, device=q.device)
    splits = _L().mb200_decode_attn_splits(ctx)
    part = torch.empty((B * H * splits * (hd + 2),), dtype=torch.float32, device=q.device)
    words = 0
    if kbits is None and kmask is not None:
        kbits = kmask_bits(kmask[:, :ctx])
    if kbits is not None:
        words = kbits.shape[1]
    _call("mb200_decode_attn_bf16", _p(q), _p(k_cache), _p(v_cache), _p(o), _p(part), B, H, Hkv, ctx, hd,
          q.stride(0), q.stride(2), k_cache.stride(0), k_cache.stride(1), k_cache.stride(2), o.stride(0), o.stride(2),
          float(scale), _p(kbits), words, _st())
    return o


def decode_attention_paged(q, cache, layer_idx, ctx, kmask, scale, kbits=None):
    """decode_attention over the paged cache (models/kv_cache.py): q [B,1,H,128] bf16 against the first `ctx` cached
    tokens of layer `layer_idx`; the kernel walks the block table itself."""
    B, _, H, hd = q.shape
    o = torch.empty((B, 1, H, hd), dtype=q.dtype, device=q.device)
    splits = _L().mb200_decode_attn_splits(ctx)
    part = torch.empty((B * H * splits * (hd + 2),), dtype=torch.float32, device=q.device)
    words = 0
    if kbits is None and kmask is not None:
        kbits = kmask_bits(kmask[:, :ctx])
    if kbits is not None:
        words = kbits.shape[1]
    tab = cache.device_table()
    _call("mb200_decode_attn_paged_bf16", _p(q), _p(tab), tab.shape[1], layer_idx * cache.layer_stride, cache.v_off,
          _p(o), _p(part), B, H, cache.Hkv, ctx, hd, q.stride(0), q.stride(2), o.stride(0), o.stride(2), float(scale),
          _p(kbits), words, _st())
    return o


ATTN_BWD_SINGLE_PASS = os.environ.get("MB200_ATTN_BWD_SINGLE_PASS", "1") == "1"
ATTN_BWD_WS_MAX = int(os.environ.get("MB200_ATTN_BWD_WS_MAX_GB", "12")) << 30
_attn_bwd_ws_cache = {}


def _attn_bwd_ws(B, H, Sq, device):
    """dS^T scratch of the single-pass attention backward ([B*H][Sq_pad][Sq_pad] bf16: 4 GB for one config-2 sample), kept and
    re-used across calls (grow-only, per device and stream); None when it would exceed MB200_ATTN_BWD_WS_MAX_GB."""
    need = int(_L().mb200_attn_bwd_ds_bytes(B, H, Sq))
    if need > ATTN_BWD_WS_MAX:
        return None
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _attn_bwd_ws_cache.get(key)
    if ws is None or ws.numel() < need:
        _attn_bwd_ws_cache[key] = None
        ws = torch.empty((need,), dtype=torch.uint8, device=device)
        _attn_bwd_ws_cache[key] = ws
    return ws


def attention_bwd(q, k, v, o, do, lse, causal, kmask, scale, kbits=None, fast=False, out=None, window=None):
    """`out`: optional (dq, dk, dv) contiguous destinations (slices of packed gradients, see attention_varlen)."""
    B, Sq, H, hd = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    do = do.contiguous()
    assert o.is_contiguous()
    if fast:
        if out is not None:
            dq, dk, dv = out
            assert dq.is_contiguous() and dk.is_contiguous() and dv.is_contiguous()
        else:
            dq = torch.empty((B, Sq, H, hd), dtype=q.dtype, device=q.device)
            dk = torch.empty((B, Sk, Hkv, hd), dtype=q.dtype, device=q.device)
            dv = torch.empty((B, Sk, Hkv, hd), dtype=q.dtype, device=q.device)
        delta = torch.empty((2 * B * H * _L().mb200_attn_bwd_sq_pad(Sq),), dtype=torch.float32, device=q.device)
        st = _strides12(q, k, v, o)
        ws = _attn_bwd_ws(B, H, Sq, q.device) if (ATTN_BWD_SINGLE_PASS and Sq == Sk) else None
        _call("mb200_attn_bwd_bf16_sp", _p(q), _p(k), _p(v), _p(o), _p(do), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv),
              B, H, Hkv, Sq, Sk, hd, st, float(scale), int(causal), _p(kmask), Sk if kmask is not None else 0,
              _p(kbits), _p(ws), _st())
        return dq, dk, dv
    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    dq = torch.empty_like(qc); dk = torch.empty_like(kc); dv = torch.empty_like(vc)
    delta = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    st = _strides12(qc, kc, vc, o)
    _call("mb200_attn_generic_bwd", _p(qc), _p(kc), _p(vc), _p(o), _p(do), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv),
          B, H, Hkv, Sq, Sk, hd, st, float(scale), _window_code(causal, window, Sk), _p(kmask),
          Sk if kmask is not None else 0, _dt(q), _st())
    if out is not None:
        out[0].copy_(dq); out[1].copy_(dk); out[2].copy_(dv)
        return out
    return dq, dk, dv


ATTN_PAD_HEAD_DIM = os.environ.get("MB200_ATTN_PAD_HD", "1") == "1"
attn_padded_calls = 0       # attention calls that took the zero-padded tensor-core route (tests read it)


def _pad_hd_ok(q, k, v, Sq):
    """head_dim 72 (SigLIP), 96 (Idefics2 perceiver), ... fit no tcgen05 tile; zero-padding the head dimension of q, k, v to 128
    changes neither q.k nor p.v, so the hd-128 tensor-core kernels can serve them (the padded lanes cost 25-44 % extra MMA work,
    still far ahead of the SIMT kernel).  The frozen ViT does this once at the WEIGHT level (models/vision.py); this is the
    activation-level version for everything that trains."""
    hd = q.shape[-1]
    return (ATTN_PAD_HEAD_DIM and not FORCE_GENERIC and q.is_cuda and q.dtype == torch.bfloat16 and 32 <= hd < 128 and hd % 8 == 0
            and Sq >= FAST_ATTN_MIN_SQ and k.shape[-1] == hd and v.shape[-1] == hd)


def _pad_hd(x):
    return torch.nn.functional.pad(x, (0, 128 - x.shape[-1]))


class _AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, kmask, scale, window):
        global attn_padded_calls
        if kmask is not None:
            kmask = kmask.contiguous().to(torch.int64)
        hd = q.shape[-1]
        pad = _pad_hd_ok(q, k, v, q.shape[1]) and _window_code(causal, window, k.shape[1]) <= 1
        if pad:
            q, k, v = _pad_hd(q), _pad_hd(k), _pad_hd(v)
            attn_padded_calls += 1
        o, lse, kbits, fast = attention_fwd(q, k, v, causal, kmask, scale, return_kbits=True, window=window)
        ctx.save_for_backward(q, k, v, o, lse, kmask, kbits)
        ctx.causal, ctx.scale, ctx.fast, ctx.window = causal, scale, fast, window
        ctx.hd = hd if pad else None
        return o[..., :hd].contiguous() if pad else o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, kmask, kbits = ctx.saved_tensors
        if ctx.hd is not None:
            do = _pad_hd(do)
        dq, dk, dv = attention_bwd(q, k, v, o, do, lse, ctx.causal, kmask, ctx.scale, kbits=kbits, fast=ctx.fast,
                                   window=ctx.window)
        if ctx.hd is not None:
            dq, dk, dv = dq[..., :ctx.hd], dk[..., :ctx.hd], dv[..., :ctx.hd]
        return dq, dk, dv, None, None, None, None


def attention(q, k, v, causal=False, kmask=None, scale=None, window=None):
    """softmax(q k^T * scale + mask) v with GQA.  q [B,Sq,H,hd]; k,v [B,Sk,Hkv,hd]; kmask [B,Sk] (non-zero = attend);
    window: causal sliding window (Mistral), None = unlimited"""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    return _AttentionFn.apply(q, k, v, bool(causal), kmask, float(scale), window)


class _VarlenAttentionFn(torch.autograd.Function):
    """Causal attention over PACKED sequences (sequence packing, ref: mantis/train/data.py:1546-1671 builds a block-diagonal
    4-D mask + per-sample position ids for this): every (batch row, start, end) segment attends only to itself.  One kernel
    launch per segment on strided views of the packed q/k/v -- outputs and gradients are written straight into slices of
    ONE packed tensor, so packing adds no copies and no masked-out FLOPs (a dense block-diagonal mask would compute
    S^2 scores and throw most of them away)."""

    @staticmethod
    def forward(ctx, q, k, v, segs, kmask, scale):
        B, S, H, hd = q.shape
        o = torch.zeros((B, S, H, hd), dtype=q.dtype, device=q.device)       # positions outside every segment stay 0
        if kmask is not None:
            kmask = kmask.contiguous().to(torch.int64)
        per_seg = []
        for (b, s0, s1) in segs:
            km = kmask[b:b + 1, s0:s1] if kmask is not None else None
            _, lse, kbits, fast = attention_fwd(q[b:b + 1, s0:s1], k[b:b + 1, s0:s1], v[b:b + 1, s0:s1], True, km, scale,
                                                return_kbits=True, out=o[b:b + 1, s0:s1])
            per_seg.append((lse, kbits, fast))
        ctx.save_for_backward(q, k, v, o, kmask)
        ctx.segs, ctx.per_seg, ctx.scale = segs, per_seg, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, kmask = ctx.saved_tensors
        do = do.contiguous()
        dq = torch.zeros_like(q, memory_format=torch.contiguous_format)
        dk = torch.zeros_like(k, memory_format=torch.contiguous_format)
        dv = torch.zeros_like(v, memory_format=torch.contiguous_format)
        for (b, s0, s1), (lse, kbits, fast) in zip(ctx.segs, ctx.per_seg):
            km = kmask[b:b + 1, s0:s1] if kmask is not None else None
            attention_bwd(q[b:b + 1, s0:s1], k[b:b + 1, s0:s1], v[b:b + 1, s0:s1], o[b:b + 1, s0:s1], do[b:b + 1, s0:s1], lse,
                          True, km, ctx.scale, kbits=kbits, fast=fast,
                          out=(dq[b:b + 1, s0:s1], dk[b:b + 1, s0:s1], dv[b:b + 1, s0:s1]))
        return dq, dk, dv, None, None, None


def attention_varlen(q, k, v, segs, kmask=None, scale=None):
    """causal self-attention over packed sequences.  q [B,S,H,hd]; k,v [B,S,Hkv,hd]; segs: list of (batch row, start, end)
    with non-overlapping [start, end) ranges; kmask [B,S] (non-zero = attend) or None."""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    return _VarlenAttentionFn.apply(q, k, v, tuple(tuple(int(x) for x in sgm) for sgm in segs), kmask, float(scale))


def packed_segments(position_ids):
    """[B, S] per-sample position ids of a packed batch (each packed sample restarts at 0) -> list of (b, start, end).
    One host sync (the segment table drives the launch loop)."""
    pos = position_ids if position_ids.dim() == 2 else position_ids.unsqueeze(0)
    B, S = pos.shape
    starts = (pos == 0)
    starts[:, 0] = True
    idx = starts.nonzero().tolist()
    segs = []
    for i, (b, s0) in enumerate(idx):
        s1 = idx[i + 1][1] if i + 1 < len(idx) and idx[i + 1][0] == b else S
        segs.append((b, s0, s1))
    return segs


# ---------------------------------------------------------------------------------------------- merge (scatter)
_merge_ws_cache = {}


def _merge_ws(B, T, device):
    need = _L().mb200_merge_ws_bytes(B, T)
    key = (device, )
    ws = _merge_ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros((max(need, 1 << 16),), dtype=torch.uint8, device=device)
        _merge_ws_cache[key] = ws
    return ws


_deferred_checks = []       # (pinned host flag, event that marks its arrival, message)
_flag_pool = None           # pinned ring of host flags (one cudaHostAlloc, not one per check)
_flag_next = 0
_FLAG_SLOTS = 512


def _defer_check(flag_dev, msg):
    """queue a device-side boolean for a later host-side check WITHOUT synchronising: the flag travels to pinned host memory
    on the current stream and is looked at once its event has completed"""
    global _flag_pool, _flag_next
    if _flag_pool is None:
        _flag_pool = torch.empty((_FLAG_SLOTS,), dtype=torch.bool).pin_memory()
    if len(_deferred_checks) >= _FLAG_SLOTS // 2:
        check_deferred(block=False)                       # inference loops never call check_deferred: keep the queue short
        if len(_deferred_checks) >= _FLAG_SLOTS - 1:
            check_deferred(block=True)
    host = _flag_pool[_flag_next:_flag_next + 1]
    _flag_next = (_flag_next + 1) % _FLAG_SLOTS
    host.copy_(flag_dev.reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _deferred_checks.append((host, ev, msg))


def check_deferred(block=True):
    """Raise the errors of sync-free merges (see merge_input_ids_with_image_features(plan_hint=...)).  block=False (what
    B200Trainer.optimizer_step uses, every step) only looks at flags that have already arrived, so the host never waits for
    the device; block=True (call it wherever the host synchronises anyway, e.g. when the loss is read) drains them all."""
    keep = []
    err = None
    for host, ev, msg in _deferred_checks:
        if block:
            ev.synchronize()
        elif not ev.query():
            keep.append((host, ev, msg))
            continue
        if err is None and not bool(host[0]):
            err = msg
    _deferred_checks[:] = keep
    if err is not None:
        raise ValueError(err)


def merge_plan(input_ids, inputs_embeds, P, image_token, pad_token, sync=True):
    """Runs the plan kernel and reads the 8-word header back (the ONE host sync of the merge: the output
    length is data dependent).  Returns (ws, header list); with sync=False the header stays on the device."""
    B, T = input_ids.shape
    D = inputs_embeds.shape[-1]
    ws = _merge_ws(B, T, input_ids.device)
    header = torch.empty((8,), dtype=torch.int64, device=input_ids.device)
    _call("mb200_merge_plan", _p(input_ids), _p(inputs_embeds), _dt(inputs_embeds), B, T, D, P, int(image_token),
          int(pad_token), _p(ws), _p(header), _st())
    return ws, (header.tolist() if sync else header)


class _MergeRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs_embeds, image_features, srcmap, S):
        B, T, D = inputs_embeds.shape
        img2 = image_features.reshape(-1, D).contiguous()
        text = inputs_embeds.contiguous()
        out = torch.empty((B, S, D), dtype=text.dtype, device=text.device)
        row_bytes = D * text.element_size()
        _call("mb200_merge_rows", _p(srcmap), _p(text), _p(img2), _p(out), B, S, T, row_bytes, img2.shape[0], _st())
        ctx.save_for_backward(srcmap)
        ctx.dims = (B, T, D, S, tuple(image_features.shape))
        return out

    @staticmethod
    def backward(ctx, gout):
        (srcmap,) = ctx.saved_tensors
        B, T, D, S, img_shape = ctx.dims
        gout = gout.contiguous()
        gtext = gimg = None
        if ctx.needs_input_grad[0]:
            gtext = torch.zeros((B, T, D), dtype=gout.dtype, device=gout.device)
        if ctx.needs_input_grad[1]:
            gimg = torch.empty(img_shape, dtype=gout.dtype, device=gout.device)
        n_img_rows = 1
        for s in img_shape[:-1]:
            n_img_rows *= s
        _call("mb200_merge_rows_bwd", _p(srcmap), _p(gout), _p(gtext), _p(gimg), B, S, T,
              D * gout.element_size(), n_img_rows, _st())
        return gtext, gimg, None, None


def merge_input_ids_with_image_features(image_features, inputs_embeds, input_ids, attention_mask, labels,
                                        image_token_index, pad_token_id, ignore_index=-100, plan_hint=None,
                                        zero_pad_rows=False):
    """CUDA re-implementation of LlavaForConditionalGeneration._merge_input_ids_with_image_features
    (mantis/models/mllava/modeling_llava.py:293-360).  Same return tuple, same ValueError.

    plan_hint = {"max_image_tokens": n, "left_padding": bool} (emitted by train.data.Collator from the host copy of
    input_ids) removes the merge's host sync: the merged length n * (P - 1) + T and the padding side are then known
    up front, the plan kernel's own result is compared with them ON THE DEVICE and a mismatch (wrong hint, or the
    reference's image-count ValueError) is raised by ops.check_deferred() at the caller's next sync point.

    zero_pad_rows: the LLaVA-NeXT variant additionally zeroes the merged rows that came from pad tokens
    (mantis/models/mllava_next/modeling_llava_next.py:455-461) -- done by pointing their source-map entries at "zero fill"."""
    _need_cuda(image_features, inputs_embeds, input_ids)
    num_images, P, D = image_features.shape
    B, T = input_ids.shape
    if (D * inputs_embeds.element_size()) % 16 != 0:
        raise ValueError("embedding rows must be multiples of 16 bytes")
    ids = input_ids.contiguous().to(torch.int64)
    emb = inputs_embeds.contiguous()
    mask_dtype = attention_mask.dtype
    am = attention_mask.contiguous().to(torch.int64)
    lab = labels.contiguous().to(torch.int64) if labels is not None else None
    if plan_hint is not None:
        ws, hdr = merge_plan(ids, emb, P, image_token_index, pad_token_id, sync=False)
        S = int(plan_hint["max_image_tokens"]) * (P - 1) + T
        left_padding = int(bool(plan_hint["left_padding"]))
        expect = torch.tensor([S, left_padding, num_images * P], dtype=torch.int64, device=ids.device)
        _defer_check((hdr[:3] == expect).all(),
                     "The input provided to the model are wrong (sync-free merge): the plan_hint or the number of"
                     f" images ({num_images}) does not match the image tokens in input_ids. This prevents correct"
                     " indexing and breaks batch generation.")
    else:
        ws, hdr = merge_plan(ids, emb, P, image_token_index, pad_token_id)
        S, left_padding, n_slots = int(hdr[0]), int(hdr[1]), int(hdr[2])
        if n_slots != num_images * P:
            raise ValueError(
                f"The input provided to the model are wrong. The number of image tokens is {int(hdr[4])} while"
                f" the number of image given to the model is {num_images}. This prevents correct indexing and breaks batch generation.")
    dev = ids.device
    srcmap = torch.empty((B, S), dtype=torch.int32, device=dev)
    out_mask = torch.empty((B, S), dtype=torch.int64, device=dev)
    out_pos = torch.empty((B, S), dtype=torch.int64, device=dev)
    out_labels = torch.empty((B, S), dtype=torch.int64, device=dev) if lab is not None else None
    _call("mb200_merge_index", _p(ids), _p(am), _p(lab), _p(ws), B, T, P, S, left_padding, int(image_token_index),
          int(ignore_index), _p(srcmap), _p(out_mask), _p(out_labels), _p(out_pos), _st())
    if zero_pad_rows:
        src_tok = torch.gather(ids, 1, srcmap.clamp(min=0).to(torch.int64))
        srcmap = torch.where((srcmap >= 0) & (src_tok == int(pad_token_id)), torch.full_like(srcmap, -1), srcmap)
    final = _MergeRowsFn.apply(emb, image_features, srcmap, S)
    if mask_dtype != torch.int64:
        out_mask = out_mask.to(mask_dtype)
    return final, out_mask, out_labels, out_pos


# ---------------------------------------------------------------------------------------------- loss
def shift_labels(labels, attention_mask, ignore_index=-100):
    """eff[b,s] = labels[b,s+1] if mask[b,s+1] != 0 and != ignore_index else -100 ; count of valid rows (fp32 scalar)"""
    B, S = labels.shape
    labels = labels.contiguous().to(torch.int64)
    am = attention_mask.contiguous().to(torch.int64) if attention_mask is not None else None
    out = torch.empty_like(labels)
    count = torch.zeros((1,), dtype=torch.float32, device=labels.device)
    _call("mb200_shift_labels", _p(labels), _p(am), _p(out), B, S, int(ignore_index), _p(count), _st())
    return out, count


LM_HEAD_CHUNK = int(os.environ.get("MB200_LM_HEAD_CHUNK", "4096"))


LM_HEAD_SKIP_IGNORED = os.environ.get("MB200_LMHEAD_SKIP_IGNORED", "1") == "1"


def gather_rows(x2, idx):
    """out[i, :] = x2[idx[i], :]  (row gather on the embedding kernel)"""
    n, D = idx.numel(), x2.shape[1]
    out = torch.empty((n, D), dtype=x2.dtype, device=x2.device)
    if n:
        _call("mb200_embedding_fwd", _p(idx), _p(x2), _p(out), n, D, x2.shape[0], _dt(x2), _st())
    return out


class _LMHeadCEFn(torch.autograd.Function):
    """Fused LM head + shifted masked cross-entropy: logits are produced chunk by chunk, consumed by the CE kernel
    (which overwrites them with dlogits), and immediately folded into d(hidden) and d(W) -- the [B,S,V] logits
    tensor (8 GB at Mantis-8B scale) is never materialised.  loss = sum_rows(lse - logit[target]) / n_valid.

    Rows whose effective label is ignored (image slots, prompt tokens, padding: the reference gathers them away *after*
    computing their logits, modeling_llava.py:526-531) contribute neither to the loss nor to any gradient, so they are
    compacted away *before* the LM-head GEMMs (LM_HEAD_SKIP_IGNORED): identical loss / gradients, fewer FLOPs."""

    @staticmethod
    def forward(ctx, hidden, weight, eff_labels, count, valid_rows_hint=None):
        h2 = hidden.reshape(-1, hidden.shape[-1])
        if h2.stride(1) != 1:
            h2 = h2.contiguous()
        n_all, D = h2.shape
        V = weight.shape[0]
        lab = eff_labels.reshape(-1)
        dev = h2.device
        need_h, need_w = hidden.requires_grad, weight.requires_grad
        # torch's cross_entropy raises on a target >= the number of classes; the kernels cannot raise, so the check is queued for
        # the caller's next synchronisation point (ops.check_deferred) instead of silently treating such rows as ignored
        _defer_check((lab < V).all(), f"cross-entropy target out of range: a label is >= the vocabulary size ({V})")
        idx = None
        if LM_HEAD_SKIP_IGNORED:
            if valid_rows_hint is not None:
                # sync-free: the collator counted the supervised rows on the host (train.data.Collator `valid_rows`); the device
                # count is compared with it at the caller's next sync point (ops.check_deferred)
                n_hint = int(valid_rows_hint)
                idx = torch.empty((n_hint,), dtype=torch.int64, device=dev)
                cnt = torch.empty((1,), dtype=torch.int32, device=dev)
                _call("mb200_compact_valid_rows", _p(lab.contiguous()), n_all, _p(idx), n_hint, _p(cnt), _st())
                _defer_check(cnt[0] == n_hint, f"loss: the batch's `valid_rows` hint ({n_hint}) does not match the number of "
                                               "supervised positions after the image-token merge")
            else:
                idx = (lab >= 0).nonzero(as_tuple=False).squeeze(1)       # one host sync (row count)
            if idx.numel() == n_all:
                idx = None
        if idx is not None:
            hsrc = gather_rows(h2, idx)
            lsrc = lab.index_select(0, idx)
        else:
            hsrc, lsrc = h2, lab
        n = hsrc.shape[0]
        ld = (V + 7) // 8 * 8
        acc = torch.zeros((2,), dtype=torch.float32, device=dev)
        inv = (1.0 / count).to(torch.float32)                # device scalar: dlogits scale (mean over valid rows)
        dh = torch.empty_like(hsrc) if need_h else None
        dw = None
        chunk = max(1, min(LM_HEAD_CHUNK, n))
        buf = torch.empty((chunk, ld), dtype=h2.dtype, device=dev)
        loss_rows = torch.empty((chunk,), dtype=torch.float32, device=dev)
        for r0 in range(0, n, chunk):
            r1 = min(n, r0 + chunk)
            m = r1 - r0
            logits = buf[:m, :V]
            gemm(hsrc[r0:r1], weight, out=logits)
            _call("mb200_ce_fwd_bwd", _p(logits), _p(lsrc[r0:r1]), _p(loss_rows), None,
                  _p(logits) if (need_h or need_w) else None, m, V, ld, _p(inv), 1.0, _dt(h2), _st())
            _call("mb200_ce_reduce", _p(loss_rows), _p(lsrc[r0:r1]), m, V, _p(acc), 1, _st())
            if need_h:
                gemm(logits, weight, trans_a=False, trans_b=False, out=dh[r0:r1])
            if need_w:
                if dw is None:
                    dw = gemm(logits, hsrc[r0:r1], trans_a=True, trans_b=False)
                else:
                    gemm(logits, hsrc[r0:r1], trans_a=True, trans_b=False, addend=dw, out=dw)
        if need_w and dw is None:
            dw = torch.zeros_like(weight)
        if need_h and idx is not None:                       # scatter the compacted rows back (others get zero gradient)
            full = torch.zeros((n_all, D), dtype=dh.dtype, device=dev)
            if n:
                _call("mb200_embedding_bwd", _p(idx), _p(dh), _p(full), n, D, n_all, _dt(dh), _st())
            dh = full
        loss = acc[0] / acc[1]
        ctx.save_for_backward(dh, dw)
        ctx.shape = hidden.shape
        return loss

    @staticmethod
    def backward(ctx, gloss):
        dh, dw = ctx.saved_tensors
        gh = gw = None
        if dh is not None:
            gh = (dh * gloss.to(dh.dtype)).reshape(ctx.shape)
        if dw is not None:
            gw = dw * gloss.to(dw.dtype)
        return gh, gw, None, None, None


def lm_head_ce(hidden, weight, eff_labels, count, valid_rows_hint=None):
    return _LMHeadCEFn.apply(hidden, weight, eff_labels, count, valid_rows_hint)


class _CEFn(torch.autograd.Function):
    """cross-entropy over materialised logits [n, V] with labels (-100 = ignore), mean over valid rows"""

    @staticmethod
    def forward(ctx, logits2, lab, count):
        n, V = logits2.shape
        assert logits2.stride(1) == 1
        dev = logits2.device
        _defer_check((lab < V).all(), f"cross-entropy target out of range: a label is >= the number of classes ({V})")
        loss_rows = torch.empty((n,), dtype=torch.float32, device=dev)
        acc = torch.zeros((2,), dtype=torch.float32, device=dev)
        dl = torch.empty((n, V), dtype=logits2.dtype, device=dev) if logits2.requires_grad else None
        inv = (1.0 / count).to(torch.float32)
        _call("mb200_ce_fwd_bwd", _p(logits2), _p(lab), _p(loss_rows), None, _p(dl), n, V, logits2.stride(0), _p(inv), 1.0,
              _dt(logits2), _st())
        _call("mb200_ce_reduce", _p(loss_rows), _p(lab), n, V, _p(acc), 0, _st())
        ctx.save_for_backward(dl)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return (dl * g.to(dl.dtype)) if dl is not None else None, None, None


def cross_entropy(logits2, labels1, count):
    return _CEFn.apply(logits2, labels1.contiguous(), count)


# ---------------------------------------------------------------------------------------------- misc
def im2col(pixels, patch, k_pad, out_dtype):
    N, C, H, W = pixels.shape
    px = pixels.contiguous()
    gh, gw = H // patch, W // patch
    out = torch.empty((N * gh * gw, k_pad), dtype=out_dtype, device=px.device)
    _call("mb200_im2col", _p(px), _dt(px), _p(out), BF16 if out_dtype == torch.bfloat16 else F32, N, C, H, W, patch,
          k_pad, _st())
    return out


def resize_u8_pass(img, bounds, coef, out_len, horizontal):
    """one pass of the antialiased 8-bit resampler: img [H, W, C] uint8 -> [H, out_len, C] (horizontal) / [out_len, W, C];
    bounds int32 [out_len, 2], coef int32 [out_len, ksize] on the device"""
    _need_cuda(img, bounds, coef)
    assert img.dtype == torch.uint8 and img.dim() == 3 and bounds.dtype == torch.int32 and coef.dtype == torch.int32
    img = img.contiguous()
    H, W, C = img.shape
    out = torch.empty((H, out_len, C) if horizontal else (out_len, W, C), dtype=torch.uint8, device=img.device)
    _call("mb200_resize_u8_pass", _p(img), _p(out), _p(bounds.contiguous()), _p(coef.contiguous()), coef.shape[1], H, W,
          int(out_len), C, int(bool(horizontal)), _st())
    return out


def image_normalize_u8(pixels_u8, lut, channels_last, out_dtype):
    """uint8 [N,H,W,C] (channels_last) or [N,C,H,W] on the device + fp32 lut [C,256] -> normalized [N,C,H,W] out_dtype"""
    _need_cuda(pixels_u8, lut)
    assert pixels_u8.dtype == torch.uint8 and lut.dtype == torch.float32 and lut.shape[1] == 256
    px = pixels_u8.contiguous()
    if channels_last:
        N, H, W, C = px.shape
    else:
        N, C, H, W = px.shape
    assert lut.shape[0] == C
    out = torch.empty((N, C, H, W), dtype=out_dtype, device=px.device)
    _call("mb200_image_normalize_u8", _p(px), _p(lut.contiguous()), _p(out), BF16 if out_dtype == torch.bfloat16 else F32,
          N, C, H, W, int(bool(channels_last)), _st())
    return out


class _AddRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, table, idx, period):
        n, D = x2.shape
        x2 = x2.contiguous(); tab = table.contiguous()
        y = torch.empty_like(x2)
        if idx is not None:
            idx = idx.contiguous().to(torch.int64)
        _call("mb200_add_rows", _p(x2), _p(tab), _p(idx), _p(y), n, D, int(period or 0), _dt(x2), _st())
        ctx.save_for_backward(idx)
        ctx.period, ctx.tshape = period, tuple(table.shape)
        return y

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        gt = None
        if ctx.needs_input_grad[1]:
            g2 = g.contiguous()
            n, D = g2.shape
            ids = idx if idx is not None else (torch.arange(n, device=g.device, dtype=torch.int64) % ctx.period)
            gt = torch.zeros(ctx.tshape, dtype=g.dtype, device=g.device)
            _call("mb200_embedding_bwd", _p(ids), _p(g2), _p(gt), n, D, ctx.tshape[0], _dt(g2), _st())
        return g, gt, None, None


def add_rows(x2, table, idx=None, period=None):
    """y[r,:] = x[r,:] + table[idx[r] or r % period, :]   (position / type embeddings)"""
    return _AddRowsFn.apply(x2, table, idx, period)


def rows_all_zero(x2):
    """int32 flags[r] = 1 iff row r of the 2-D tensor is entirely zero"""
    x2 = x2.contiguous()
    flags = torch.empty((x2.shape[0],), dtype=torch.int32, device=x2.device)
    _call("mb200_rows_all_zero", _p(x2), x2.shape[0], x2.shape[1], _p(flags), _dt(x2), _st())
    return flags


def adamw_flat(p, lo, g, m, v, blk_group, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, norm_sq=None, max_norm=0.0,
               zero_grad=True):
    """ONE launch of AdamW over flat buffers.  p bf16 + lo (int16 low halves = fp32 master) or p fp32 (lo None); g fp32/bf16."""
    _need_cuda(p, g, m, v)
    if p.dtype == torch.bfloat16 and lo is None:
        raise ValueError("bf16 parameters need their fp32-master low halves (lo)")
    _call("mb200_adamw_flat", _p(p), _p(lo), _p(g), _p(m), _p(v), _p(blk_group), p.numel(), float(lr), float(beta1),
          float(beta2), float(eps), float(wd), int(step), float(grad_scale), _p(norm_sq), float(max_norm or 0.0),
          int(bool(zero_grad)), _dt(p), _dt(g), _st())


def master_split(master, hi, lo):
    """fp32 master -> (bf16 weight `hi`, int16 low half `lo`), flat tensors of equal length"""
    _call("mb200_master_split", _p(master), _p(hi), _p(lo), master.numel(), _st())


def master_join(hi, lo, out=None):
    """(bf16 weight, int16 low half or None) -> fp32 master"""
    if out is None:
        out = torch.empty(hi.shape, dtype=torch.float32, device=hi.device)
    _call("mb200_master_join", _p(hi), _p(lo), _p(out), hi.numel(), _st())
    return out


def accum_f32(dst, src, scale=1.0, accumulate=True):
    """dst (fp32, contiguous) = (dst if accumulate else 0) + scale * src (bf16 / fp32, contiguous, same numel)"""
    assert dst.dtype == torch.float32 and dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    _call("mb200_accum_f32", _p(dst), _p(src), dst.numel(), float(scale), int(bool(accumulate)), _dt(src), _st())


def sumsq(g, out):
    _call("mb200_sumsq", _p(g), g.numel(), _p(out), _dt(g), _st())
