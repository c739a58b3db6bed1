"""fp32-accurate tensor-core GEMMs for the training step's dense layers (csrc/gemm3_tc.cu: tcgen05 kind::tf32, 3xTF32
split, TMEM accumulators) and the `nn.Linear` replacement built on them.

    y = tc_linear([x_0, x_1, ...], weight, bias, relu)        y = relu?([x_0 | x_1 | ...] @ weight.T + bias)

is what the reference's MLP layers compute with cuBLAS fp32 SGEMMs (internal/models.py:L438-441 `density_layer`,
L643-652 `lin_second_stage_*` on `torch.cat([x, inputs])`, L650 `rgb_layer`); the segments replace the reference's
`torch.cat` copies (the K loop simply walks over several operand tensors).  Backward = the same kernels:
dx_s = dy W_s (NT form), dW_s = dy^T x_s (TN form, reduction over the rows), db = column sums.  There is no CPU path."""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib


def _chk(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _f32c(t, name):
    _chk(t.device.type == "cuda", f"{name} must be a CUDA tensor (no CPU path)")
    _chk(t.dtype == torch.float32, f"{name} must be float32")
    return t if (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]) or t.is_contiguous() else t.contiguous()


def gemm_nt(pairs: Sequence, bias: Optional[torch.Tensor] = None, relu: bool = False, out: Optional[torch.Tensor] = None):
    """C[M,N] = sum_s A_s[M,k_s] @ B_s[N,k_s].T (+ bias) (relu).  pairs: [(A_s, B_s), ...]; N <= 256."""
    lib = _lib.load()
    A0, B0 = pairs[0]
    M, N = A0.shape[0], B0.shape[0]
    segs = (_lib.GemmSeg * len(pairs))()
    keep = []
    for i, (a, b) in enumerate(pairs):
        a, b = _f32c(a, "A"), _f32c(b, "B")
        _chk(a.dim() == 2 and b.dim() == 2 and a.shape[0] == M and b.shape[0] == N and a.shape[1] == b.shape[1],
             f"gemm_nt: segment {i} has shapes {tuple(a.shape)} x {tuple(b.shape)}")
        keep += [a, b]
        segs[i].a, segs[i].b = a.data_ptr(), b.data_ptr()
        segs[i].lda, segs[i].ldb, segs[i].k = a.stride(0), b.stride(0), a.shape[1]
    if bias is not None:
        bias = _f32c(bias, "bias")
        _chk(bias.numel() == N, "gemm_nt: bias must have N elements")
    if out is None:
        out = torch.empty((M, N), device=A0.device, dtype=torch.float32)
    _chk(out.shape == (M, N) and out.stride(1) == 1 and out.dtype == torch.float32, "gemm_nt: bad output tensor")
    with torch.cuda.device(A0.device):
        rc = lib.ucnerf_gemm_nt(M, N, len(pairs), segs, None if bias is None else bias.data_ptr(), int(bool(relu)),
                                out.data_ptr(), out.stride(0), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gemm_nt")
    return out


def gemm_tn(A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None):
    """C[N1,N2] (+)= A[M,N1].T @ B[M,N2] (reduction over the M rows).  `out` (optional) is accumulated into and may be a
    column slice of a larger matrix (unit stride along N2)."""
    lib = _lib.load()
    A, B = _f32c(A, "A"), _f32c(B, "B")
    _chk(A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0], "gemm_tn: A [M,N1], B [M,N2]")
    M, N1, N2 = A.shape[0], A.shape[1], B.shape[1]
    if out is None:
        out = torch.zeros((N1, N2), device=A.device, dtype=torch.float32)
    _chk(out.shape == (N1, N2) and out.stride(1) == 1 and out.dtype == torch.float32, "gemm_tn: bad output tensor")
    with torch.cuda.device(A.device):
        rc = lib.ucnerf_gemm_tn(M, N1, N2, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(), out.stride(0),
                                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gemm_tn")
    return out


def status():
    """Raises if a GEMM kernel's pipeline watchdog fired since the last call (synchronises the device)."""
    buf = (C.c_uint32 * 32)()
    _lib.check(_lib.load().ucnerf_gemm_status(buf), "gemm_status")


def relu_mask_colsum(gy: torch.Tensor, y: Optional[torch.Tensor], want_g: bool = True, want_colsum: bool = True):
    """One pass: g = gy * (y > 0) (y None: g is gy itself) and the column sums of g.  Returns (g, colsum)."""
    M, N = gy.shape
    if N % 4 != 0 or N > 1024 or M == 0:          # tiny / odd widths (rgb layer, proposal density): torch
        g = gy if y is None else gy * (y > 0)
        return g, (g.sum(dim=0) if want_colsum else None)
    lib = _lib.load()
    gy = _f32c(gy, "gy").contiguous()
    g = torch.empty_like(gy) if (want_g and y is not None) else None
    cs = torch.empty(N, device=gy.device, dtype=torch.float32) if want_colsum else None
    if g is None and cs is None:
        return gy, None
    with torch.cuda.device(gy.device):
        rc = lib.ucnerf_relu_mask_colsum(gy.data_ptr(), None if y is None else y.contiguous().data_ptr(),
                                         None if g is None else g.data_ptr(), None if cs is None else cs.data_ptr(), M, N,
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "relu_mask_colsum")
    return (g if g is not None else gy), cs


class _TcLinear(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, relu, weight, bias, *xs):
        lead = xs[0].shape[:-1]
        x2 = [x.reshape(-1, x.shape[-1]) for x in xs]
        ks = [x.shape[1] for x in x2]
        _chk(sum(ks) == weight.shape[1], f"tc_linear: inputs have {sum(ks)} columns, weight expects {weight.shape[1]}")
        offs = [0]
        for k in ks:
            offs.append(offs[-1] + k)
        wsegs = [weight.detach()[:, offs[i]:offs[i + 1]].contiguous() for i in range(len(ks))]
        y = gemm_nt([(x.detach(), w) for x, w in zip(x2, wsegs)], None if bias is None else bias.detach(), relu)
        ctx.relu, ctx.offs, ctx.has_bias, ctx.lead = relu, offs, bias is not None, lead
        ctx.save_for_backward(weight, y if relu else None, *x2)
        return y.reshape(*lead, weight.shape[0])

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        weight, y, *x2 = ctx.saved_tensors
        offs = ctx.offs
        need = ctx.needs_input_grad            # (relu, weight, bias, *xs)
        # relu'(pre-activation) == (output > 0); the mask and the bias gradient (column sums) in one pass over the rows
        g, gb = relu_mask_colsum(gy.reshape(-1, weight.shape[0]).float().contiguous(), y if ctx.relu else None,
                                 want_g=True, want_colsum=ctx.has_bias and need[2])
        gw = None
        if need[1]:
            gw = torch.zeros_like(weight)
            for i, x in enumerate(x2):
                gemm_tn(g, x, out=gw[:, offs[i]:offs[i + 1]])
        gxs = []
        for i, x in enumerate(x2):
            if need[3 + i]:
                wt = weight.detach()[:, offs[i]:offs[i + 1]].t().contiguous()     # [k_s, N]: B of dx_s = g @ W_s
                gxs.append(gemm_nt([(g, wt)]).reshape(*ctx.lead, x.shape[1]))
            else:
                gxs.append(None)
        return (None, gw, gb, *gxs)


def tc_linear(xs: List[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False):
    """relu?([x_0 | x_1 | ...] @ weight.T + bias) on the tensor cores with fp32 accuracy; autograd for xs, weight, bias.
    weight [N, sum k_s] with N <= 256 and every k_s <= 256 (the input-gradient GEMM has N = k_s)."""
    if isinstance(xs, torch.Tensor):
        xs = [xs]
    _chk(weight.shape[0] <= 256 and all(x.shape[-1] <= 256 for x in xs), "tc_linear: widths up to 256")
    return _TcLinear.apply(bool(relu), weight, bias, *xs)
