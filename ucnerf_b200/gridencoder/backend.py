"""`_backend` of the gridencoder mirror: the three functions of the reference's pybind module
(gridencoder/src/bindings.cpp:L5-9) with identical signatures and error behaviour, implemented as thin
ctypes calls into libucnerf_b200.so on torch's current CUDA stream.

Error behaviour follows the reference's TORCH_CHECKs (gridencoder.cu:L15-18, L449-465): RuntimeError for
non-CUDA / non-contiguous tensors and wrong dtypes, RuntimeError for unsupported D / C (L381, L398)."""
import torch

from .. import _lib

_DTYPE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.float64: _lib.F64}


def _chk(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _cuda_contig(t, name):
    _chk(t.device.type == "cuda", f"{name} must be a CUDA tensor")
    _chk(t.is_contiguous(), f"{name} must be a contiguous tensor")


def _floating(t, name):
    _chk(t.dtype in _DTYPE, f"{name} must be a floating tensor")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners,
                        interp):
    for t, n in ((inputs, "inputs"), (embeddings, "embeddings"), (offsets, "offsets"), (outputs, "outputs")):
        _cuda_contig(t, n)
    for t, n in ((inputs, "inputs"), (embeddings, "embeddings"), (outputs, "outputs")):
        _floating(t, n)
    _chk(offsets.dtype == torch.int32, "offsets must be an int tensor")
    _chk(inputs.dtype == torch.float32, "inputs must be float32")
    _chk(outputs.dtype == embeddings.dtype, "outputs must have the embeddings dtype")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.ucnerf_grid_encode_forward(_ptr(inputs), _ptr(embeddings), _ptr(offsets), _ptr(outputs), B, D, C, L,
                                            float(S), H, _ptr(dy_dx), gridtype, int(bool(align_corners)), interp,
                                            _DTYPE[embeddings.dtype], _stream())
    _lib.check(rc, "grid_encode_forward")


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                         gridtype, align_corners, interp):
    for t, n in ((grad, "grad"), (inputs, "inputs"), (embeddings, "embeddings"), (offsets, "offsets"),
                 (grad_embeddings, "grad_embeddings")):
        _cuda_contig(t, n)
    for t, n in ((grad, "grad"), (inputs, "inputs"), (embeddings, "embeddings"), (grad_embeddings, "grad_embeddings")):
        _floating(t, n)
    _chk(offsets.dtype == torch.int32, "offsets must be an int tensor")
    _chk(grad.dtype == grad_embeddings.dtype, "grad and grad_embeddings must share a dtype")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.ucnerf_grid_encode_backward(_ptr(grad), _ptr(inputs), _ptr(embeddings), _ptr(offsets),
                                             _ptr(grad_embeddings), B, D, C, L, float(S), H, _ptr(dy_dx),
                                             _ptr(grad_inputs), gridtype, int(bool(align_corners)), interp,
                                             _DTYPE[grad.dtype], _stream())
    _lib.check(rc, "grid_encode_backward")


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    _chk(embeddings.dtype in _DTYPE, "embeddings must be a floating tensor")
    _chk(inputs.dtype == embeddings.dtype and grad.dtype == embeddings.dtype,
         "inputs / grad must have the embeddings dtype")
    for t, n in ((inputs, "inputs"), (embeddings, "embeddings"), (grad, "grad"), (offsets, "offsets")):
        _cuda_contig(t, n)
    lib = _lib.load()
    with torch.cuda.device(embeddings.device):
        rc = lib.ucnerf_grad_total_variation(_ptr(inputs), _ptr(embeddings), _ptr(grad), _ptr(offsets), float(weight),
                                             B, D, C, L, float(S), H, gridtype, int(bool(align_corners)),
                                             _DTYPE[embeddings.dtype], _stream())
    _lib.check(rc, "grad_total_variation")
