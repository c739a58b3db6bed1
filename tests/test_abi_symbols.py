"""CPU: the C-ABI library loads without a GPU and exports every symbol include/ucnerf_b200.h declares.
No compute entry point is called here."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ucnerf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ucnerf_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ucnerf_b200.h but not exported"
    from ucnerf_b200 import _lib
    assert sorted(_lib.EXPORTS) == syms


def test_abi_version_and_error_string(lib):
    assert lib.ucnerf_abi_version() == 1
    assert isinstance(lib.ucnerf_last_error(), bytes)


def test_struct_sizes_match_header(lib):
    """ctypes mirrors of the PODs must have the C layout (LP64)."""
    from ucnerf_b200 import _lib
    assert ctypes.sizeof(_lib.MlpDesc) == 3 * 8 + 4 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.ModelDesc) == 6 * 4 + 8 * 8 + 5 * ctypes.sizeof(_lib.MlpDesc) + 6 * 8
    assert ctypes.sizeof(_lib.Rays) == 8 * 8
    assert ctypes.sizeof(_lib.Outputs) == (8 + 5 + 5 + 4) * 8
    assert ctypes.sizeof(_lib.Camera) == 21 * 8 + 4 * 4 + 8
    assert ctypes.sizeof(_lib.RayBuffers) == 9 * 8


def test_argument_validation_without_gpu(lib):
    """Bad D / C are rejected before any CUDA call (reference: std::runtime_error, gridencoder.cu:L381,L398)."""
    rc = lib.ucnerf_grid_encode_forward(None, None, None, None, 0, 7, 4, 1, 1.0, 16, None, 0, 0, 0, 0, None)
    assert rc != 0 and b"D must be" in lib.ucnerf_last_error()
    rc = lib.ucnerf_grid_encode_forward(None, None, None, None, 0, 3, 3, 1, 1.0, 16, None, 0, 0, 0, 0, None)
    assert rc != 0 and b"C must be" in lib.ucnerf_last_error()
    rc = lib.ucnerf_grid_encode_forward(None, None, None, None, 0, 3, 4, 1, 1.0, 16, None, 0, 0, 0, 9, None)
    assert rc != 0 and b"dtype" in lib.ucnerf_last_error()
    # empty batch is a no-op success
    assert lib.ucnerf_grid_encode_forward(None, None, None, None, 0, 3, 4, 1, 1.0, 16, None, 0, 0, 0, 0, None) == 0


def test_training_ops_validate_arguments_without_gpu(lib):
    """The training-step entry points reject bad shapes before any CUDA call; empty batches are no-op successes."""
    import numpy as np
    offs = np.array([0, 8, 16], np.int32)
    gsz = np.array([17, 33], np.int32)
    op, gp = offs.ctypes.data, gsz.ctypes.data
    assert lib.ucnerf_pooled_encode_forward(None, None, 0, 6, 1, None, op, gp, 2, 4, 1.0, 16, None, None, None) == 0
    assert lib.ucnerf_pooled_encode_forward(None, None, 0, 6, 1, None, op, gp, 2, 2, 1.0, 16, None, None, None) != 0
    assert b"level_dim" in lib.ucnerf_last_error()
    assert lib.ucnerf_pooled_encode_backward(None, None, None, 0, 9, 1, op, gp, 2, 4, 1.0, 16, None, None) != 0
    assert b"multisample" in lib.ucnerf_last_error()
    assert lib.ucnerf_pooled_encode_forward(None, None, 5, 6, 1, None, op, gp, 2, 4, 1.0, 16, None, None, None) != 0
    assert b"null" in lib.ucnerf_last_error()
    assert lib.ucnerf_resample_intervals(None, None, 0, 1, 0, 0.0, 1.0, 0.0, 32, None, None, 0, None, None) == 0
    assert lib.ucnerf_resample_intervals(None, None, 4, 1, 0, 0.0, 1.0, 0.0, 1, None, None, 0, None, None) != 0
    assert b"S >= 2" in lib.ucnerf_last_error()
    assert lib.ucnerf_composite_train_forward(None, None, None, None, 0, 32, 1.0, None, None, None, None) == 0
    assert lib.ucnerf_composite_train_forward(None, None, None, None, 3, 0, 1.0, None, None, None, None) != 0
    assert lib.ucnerf_cast_rays(None, None, None, None, None, None, None, None, 0, 8, 0.5, None, None, None, None) == 0
    assert lib.ucnerf_cast_rays(None, None, None, None, None, None, None, None, 2, 8, 0.5, None, None, None, None) != 0
    assert lib.ucnerf_composite_train_backward(None, None, None, None, None, None, None, None, None, 3, 8, 1.0, None, None, None) != 0
    assert b"null" in lib.ucnerf_last_error()


def test_ctypes_argtypes_match_the_header_declarations(lib):
    """Every function declared in include/ucnerf_b200.h has ctypes argtypes of the same length and kind (pointer vs the
    exact scalar type) in ucnerf_b200/_lib.py - the Python wrappers cannot drift from the C ABI unnoticed."""
    txt = open(os.path.join(ROOT, "include", "ucnerf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    decls = re.findall(r"\b(ucnerf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    scalar = {"uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "int": ctypes.c_int, "float": ctypes.c_float,
              "double": ctypes.c_double, "uint64_t": ctypes.c_uint64, "int64_t": ctypes.c_int64}
    assert len(decls) >= 27
    for name, params in decls:
        plist = [q.strip() for q in params.split(",") if q.strip() and q.strip() != "void"]
        at = getattr(lib, name).argtypes
        if not plist:
            continue
        assert at is not None and len(at) == len(plist), (name, at, plist)
        for i, (q, a) in enumerate(zip(plist, at)):
            if "*" in q:
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, ctypes._Pointer), (name, i, q, a)
            else:
                base = re.sub(r"\bconst\b", "", q).split()[0]
                assert scalar.get(base) is a, (name, i, q, a)
