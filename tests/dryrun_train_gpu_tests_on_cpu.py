"""TEST INFRASTRUCTURE - never imported by the product, never collected by pytest (run by hand:
`python tests/dryrun_train_gpu_tests_on_cpu.py`, after the CPU suite has built tests/_build/cpu_harness.so).

Rehearses the `-m gpu` tests of the training-step ops (tests/test_train_*.py) in a container WITHOUT a GPU, so that their
Python plumbing (argument order of the ctypes calls, shapes, autograd wiring) and their tolerances are exercised before
GPU minutes are spent: (1) `.cuda()` becomes a copy, (2) the product's Python wrappers are loaded from source with their
`device.type == "cuda"` checks inverted, (3) the C-ABI calls are served by a stand-in whose methods have the C
signatures of include/ucnerf_b200.h and forward the raw host pointers to the serial CPU instantiation of the SAME
algorithm templates (tests/cpu_harness.cpp).  It proves nothing about the CUDA instantiation itself - that is what the
real `-m gpu` run is for - and it is not a CPU path of the product: the product raises without CUDA
(tests/test_host_logic.py::test_renderer_refuses_to_run_without_cuda)."""
import sys, types, ctypes, contextlib, importlib, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
H = ctypes.CDLL(ROOT + '/tests/_build/cpu_harness.so')
vp = ctypes.c_void_p; cf = ctypes.c_float

torch.Tensor.cuda = lambda self, *a, **k: self.detach().clone().requires_grad_(False) if not self.requires_grad else self.clone()
torch.nn.Module.cuda = lambda self, *a, **k: self
class _S: cuda_stream = 0
torch.cuda.current_stream = lambda *a, **k: _S()
torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
_OrigGen = torch.Generator
def Gen(device=None): return _OrigGen()
torch.Generator = Gen
_rand, _randn = torch.rand, torch.randn
def _strip(k): k.pop('device', None); return k
torch.rand = lambda *a, **k: _rand(*a, **_strip(k))
torch.randn = lambda *a, **k: _randn(*a, **_strip(k))
_empty, _zeros, _linspace = torch.empty, torch.zeros, torch.linspace

class FakeLib:
    def ucnerf_last_error(self): return b"fake"
    def ucnerf_resample_intervals(self, t, w, N, n, dil, dilation, anneal, padding, S, u, jit, cols, out, stream):
        H.h_resample_jitter(N, n, vp(t), vp(w), dil, cf(dilation), cf(anneal), cf(padding), S, vp(u), vp(jit) if jit else None, cols, vp(out)); return 0
    def ucnerf_cast_rays(self, t, o, d, c, r, rv, rot, flip, N, S, ss, means, stds, ts, stream):
        H.h_cast_rays(N, S, vp(t), vp(o), vp(d), vp(c), vp(r), vp(rv), vp(rot) if rot else None, vp(flip) if flip else None, cf(ss), vp(means), vp(stds), vp(ts)); return 0
    def ucnerf_composite_train_forward(self, t, d, c, dr, N, S, bg, w, rgb, acc, stream):
        H.h_composite_train_forward(N, S, vp(t), vp(d), vp(c) if c else None, vp(dr), cf(bg), vp(w), vp(rgb), vp(acc)); return 0
    def ucnerf_composite_train_backward(self, t, d, c, dr, w, acc, gw, gr, ga, N, S, bg, dd, dc, stream):
        H.h_composite_train_backward(N, S, vp(t), vp(d), vp(c) if c else None, vp(dr), cf(bg), vp(w), vp(acc), vp(gw) if gw else None,
                                     vp(gr) if gr else None, vp(ga) if ga else None, vp(dd), vp(dc) if dc else None); return 0
    def ucnerf_pooled_encode_forward(self, m, s, B, M, flags, emb, offs, gsz, L, C, S, Hres, feats, coord, stream):
        H.h_pooled_forward(B, M, flags & 1, L, vp(offs), vp(gsz), cf(S), Hres, vp(emb), vp(m), vp(s), vp(feats), vp(coord) if coord else None); return 0
    def ucnerf_pooled_encode_backward(self, g, m, s, B, M, flags, offs, gsz, L, C, S, Hres, ge, stream):
        offs_np = np.ctypeslib.as_array(ctypes.cast(offs, ctypes.POINTER(ctypes.c_int32)), shape=(L + 1,))
        T = int(offs_np[-1])
        acc = np.zeros((T, 4), np.float64)
        H.h_pooled_backward(B, M, flags, L, vp(offs), vp(gsz), cf(S), Hres, vp(g), vp(m), vp(s), acc.ctypes.data_as(vp))
        dst = np.ctypeslib.as_array(ctypes.cast(ge, ctypes.POINTER(ctypes.c_float)), shape=(T, 4))
        dst += acc.astype(np.float32); return 0

import ucnerf_b200._lib as L
fake = FakeLib()
L.load = lambda: fake
L.check = lambda rc, what="": (_ for _ in ()).throw(RuntimeError(what)) if rc else None

def load_inverted(modname, path):
    src = open(path).read().replace('!= "cuda"', '!= "cpu"').replace('== "cuda"', '== "cpu"').replace('device_type="cuda"', 'device_type="cpu"')
    mod = types.ModuleType(modname); mod.__file__ = path; mod.__package__ = modname.rpartition('.')[0]
    sys.modules[modname] = mod
    exec(compile(src, path, 'exec'), mod.__dict__)
    return mod
import ucnerf_b200, ucnerf_b200.gridencoder
load_inverted('ucnerf_b200.gridencoder.pooled', ROOT + '/ucnerf_b200/gridencoder/pooled.py')
load_inverted('ucnerf_b200.stepfun', ROOT + '/ucnerf_b200/stepfun.py')
load_inverted('ucnerf_b200.render_train', ROOT + '/ucnerf_b200/render_train.py')

# the tensor-core dense layers (ucnerf_b200.gemm.tc_linear, CUDA only) are served by torch's own fp32 linear here: the
# rehearsal covers the wiring of _mlp_forward_native (segments instead of torch.cat), tests/test_gpu_gemm.py the kernels
import torch as _torch
import ucnerf_b200.gemm as _G
def _tc_linear_cpu(xs, weight, bias=None, relu=False):
    y = _torch.nn.functional.linear(_torch.cat(list(xs), dim=-1), weight, bias)
    return _torch.relu(y) if relu else y
_G.tc_linear = _tc_linear_cpu

import pytest
import re
NOT_EXERCISED = []
@contextlib.contextmanager
def strict_raises(exc, match=None):
    """pytest.raises with its real semantics (must raise, message must match).  The one exception: assertions about
    errors only the CUDA build can produce (device checks - the rehearsal inverts them to run on CPU tensors) are recorded
    as not exercised instead of being reported as passed."""
    try:
        yield
    except exc as e:
        if match is not None and not re.search(match, str(e)):
            raise AssertionError(f"raised {e!r}, which does not match {match!r}")
        return
    if match is not None and re.search(r"cuda|CUDA|device", match):
        NOT_EXERCISED.append(match)
        return
    raise AssertionError(f"DID NOT RAISE {exc}")
pytest.raises = strict_raises

import inspect
def run_module(name, skip=()):
    mod = importlib.import_module(name)
    fixtures = {}
    ok = 0
    for tname, fn in list(vars(mod).items()):
        if not tname.startswith('test_') or not callable(fn): continue
        marks = [m for m in getattr(fn, 'pytestmark', [])]
        if not any(m.name == 'gpu' for m in marks): continue
        if tname in skip: print('  skip', tname); continue
        params = [m for m in marks if m.name == 'parametrize']
        combos = [{}]
        for pm in params:
            names = [n.strip() for n in pm.args[0].split(',')]
            new = []
            for c in combos:
                for v in pm.args[1]:
                    vals = v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,)
                    d = dict(c); d.update(dict(zip(names, vals))); new.append(d)
            combos = new
        sig = inspect.signature(fn)
        for c in combos:
            kw = dict(c)
            for p in sig.parameters:
                if p in kw: continue
                if p in ('case', 'gold'):
                    fx = getattr(mod, p)
                    inner = getattr(fx, '__wrapped__', None) or getattr(fx, '__pytest_wrapped__').obj
                    kw[p] = fixtures.setdefault((name, p), inner())
            fn(**kw); ok += 1
            print('  ok', tname, c)
    return ok

for m, skip in (('test_train_resample', ()), ('test_train_render_composite', ()), ('test_train_sample_cast_rays', ()),
                ('test_train_pooled_encode', ('test_cuda_equals_the_unfused_gridencoder_chain_and_errors', 'test_cuda_full_training_size_properties')),
                ('test_train_zz_level_chain', ()), ('test_train_zzz_forward', ('test_mirror_model_forward_dispatch_eval_is_the_fused_path_and_follows_weight_updates',))):
    print(m); run_module(m, skip)
print("CUDA-only error paths not exercised:", NOT_EXERCISED)
print("ALL DRY RUNS OK")
