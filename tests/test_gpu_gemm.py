"""GPU: the fp32-accurate tensor-core GEMMs of the training step (csrc/gemm3_tc.cu, 3xTF32 on tcgen05) and the nn.Linear
replacement built on them, against torch fp64 / fp32 matmuls of the same operands.  Shapes follow the reference's MLPs
(internal/models.py:L438-441, L475-483, L643-652): 40 -> 64 -> 256, [256 | 27] -> 256, [256 | 256 | 27] -> 256, 256 -> 3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


NT_CASES = [
    # M, N, [k_s], bias, relu
    (1000, 64, [40], True, True),              # density_layer.0 (ragged M, K = 40 -> chunk with 8 valid columns)
    (4096, 256, [64], True, False),            # density_layer.2
    (4099, 256, [256, 27], True, True),        # lin_second_stage_0 on [bottleneck | dir_enc]
    (2048, 256, [256, 256, 27], True, True),   # lin_second_stage_1 on [x | bottleneck | dir_enc]
    (777, 3, [256], True, False),              # rgb_layer (N = 3 -> padded MMA N = 16)
    (513, 1, [64], True, False),               # proposal density_layer.2 (N = 1)
    (300, 27, [256], False, False),            # an input gradient with N = 27
    (1, 256, [64], False, False),              # a single row
    (128 * 151 + 5, 256, [256], False, True),  # more tiles than SMs: the persistent loop wraps, accumulators alternate
]


@pytest.mark.parametrize("M,N,ks,bias,relu", NT_CASES)
def test_gemm_nt_matches_fp64(M, N, ks, bias, relu):
    from ucnerf_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(M + N)
    As = [torch.randn((M, k), device="cuda", generator=g) * (10.0 ** (i - 1)) for i, k in enumerate(ks)]
    W = torch.randn((N, sum(ks)), device="cuda", generator=g) / sum(ks) ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    off, pairs = 0, []
    for a in As:
        pairs.append((a, W[:, off:off + a.shape[1]].contiguous()))
        off += a.shape[1]
    got = gemm.gemm_nt(pairs, b, relu)
    gemm.status()
    ref = torch.cat(As, 1).double() @ W.double().T
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp_min(0)
    fp32 = torch.cat(As, 1) @ W.T + (b if bias else 0)
    if relu:
        fp32 = fp32.clamp_min(0)
    err, err32 = _rel(got, ref), _rel(fp32, ref)
    assert err < 5e-6, (err, err32)               # fp32-level: torch's own fp32 matmul sits at ~1e-6 on these
    # strided B (a column slice of the full weight, ldb = sum k) gives the same numbers as the contiguous copy
    off, pairs2 = 0, []
    for a in As:
        pairs2.append((a, W[:, off:off + a.shape[1]]))
        off += a.shape[1]
    got2 = gemm.gemm_nt(pairs2, b, relu)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("M,N1,N2", [(4096, 256, 256), (1000, 256, 283), (5000, 64, 40), (333, 3, 256), (31, 256, 64),
                                     (262144, 256, 256)])
def test_gemm_tn_matches_fp64(M, N1, N2):
    from ucnerf_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(M + N1 + N2)
    A = torch.randn((M, N1), device="cuda", generator=g) * 1e-3        # gradient-sized values: no scaling needed (TF32 range)
    B = torch.randn((M, N2), device="cuda", generator=g)
    C0 = torch.randn((N1, N2), device="cuda", generator=g) * 1e-3
    got = gemm.gemm_tn(A, B, out=C0.clone())                            # accumulates into `out`
    gemm.status()
    ref = C0.double() + A.double().T @ B.double()
    err = _rel(got, ref)
    # each CTA accumulates its ~M / 148 rows in ONE fp32 TMEM accumulator (hundreds of accumulate steps; the tensor core
    # truncates rather than rounds when it adds into the accumulator), then the partials meet through fp32 atomics:
    # 2.7e-5 at M = 262,144 - far inside what a weight gradient needs (the training tests compare at 1e-2)
    assert err < (1e-5 if M <= 8192 else 1e-4), err
    # into a column slice of a wider matrix (how the weight gradient of a concatenated input is assembled)
    wide = torch.zeros((N1, N2 + 40), device="cuda")
    gemm.gemm_tn(A, B, out=wide[:, 8:8 + N2])
    assert _rel(wide[:, 8:8 + N2], A.double().T @ B.double()) < (1e-5 if M <= 8192 else 1e-4)
    assert float(wide[:, :8].abs().max()) == 0 and float(wide[:, 8 + N2:].abs().max()) == 0


def test_tc_linear_forward_backward_matches_torch_linear():
    """The nn.Linear replacement on the reference's widest layer ([x | bottleneck | dir_enc] -> 256, relu) and the rgb
    layer: outputs and every gradient against torch's fp64 autograd of the same expression."""
    from ucnerf_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(5)
    M = 3000
    xs = [torch.randn((M, k), device="cuda", generator=g) for k in (256, 256, 27)]
    W = (torch.randn((256, 539), device="cuda", generator=g) / 539 ** 0.5).requires_grad_(True)
    b = torch.randn(256, device="cuda", generator=g).requires_grad_(True)
    R = (torch.randn((3, 256), device="cuda", generator=g) / 16).requires_grad_(True)
    r0 = torch.zeros(3, device="cuda", requires_grad=True)
    xs[0].requires_grad_(True)
    xs[1].requires_grad_(True)
    y = gemm.tc_linear(xs, W, b, relu=True)
    out = gemm.tc_linear([y], R, r0, relu=False)
    loss = (torch.sigmoid(out) * torch.arange(1, 4, device="cuda")).sum()
    loss.backward()
    gemm.status()
    got = [xs[0].grad, xs[1].grad, W.grad, b.grad, R.grad, r0.grad]
    xd = [x.detach().double().requires_grad_(x.requires_grad) for x in xs]
    Wd, bd, Rd, rd = (t.detach().double().requires_grad_(True) for t in (W, b, R, r0))
    yd = torch.relu(torch.cat(xd, 1) @ Wd.T + bd)
    outd = yd @ Rd.T + rd
    (torch.sigmoid(outd) * torch.arange(1, 4, device="cuda")).sum().backward()
    assert _rel(y, yd) < 5e-6 and _rel(out, outd) < 2e-5          # (the second layer inherits the round-off of the first)
    for a, r in zip(got, [xd[0].grad, xd[1].grad, Wd.grad, bd.grad, Rd.grad, rd.grad]):
        assert _rel(a, r) < 2e-5, _rel(a, r)
    assert xs[2].grad is None


def test_errors():
    from ucnerf_b200 import gemm
    with pytest.raises(RuntimeError, match="CUDA"):
        gemm.gemm_nt([(torch.zeros(4, 8), torch.zeros(4, 8))])
    with pytest.raises(RuntimeError):
        gemm.gemm_nt([(torch.zeros(4, 8, device="cuda"), torch.zeros(300, 8, device="cuda"))])      # N > 256
    with pytest.raises(RuntimeError, match="segment"):
        gemm.gemm_nt([(torch.zeros(4, 8, device="cuda"), torch.zeros(4, 9, device="cuda"))])


@pytest.mark.parametrize("M,N", [(5000, 256), (262144, 64), (77, 4), (1000, 3)])
def test_relu_mask_colsum(M, N):
    from ucnerf_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(M + N)
    gy = torch.randn((M, N), device="cuda", generator=g)
    y = torch.relu(torch.randn((M, N), device="cuda", generator=g))
    got_g, got_s = gemm.relu_mask_colsum(gy, y)
    ref_g = gy * (y > 0)
    assert torch.equal(got_g, ref_g)
    assert _rel(got_s, ref_g.double().sum(0)) < 1e-5
    g2, s2 = gemm.relu_mask_colsum(gy, None)
    assert g2 is gy or torch.equal(g2, gy)
    assert _rel(s2, gy.double().sum(0)) < 1e-5
