"""GPU parity of the convolution kernels against a plain PyTorch fp32 reference (F.conv2d on CPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(x, w):
    n, _, cin = x.shape
    xt = torch.from_numpy(x).reshape(n, 19, 19, cin).permute(0, 3, 1, 2).double()
    y = F.conv2d(xt, torch.from_numpy(w).double(), padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).reshape(n, 361, -1).numpy()


def _bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("cin,cout,k", [(16, 8, 1), (8, 8, 3), (64, 128, 3), (256, 128, 1), (96, 40, 3)])
def test_conv_fp32(cin, cout, k):
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(cin * 1000 + cout + k)
    x = rng.standard_normal((3, 361, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    y = E.conv_test(x, w, E.PRECISION_FP32)
    ref = _ref_conv(x, w)
    assert np.abs(y - ref).max() < 2e-5  # fp32 accumulation over K <= 2304


@pytest.mark.parametrize("cin,cout,k,n", [(64, 64, 1, 2), (128, 128, 3, 3), (256, 128, 1, 2), (128, 256, 1, 5),
                                          (256, 96, 1, 2), (192, 192, 3, 2), (384, 192, 1, 1), (192, 384, 1, 1),
                                          (64, 64, 3, 40)])
def test_conv_tcgen05(cin, cout, k, n):
    """tcgen05 path: bf16 operands, fp32 accumulate. Against the fp64 conv of the SAME bf16-rounded operands the
    only error is fp32 accumulation order, so the bound is tight (1e-3 would hide a wrong tap or swizzle)."""
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(cin * 1000 + cout + k)
    x = _bf16_round(rng.standard_normal((n, 361, cin)).astype(np.float32))
    w = _bf16_round((rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32))
    y = E.conv_test(x, w, E.PRECISION_BF16)
    ref = _ref_conv(x, w)
    err = np.abs(y - ref).max()
    assert err < 5e-5, f"max abs err {err}"


@pytest.mark.parametrize("cin,cout,n", [(64, 64, 3), (128, 128, 40), (128, 64, 2), (64, 128, 5), (192, 192, 3), (128, 128, 1),
                                        (64, 32, 2), (128, 256, 7)])
def test_conv_tcgen05_resident_weights(cin, cout, n, monkeypatch):
    """The resident-weight / tap-reuse 3x3 kernel (shifted shared-memory views of one haloed A tile). It only writes the
    bf16 activated copy, so the comparison allows one bf16 rounding of the output (2^-8 relative)."""
    from p3achygo_b200 import engine as E
    monkeypatch.setenv("P3_CONV_TEST_ACT", "1")
    rng = np.random.default_rng(cin * 7 + cout)
    x = _bf16_round(rng.standard_normal((n, 361, cin)).astype(np.float32))
    w = _bf16_round((rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))
    y = E.conv_test(x, w, E.PRECISION_BF16)
    ref = _ref_conv(x, w)
    assert np.all(np.abs(y - ref) <= np.abs(ref) * 2.0 ** -8 + 1e-4), f"max abs err {np.abs(y - ref).max()}"


def _mish(x):
    return x * np.tanh(np.log1p(np.exp(x)))


@pytest.mark.parametrize("C,n,prec", [(16, 2, "fp32"), (128, 3, "fp32"), (128, 3, "bf16"), (256, 5, "bf16"), (384, 2, "bf16"),
                                      (192, 2, "bf16")])
def test_broadcast_mix(C, n, prec):
    """Board-mixing Dense(361->361) of the broadcast block (python/model.py:570-581) vs numpy in fp64."""
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(C + n)
    x = rng.standard_normal((n, 361, C)).astype(np.float32)
    w = (rng.standard_normal((361, 361)) / 19.0).astype(np.float32)
    bias = (rng.standard_normal(361) * 0.1).astype(np.float32)
    if prec == "bf16":
        x, w = _bf16_round(x), _bf16_round(w)
    y = E.broadcast_test(x, w, bias, E.PRECISION_BF16 if prec == "bf16" else E.PRECISION_FP32)
    ref = _mish(np.einsum("pq,bpc->bqc", w.astype(np.float64), x.astype(np.float64)) + bias[None, :, None])
    if prec == "fp32":
        assert np.abs(y - ref).max() < 2e-5
    else:
        assert np.all(np.abs(y - ref) <= np.abs(ref) * 2.0 ** -8 + 2e-4), f"max abs err {np.abs(y - ref).max()}"


def _f16_round(a):
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("k1,n1,n2,n", [(128, 256, 128, 3), (64, 128, 64, 5), (128, 256, 128, 41), (64, 128, 128, 2),
                                        (128, 256, 256, 4), (256, 256, 128, 4), (128, 128, 128, 37)])
@pytest.mark.parametrize("fused", [True, False])
def test_block_boundary(k1, n1, n2, n, fused):
    """Expand 1x1 + residual -> BN/mish -> reduce 1x1 -> BN/mish (python/model.py:372-427, 276-281) on the fused CTA-pair
    kernel (chain_tc.cu) and on the two stand-alone 1x1 launches (pw_tc.cu), against numpy in fp64 on the SAME bf16 / fp16
    rounded operands.  x' is exact up to its fp16 store; `out` additionally sees the bf16 rounding of u and of itself."""
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(k1 + 7 * n1 + 13 * n2 + n)
    t = _bf16_round(rng.standard_normal((n, 361, k1)).astype(np.float32))
    x = _f16_round(rng.standard_normal((n, 361, n1)).astype(np.float32) * 2.0)
    w1 = _bf16_round((rng.standard_normal((n1, k1)) / np.sqrt(k1)).astype(np.float32))
    w2 = _bf16_round((rng.standard_normal((n2, n1)) / np.sqrt(n1)).astype(np.float32))
    s1 = rng.uniform(0.5, 1.5, n1).astype(np.float32)
    h1 = rng.uniform(-0.5, 0.5, n1).astype(np.float32)
    s2 = rng.uniform(0.5, 1.5, n2).astype(np.float32)
    h2 = rng.uniform(-0.5, 0.5, n2).astype(np.float32)
    xp, out = E.block_boundary_test(t, x, w1, w2, s1, h1, s2, h2, fused)
    ref_x = x.astype(np.float64) + t.astype(np.float64) @ w1.astype(np.float64).T
    # x' is stored as fp16: half an ulp of fp16 (2^-11 relative) plus fp32 accumulation noise
    assert np.all(np.abs(xp - ref_x) <= np.abs(ref_x) * 2.0 ** -11 + 2e-4), f"x' max abs err {np.abs(xp - ref_x).max()}"
    u = _bf16_round(_mish(ref_x * s1 + h1).astype(np.float32)).astype(np.float64)
    ref_out = _mish((u @ w2.astype(np.float64).T) * s2 + h2)
    # u is rounded to bf16 before the second GEMM (a value on a rounding boundary may flip: n1 terms of 2^-9 relative,
    # averaged down by the sum) and `out` is rounded to bf16: a pre-rounding difference can move it by one whole bf16 ulp
    err = np.abs(out - ref_out)
    assert np.all(err <= np.abs(ref_out) * 2.0 ** -7 + 6e-3), f"out max abs err {err.max()}"
    assert err.mean() < 2.5e-3  # ~a quarter ulp of bf16 at |out| ~ 1


def test_block_boundary_fused_equals_unfused():
    """The fused launch and the two stand-alone launches compute the same function with the same roundings: bit-identical
    x', and identical `out` (same operands, same MMA shapes along K)."""
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(5)
    n, k1, n1, n2 = 6, 128, 256, 128
    t = _bf16_round(rng.standard_normal((n, 361, k1)).astype(np.float32))
    x = _f16_round(rng.standard_normal((n, 361, n1)).astype(np.float32))
    w1 = _bf16_round((rng.standard_normal((n1, k1)) / np.sqrt(k1)).astype(np.float32))
    w2 = _bf16_round((rng.standard_normal((n2, n1)) / np.sqrt(n1)).astype(np.float32))
    s1, h1 = np.ones(n1, np.float32), np.zeros(n1, np.float32)
    s2, h2 = np.ones(n2, np.float32), np.zeros(n2, np.float32)
    xa, oa = E.block_boundary_test(t, x, w1, w2, s1, h1, s2, h2, True)
    xb, ob = E.block_boundary_test(t, x, w1, w2, s1, h1, s2, h2, False)
    assert np.array_equal(xa, xb)
    assert np.abs(oa - ob).max() <= 2.0 ** -7 * max(1.0, float(np.abs(ob).max()))  # at most one bf16 ulp apart
