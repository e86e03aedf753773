"""Numerical model of the tensor-core operand split used by `pw_mma_kernel` (CPU only).

The kernel feeds fp32 activations and weights to bf16 tensor cores as
    x ~ xh + xl,  xh = x truncated to bf16 (top 16 bits),  xl = bf16_rn(x - xh)     (pointwise_mma.cu transform loop)
    w ~ wh + wl,  wh = bf16_rn(w),                        wl = bf16_rn(w - wh)     (pack_weight_mma_kernel)
and accumulates xh*wh + xl*wh + xh*wl in fp32.  This test pins the error budget that design rests on
(~1e-5 relative, against 5e-3 for a single bf16 pass and the 1e-3 parity budget), independently of
the GPU: the dropped term is xl*wl = O(2^-16 |x||w|) and each kept term is exact in fp32 before accumulation."""
import torch


def bf16_trunc(x):
    return (x.view(torch.int32) & -65536).view(torch.float32)


def bf16_rn(x):
    return x.to(torch.bfloat16).to(torch.float32)


def split_gemm(W, X):
    xh = bf16_trunc(X)
    xl = bf16_rn(X - xh)
    wh = bf16_rn(W)
    wl = bf16_rn(W - wh)
    assert torch.equal(X - xh, (X.double() - xh.double()).float())          # the remainder is exact in fp32
    return (wh @ xh) + (wh @ xl) + (wl @ xh)                                 # fp32 accumulation of exact products


def test_three_product_split_is_fp32_grade():
    g = torch.Generator().manual_seed(0)
    for K in (64, 256, 512, 2048):
        X = torch.randn(K, 640, generator=g) * 2 + 0.5
        W = torch.randn(256, K, generator=g) / K ** 0.5
        want = W.double() @ X.double()
        scale = want.abs().max()
        err3 = float((split_gemm(W, X).double() - want).abs().max() / scale)
        err1 = float(((bf16_rn(W) @ bf16_rn(X)).double() - want).abs().max() / scale)
        fp32 = float(((W @ X).double() - want).abs().max() / scale)
        assert err3 < 2e-5, (K, err3)                 # two orders inside the 1e-3 parity budget
        assert err1 > 20 * err3, (K, err1, err3)      # what one bf16 pass would cost
        assert err3 < 60 * max(fp32, 1e-7), (K, err3, fp32)


def test_split_pieces_are_bf16_representable():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10000, generator=g) * torch.logspace(-20, 20, 10000)
    xh = bf16_trunc(x)
    xl = bf16_rn(x - xh)
    for t in (xh, xl):
        assert torch.equal(t, bf16_rn(t))
    # |x - xh - xl| <= 2^-9 |x - xh| <= 2^-16 |x|  (bound quoted in the kernel)
    rest = (x.double() - xh.double() - xl.double()).abs()
    assert bool((rest <= 2.0 ** -16 * x.double().abs() + 1e-300).all())
