"""LayerNorm / im2col / attention kernels against PyTorch fp32 references."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from clipself_b200 import _lib
    _lib.require_device()
    return torch.device("cuda")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("D", [128, 768, 2048])
def test_layernorm(dev, dtype, D):
    from clipself_b200 import ops
    M = 1003
    x = (torch.randn(M, D, device=dev) * 2 + 0.5).to(dtype)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    ops.layernorm_fwd(x, M, D, g, b, 1e-6, y, mean=mean, rstd=rstd)
    ref = F.layer_norm(x.float(), (D,), g, b, 1e-6)
    assert (y.float() - ref).abs().max() <= 0.02 * ref.abs().max()
    torch.testing.assert_close(mean, x.float().mean(1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rstd, (x.float().var(1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-4, atol=1e-5)


def test_layernorm_row_maps(dev):
    from clipself_b200 import ops
    B, N, D = 4, 17, 128
    x = torch.randn(B * N, D, device=dev)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    ref = F.layer_norm(x, (D,), g, b, 1e-6).view(B, N, D)
    cls = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    ops.layernorm_fwd(x, B, D, g, b, 1e-6, cls, row_mul=N)
    assert (cls.float() - ref[:, 0]).abs().max() < 0.03
    pat = torch.empty(B * (N - 1), D, device=dev, dtype=torch.bfloat16)
    ops.layernorm_fwd(x, B * (N - 1), D, g, b, 1e-6, pat, row_div=N - 1, row_off=1)
    assert (pat.float().view(B, N - 1, D) - ref[:, 1:]).abs().max() < 0.03


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("S,P", [(224, 16), (64, 16), (42, 14)])
def test_im2col(dev, dtype, S, P):
    from clipself_b200 import ops
    B = 3
    img = torch.randn(B, 3, S, S, device=dev).to(dtype)
    k = 3 * P * P
    ld = (k + 7) // 8 * 8
    out = ops.im2col_patches(img, P, ld)
    ref = F.unfold(img.float(), kernel_size=P, stride=P).transpose(1, 2).reshape(-1, k)
    assert torch.equal(out[:, :k].float(), ref.to(torch.bfloat16).float())
    assert (out[:, k:] == 0).all()


@pytest.mark.parametrize("B,N,H", [(3, 197, 12), (2, 17, 2), (1, 577, 16), (2, 64, 1), (2, 65, 3),
                                   (2, 129, 2), (5, 208, 3), (3, 161, 1), (2, 193, 2), (150, 197, 12)])
def test_attention_fwd(dev, B, N, H):
    from clipself_b200 import ops
    D = H * 64
    torch.manual_seed(1000 * B + 10 * N + H)
    qkv = torch.randn(B * N, 3 * D, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    ops.attention_fwd(qkv, B, N, H, 0.125, out, lse)
    q, k, v = (t.reshape(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.float().view(B, N, 3, D).unbind(2))
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    err = (out.float() - ref).abs().max().item()
    print(f"attention B={B} N={N} H={H}: max err {err:.4e}")
    assert err < 2e-2
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), rtol=1e-3, atol=1e-3)
    if True:
        stats = torch.full((B * N, 4 * H, 2), float("nan"), device=dev)
        out2 = torch.empty_like(out)
        ops.attention_fwd(qkv, B, N, H, 0.125, out2, None, stats)
        assert torch.equal(out2, out)
        # the statistics are taken from the f32 accumulators before the bf16 rounding of the output:
        # compare with the f32 reference, tolerance = 32 elements x the kernel's per-element error
        o = ref.view(B * N, 4 * H, 16)
        torch.testing.assert_close(stats[..., 0], o.sum(-1), rtol=1e-2, atol=5e-2)
        torch.testing.assert_close(stats[..., 1], (o * o).sum(-1), rtol=1e-2, atol=5e-2)


@pytest.mark.parametrize("B,N,H", [(5, 197, 12), (3, 17, 2), (2, 577, 16), (2, 1024, 1), (9, 33, 3)])
def test_attention_cls_fwd(dev, B, N, H):
    """The CLS-query attention of the teacher's last block (cs_attention_cls_fwd) against torch and against row 0 of the
    full attention kernel."""
    from clipself_b200 import ops
    D = H * 64
    torch.manual_seed(31 * B + N + H)
    qkv = torch.randn(B * N, 3 * D, device=dev).to(torch.bfloat16)
    out = torch.full((B, D), float("nan"), device=dev, dtype=torch.bfloat16)
    stats = torch.full((B, 4 * H, 2), float("nan"), device=dev)
    ops.attention_cls_fwd(qkv, B, N, H, 0.125, out, stats)
    q, k, v = (t.reshape(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.float().view(B, N, 3, D).unbind(2))
    s = (q[:, :, :1] @ k.transpose(-1, -2)) * 0.125
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, D)
    err = (out.float() - ref).abs().max().item()
    print(f"cls attention B={B} N={N} H={H}: max err {err:.4e}")
    assert err < 1e-2
    o = ref.view(B, 4 * H, 16)
    torch.testing.assert_close(stats[..., 0], o.sum(-1), rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(stats[..., 1], (o * o).sum(-1), rtol=1e-2, atol=2e-2)
    full = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    ops.attention_fwd(qkv, B, N, H, 0.125, full)
    assert (out.float() - full.view(B, N, D)[:, 0].float()).abs().max().item() < 8e-3      # same rounding points, other summation order


@pytest.mark.parametrize("hin,hout", [(1024, 320), (896, 672), (64, 160), (224, 224), (37, 53)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resize_bilinear(dev, hin, hout, dtype):
    """--multiscale input resize (clipself.py:27) against torch's own bilinear interpolate."""
    from clipself_b200 import ops
    torch.manual_seed(hin + hout)
    x = torch.randn(2, 3, hin, hin, device=dev).to(dtype)
    y = ops.resize_bilinear(x, hout)
    ref = torch.nn.functional.interpolate(x.float(), size=(hout, hout), mode="bilinear").to(dtype)
    assert y.shape == ref.shape and y.dtype == dtype
    if dtype == torch.float32:
        torch.testing.assert_close(y, ref, rtol=0, atol=1e-5)          # same index math; FMA contraction only
    else:
        torch.testing.assert_close(y.float(), ref.float(), rtol=8e-3, atol=1e-6)   # 1 bf16 ulp
    if hin == hout:
        assert torch.equal(y, x)


@pytest.mark.parametrize("B,N,H", [(1, 225, 1), (2, 577, 16), (1, 1000, 2), (1, 4097, 12), (3, 129, 2)])
def test_attention_fwd_long_tc(dev, B, N, H, monkeypatch):
    from clipself_b200 import ops
    D = H * 64
    torch.manual_seed(7 * N + H)
    qkv = torch.randn(B * N, 3 * D, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    stats = torch.full((B * N, 4 * H, 2), float("nan"), device=dev)
    ops.attention_fwd(qkv, B, N, H, 0.125, out, lse, stats)
    torch.cuda.synchronize()
    q, k, v = (t.reshape(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.float().view(B, N, 3, D).unbind(2))
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    err = (out.float() - ref).abs().max().item()
    print(f"long tc attention B={B} N={N} H={H}: max err {err:.4e}")
    assert err < 2e-2
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), rtol=1e-3, atol=1e-3)
    o = ref.view(B * N, 4 * H, 16)
    torch.testing.assert_close(stats[..., 0], o.sum(-1), rtol=1e-2, atol=5e-2)
    torch.testing.assert_close(stats[..., 1], (o * o).sum(-1), rtol=1e-2, atol=5e-2)
