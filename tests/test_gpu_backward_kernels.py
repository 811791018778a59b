"""Backward kernels against PyTorch autograd references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from clipself_b200 import _lib
    _lib.require_device()
    return torch.device("cuda")


@pytest.mark.parametrize("B,N,H,rope", [(2, 197, 12, True), (3, 17, 2, True), (1, 577, 16, False), (2, 64, 1, False),
                                        (2, 65, 3, True)])
def test_attention_bwd(dev, B, N, H, rope):
    from clipself_b200 import ops
    from clipself_b200.tower import rope_tables
    D = H * 64
    torch.manual_seed(0)
    raw = (torch.randn(B, N, 3, H, 64, device=dev) * 0.7).requires_grad_(True)
    cos = sin = None
    if rope:
        g = int(round((N - 1) ** 0.5))
        assert g * g == N - 1
        cos, sin = (t.to(dev) for t in rope_tables(g, 64, 16))

    def rot(t):   # t [B, N-1, 2, H, 64]
        pairs = t.reshape(*t.shape[:-1], 32, 2)
        r = torch.stack((-pairs[..., 1], pairs[..., 0]), -1).reshape(t.shape)
        return t * cos[None, :, None, None, :] + r * sin[None, :, None, None, :]

    if rope:
        qk = torch.cat([raw[:, :1, :2], rot(raw[:, 1:, :2])], dim=1)
        x = torch.cat([qk, raw[:, :, 2:]], dim=2)
    else:
        x = raw
    xb = x.to(torch.bfloat16)                                  # what the forward kernel consumes
    q, k, v = (t.permute(0, 2, 1, 3).float() for t in xb.unbind(2))
    s = (q @ k.transpose(-1, -2)) * 0.125
    o_ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    d_out = torch.randn(B * N, D, device=dev).to(torch.bfloat16)
    o_ref.backward(d_out.float())
    ref = raw.grad.reshape(B * N, 3 * D)

    qkv = xb.detach().reshape(B * N, 3 * D).contiguous()
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    ops.attention_fwd(qkv, B, N, H, 0.125, out, lse)
    dqkv = torch.full((B * N, 3 * D), float("nan"), device=dev, dtype=torch.bfloat16)
    delta = torch.empty(B * H * N, device=dev)
    ops.attention_bwd(qkv, out, d_out, lse, B, N, H, 0.125, (cos, sin) if rope else None, delta, dqkv)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        a, b = dqkv[:, sl].float(), ref[:, sl]
        r = ((a - b).norm() / b.norm()).item()
        print(f"attention_bwd B={B} N={N} H={H} rope={rope} {name}: rel-L2 {r:.3e}")
        assert r < 2e-2, name


@pytest.mark.parametrize("xdt,dydt,outdt", [(torch.float32, torch.bfloat16, torch.float32),
                                            (torch.bfloat16, torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("D", [128, 768, 2048])
def test_layernorm_bwd(dev, xdt, dydt, outdt, D):
    from clipself_b200 import ops
    M = 517
    torch.manual_seed(1)
    x = (torch.randn(M, D, device=dev) * 1.5 + 0.3).to(xdt)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    dy = torch.randn(M, D, device=dev).to(dydt)
    xr = x.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy.float())
    y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    ops.layernorm_fwd(x, M, D, g, b, 1e-6, y, mean=mean, rstd=rstd)
    add = torch.randn(M, D, device=dev) if outdt == torch.float32 else None
    dx = add.clone() if add is not None else torch.empty(M, D, device=dev, dtype=outdt)
    ops.layernorm_bwd_dx(dy, x, M, D, mean, rstd, g, dx, add=dx if add is not None else None)
    ref = xr.grad + (add if add is not None else 0)
    tol = 2e-2 if outdt == torch.bfloat16 else 2e-4
    assert ((dx.float() - ref).norm() / ref.norm()).item() < tol
    dgamma, dbeta = torch.empty(D, device=dev), torch.empty(D, device=dev)
    ws = torch.empty(128 * 2 * D, device=dev)
    ops.col_reduce(dy, M, D, dbeta, ws, x=x, mean=mean, rstd=rstd, dgamma=dgamma)
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-3, atol=2e-3)
    ops.col_reduce(dy, M, D, dbeta, ws)
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("M,N", [(100, 64), (12608, 768), (777, 2048), (33, 40)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_cast_transpose(dev, M, N, dt):
    from clipself_b200 import ops
    src = torch.randn(M, N, device=dev).to(dt)
    Mpad = (M + 7) // 8 * 8
    dst = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    dst_t = torch.zeros(N, Mpad, device=dev, dtype=torch.bfloat16)
    ops.cast_transpose(src, M, N, dst=dst, dst_t=dst_t)
    assert torch.equal(dst, src.to(torch.bfloat16))
    assert torch.equal(dst_t[:, :M], src.to(torch.bfloat16).t())


@pytest.mark.parametrize("split", [True, False])
def test_swiglu_fwd_bwd(dev, split):
    from clipself_b200 import ops
    M, Hd = 300, 384
    torch.manual_seed(2)
    gate = torch.randn(M, Hd, device=dev).to(torch.bfloat16)
    up = torch.randn(M, Hd, device=dev).to(torch.bfloat16)
    if split:
        x12 = torch.cat([gate, up], dim=1).contiguous()
    else:
        x12 = torch.stack([gate.view(M, Hd // 128, 128), up.view(M, Hd // 128, 128)], dim=2).reshape(M, 2 * Hd).contiguous()
    gr, ur = gate.float().requires_grad_(True), up.float().requires_grad_(True)
    href = F.silu(gr) * ur
    dh = torch.randn(M, Hd, device=dev).to(torch.bfloat16)
    href.backward(dh.float())
    h = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
    ops.swiglu_fwd(x12, M, Hd, h, split=split)
    assert (h.float() - href).abs().max() < 0.03
    dx12 = torch.empty_like(x12)
    ops.swiglu_bwd(x12, dh, M, Hd, dx12, split=split)
    if split:
        dg, du = dx12[:, :Hd], dx12[:, Hd:]
    else:
        v = dx12.view(M, Hd // 128, 2, 128)
        dg, du = v[:, :, 0].reshape(M, Hd), v[:, :, 1].reshape(M, Hd)
    assert ((dg.float() - gr.grad).norm() / gr.grad.norm()).item() < 1e-2
    assert ((du.float() - ur.grad).norm() / ur.grad.norm()).item() < 1e-2


def test_adamw_matches_torch(dev):
    from clipself_b200 import ops
    torch.manual_seed(3)
    n = 10007
    p = torch.randn(n, device=dev)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 4):
        g = torch.randn(n, device=dev)
        ref.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.1, step)
        torch.testing.assert_close(p, ref.detach(), rtol=2e-6, atol=2e-7)
