"""GPU parity of every C-ABI operator against plain PyTorch fp32 math on the same (bf16-rounded) inputs.
All calls go through libmic_b200.so (ops.py is only pointer plumbing)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import ops  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16


def rnd(*shape, scale=1.0, seed=0, dtype=BF16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def close(got, want, atol, rtol, what=""):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = (err > tol)
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {float(err.max()):.4g}, " \
                          f"first bad idx {bad.nonzero()[:4].tolist()}"


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["kn", "kk", "mm"])
@pytest.mark.parametrize("shape", [(304, 520, 200), (1024, 768, 768), (128, 256, 64), (72, 1000, 136),
                                   (2048, 1024, 4096)])
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_gemm_layouts(layout, shape, bn):
    M, N, K = shape
    a = rnd(M, K, seed=1)
    b = rnd(N, K, seed=2)
    want = a.float() @ b.float().t()
    a_mn = layout == "mm"
    b_mn = layout in ("kn", "mm")
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    out = ops.gemm(A, Bm, a_mn=a_mn, b_mn=b_mn, block_n=bn)
    torch.cuda.synchronize()
    close(out, want, atol=0.05 * math.sqrt(K / 64), rtol=1e-2, what=f"gemm {layout} {shape} bn={bn}")


def test_gemm_epilogue_bias_act_residual_preact():
    M, N, K = 384, 512, 256
    a, b = rnd(M, K, seed=3, scale=0.5), rnd(K, N, seed=4, scale=0.1)
    bias = rnd(N, seed=5, dtype=torch.float32)
    res = rnd(M, N, seed=6)
    for act, fn in (("gelu", lambda x: torch.nn.functional.gelu(x)), ("quick_gelu", lambda x: x * torch.sigmoid(1.702 * x)),
                    ("none", lambda x: x)):
        pre = torch.empty(M, N, dtype=BF16, device=DEV)
        out = ops.gemm(a, b, b_mn=True, bias=bias, act=act, pre_act_out=pre, residual=res)
        torch.cuda.synchronize()
        u = a.float() @ b.float() + bias
        close(pre, u, 2e-2, 1e-2, f"pre-act {act}")
        close(out, fn(u) + res.float(), 3e-2, 1e-2, f"epilogue {act}")
    # in-place residual (x += f(x)) as the layer code does
    x = res.clone()
    ops.gemm(a, b, b_mn=True, bias=bias, residual=x, out=x)
    torch.cuda.synchronize()
    close(x, a.float() @ b.float() + bias + res.float(), 3e-2, 1e-2, "in-place residual")


def test_gemm_fp32_out_accumulate_and_strided_views():
    M, N, K = 256, 384, 512
    # wgrad shape: dW[K_in, N] = X^T dY, operands are token-major => both MN-major
    x, dy = rnd(K, M, seed=7), rnd(K, N, seed=8)       # [tokens, feat]
    want = x.float().t() @ dy.float()
    buf = torch.zeros(M, N + 128, dtype=torch.float32, device=DEV)   # strided destination (fused-param view)
    out = buf[:, 64:64 + N]
    ops.gemm(x, dy, a_mn=True, b_mn=True, out=out)
    ops.gemm(x, dy, a_mn=True, b_mn=True, out=out, accumulate=True)
    torch.cuda.synchronize()
    close(out, 2 * want, 0.2, 1e-2, "fp32 accumulate")
    assert float(buf[:, :64].abs().max()) == 0 and float(buf[:, 64 + N:].abs().max()) == 0
    # operand that is a column slice of a wider buffer (q | k | v fused activations)
    wide = rnd(300, 3 * 128, seed=9)
    w = rnd(128, 256, seed=10, scale=0.1)
    got = ops.gemm(wide[:, 128:256], w, b_mn=True)
    torch.cuda.synchronize()
    close(got, wide[:, 128:256].float() @ w.float(), 0.05, 1e-2, "strided A view")


# ------------------------------------------------------------------------------------------------
def _ce_ws(M, V):
    n = ops.lm_head_num_partials(V)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=DEV)
    return {"nparts": n, "pmax": f(n, M), "psum": f(n, M), "psumz": f(n, M), "zlabel": f(M), "lse": f(M),
            "row_loss": f(M), "row_w": f(M), "out": f(2)}


@pytest.mark.parametrize("V,M,eps", [(1003, 40, 0.0), (1003, 200, 0.1), (5000, 300, 0.0)])
def test_lm_head_ce_forward_backward(V, M, eps):
    K = 128
    h = rnd(M, K, seed=11)
    E = rnd(V, K, seed=12, scale=0.2)
    bias = rnd(V, seed=13, dtype=torch.float32, scale=0.1)
    g = torch.Generator().manual_seed(14)
    labels = torch.randint(0, V, (M,), generator=g, dtype=torch.int32).to(DEV)
    mask = (torch.rand(M, generator=g) > 0.3).to(torch.int32).to(DEV)
    mask[0] = 1
    ws = _ce_ws(M, V)
    ops.lm_head_ce_stats(h, E, bias, labels, ws)
    ops.ce_finalize(ws, mask, M, V, eps)
    Vp = (V + 255) // 256 * 256
    dl = torch.full((M, Vp), 7.0, dtype=BF16, device=DEV)
    conf, low = 1.0 - eps, eps / (V - 1)
    ops.lm_head_ce_grad(h, E, bias, labels, ws, conf, low, dl)
    torch.cuda.synchronize()
    z = (h.float() @ E.float().t() + bias).requires_grad_(True)
    lsm = torch.log_softmax(z, -1)
    soft = torch.full_like(z, low).scatter_(1, labels.long()[:, None], conf)
    const = 0.0 if eps == 0 else -(conf * math.log(conf) + (V - 1) * low * math.log(low + 1e-20))
    row = -(soft * lsm).sum(-1) - const
    loss = (row * mask.float()).sum() / mask.float().sum()
    loss.backward()
    close(ws["lse"], torch.logsumexp(z.detach(), -1), 1e-4, 1e-5, "lse")
    close(ws["row_loss"], row.detach(), 2e-3, 1e-4, "row loss")
    assert abs(float(ws["out"][0]) - float(loss)) < 1e-3
    assert float(ws["out"][1]) == float(mask.sum())
    close(dl[:, :V], z.grad, 2e-5, 2e-2, "dlogits")
    assert float(dl[:, V:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,d", [(37, 128), (1000, 768), (513, 1024)])
def test_layernorm_fwd_bwd(M, d):
    x = rnd(M, d, seed=20, scale=2.0)
    gamma = (1 + 0.1 * torch.randn(d)).to(DEV)
    beta = (0.1 * torch.randn(d)).to(DEV)
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    y = ops.layernorm_fwd(x, gamma, beta, 1e-5, mean=mean, rstd=rstd)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    want = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    torch.cuda.synchronize()
    close(y, want, 2e-2, 1e-2, "ln fwd")
    dy = rnd(M, d, seed=21)
    dres = rnd(M, d, seed=22)
    dx = torch.empty_like(x)
    dg, db = torch.empty(d, device=DEV), torch.empty(d, device=DEV)
    ws = torch.empty(ops.ln_bwd_workspace_floats(M, d), device=DEV)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dg, db, ws)
    want.backward(dy.float())
    torch.cuda.synchronize()
    close(dx, xr.grad + dres.float(), 3e-2, 2e-2, "ln dx")
    close(dg, gr.grad, 0.05 * math.sqrt(M / 37), 1e-2, "ln dgamma")
    close(db, br.grad, 0.05 * math.sqrt(M / 37), 1e-2, "ln dbeta")


@pytest.mark.parametrize("act", ["gelu", "quick_gelu", "none"])
def test_act_bwd_colsum(act):
    M, N = 777, 1032
    dy, u = rnd(M, N, seed=30), rnd(M, N, seed=31, scale=2.0)
    du = torch.empty_like(dy) if act != "none" else None
    dbias = torch.ones(N, device=DEV)
    ws = torch.empty(ops.colsum_workspace_floats(M, N), device=DEV)
    ops.act_bwd_colsum(dy, u if act != "none" else None, act, du, dbias, ws, accumulate=True)
    torch.cuda.synchronize()
    uf = u.float().requires_grad_(True)
    f = {"gelu": torch.nn.functional.gelu, "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "none": lambda x: x}[act]
    f(uf).backward(dy.float())
    g = uf.grad if act != "none" else dy.float()
    if du is not None:
        close(du, g, 1e-2, 1e-2, "dU")
    close(dbias, 1 + g.to(BF16).float().sum(0), 0.3, 1e-2, "dbias")


def test_embed_fwd_bwd():
    V, d, B, T = 1003, 128, 5, 16
    table, pos = rnd(V, d, seed=40, scale=0.05), rnd(T + 2, d, seed=41, scale=0.05)
    gamma = (1 + 0.1 * torch.randn(d)).to(DEV)
    beta = (0.1 * torch.randn(d)).to(DEV)
    ids = torch.randint(0, V, (B * T,), dtype=torch.int32).to(DEV)
    ids[::3] = 1
    emb = torch.empty(B * T, d, dtype=BF16, device=DEV)
    y = torch.empty_like(emb)
    scale = math.sqrt(d)
    ops.embed_ln_fwd(ids, None, T, 2, table, pos, scale, gamma, beta, 1e-6, emb, y)
    torch.cuda.synchronize()
    posi = (torch.arange(B * T, device=DEV) % T) + 2
    e = table.float()[ids.long()] * scale + pos.float()[posi]
    close(emb, e, 1e-2, 1e-2, "emb")
    close(y, torch.nn.functional.layer_norm(emb.float(), (d,), gamma, beta, 1e-6), 2e-2, 1e-2, "embed ln")
    d_emb = rnd(B * T, d, seed=42)
    d_table = torch.zeros(V, d, device=DEV)
    d_pos = torch.zeros(T + 2, d, device=DEV)
    ops.embed_bwd(ids, d_emb, scale, d_table, d_pos[2:], B, T, hot_id=1)
    torch.cuda.synchronize()
    want = torch.zeros(V, d, device=DEV).index_add_(0, ids.long(), d_emb.float() * scale)
    close(d_table, want, 1e-3, 1e-4, "d_table")
    close(d_pos[2:], d_emb.float().view(B, T, d).sum(0), 1e-3, 1e-4, "d_pos")


@pytest.mark.parametrize("nchw,trunc", [(False, False), (True, False), (False, True)])
def test_patchify_and_vit_embed(nchw, trunc):
    B, img, p, d = 3, 64, 32, 128
    g = img // p
    px = torch.randn(B, img, img, 3, device=DEV) * 2
    src = px.permute(0, 3, 1, 2).contiguous() if nchw else px
    out = torch.empty(B * g * g, p * p * 3, dtype=BF16, device=DEV)
    ops.patchify(src, out, B, img, p, channel_first=nchw, trunc_int=trunc)
    torch.cuda.synchronize()
    x = torch.trunc(px) if trunc else px
    want = x.reshape(B, g, p, g, p, 3).permute(0, 1, 3, 2, 4, 5).reshape(B * g * g, p * p * 3)
    close(out, want, 1e-2, 1e-2, "patchify")
    S = g * g + 1
    po, cls, pos = rnd(B * g * g, d, seed=50), rnd(d, seed=51), rnd(S, d, seed=52)
    gamma, beta = (1 + 0.1 * torch.randn(d)).to(DEV), (0.1 * torch.randn(d)).to(DEV)
    emb = torch.empty(B * S, d, dtype=BF16, device=DEV)
    y = torch.empty_like(emb)
    mean, rstd = torch.empty(B * S, device=DEV), torch.empty(B * S, device=DEV)
    ops.vit_embed_ln_fwd(po, None, cls, pos, gamma, beta, 1e-5, True, emb, y, mean, rstd, B, S)
    torch.cuda.synchronize()
    e = torch.cat([cls.float().expand(B, 1, d), po.float().view(B, g * g, d)], 1) + pos.float()[None]
    close(emb, e.reshape(B * S, d), 2e-2, 1e-2, "vit emb")
    close(y, torch.nn.functional.layer_norm(emb.float(), (d,), gamma, beta, 1e-5), 3e-2, 1e-2, "vit pre-ln")
    dd = torch.empty(B * (S - 1), d, dtype=BF16, device=DEV)
    ops.drop_cls_rows(emb, dd, B, S)
    torch.cuda.synchronize()
    assert torch.equal(dd, emb.view(B, S, d)[:, 1:].reshape(-1, d))


def test_adamw_matches_oracle():
    from oracle import reference_model as rm
    n = 10007
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 0.1
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    shadow = torch.empty(n, dtype=BF16, device=DEV)
    pn, mn, vn = p.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy()
    for count in range(3):
        lr = rm.linear_warmup_decay_lr(count + 5, 5e-3, 10, 100)
        t = count + 1
        ops.adamw(p, m, v, g, shadow, lr, 0.9, 0.999, 1e-8, 0.01, 1 / (1 - 0.9 ** t), 1 / (1 - 0.999 ** t), 1.0)
        pn, mn, vn = rm.adamw_update(pn, g.cpu().numpy(), mn, vn, count, lr, weight_decay=0.01)
    torch.cuda.synchronize()
    np.testing.assert_allclose(p.cpu().numpy(), pn, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(v.cpu().numpy(), vn, rtol=2e-5, atol=1e-9)
    assert torch.equal(shadow, p.to(BF16))


# ------------------------------------------------------------------------------------------------
def _ref_attn(q, k, v, key_mask, causal, scale):
    B, Tq, H, hd = q.shape
    Tk = k.shape[1]
    w = torch.einsum("bqhd,bkhd->bhqk", q * scale, k)
    allow = torch.ones(B, 1, Tq, Tk, dtype=torch.bool, device=q.device)
    if causal:
        allow = allow & torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device=q.device))[None, None]
    if key_mask is not None:
        allow = allow & (key_mask[:, None, None, :] > 0)
    w = w.masked_fill(~allow, float("-inf"))
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(w, -1), v)


@pytest.mark.parametrize("Tq,Tk,causal,masked", [(64, 64, True, True), (50, 50, False, False), (64, 50, False, False),
                                                 (5, 5, False, False), (16, 16, True, True), (33, 7, False, False),
                                                 (197, 197, False, False), (64, 197, False, False),
                                                 (130, 100, True, True), (82, 82, False, True), (16, 82, False, False),
                                                 (256, 256, True, True), (1, 1, False, False), (200, 17, False, True),
                                                 (17, 256, False, False), (144, 144, True, False)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_attention_fwd_bwd(Tq, Tk, causal, masked, impl):
    """impl 0: the row-tiled kernels (default; single-kernel backward); impl 1: the one-CTA-per-head kernels
    (<= 64 tokens); impl 2: row-tiled with the backward as dQ + dK/dV kernels."""
    if impl == 1 and max(Tq, Tk) > 64:
        pytest.skip("one-CTA-per-head kernels hold at most 64 tokens")
    ops.attention_impl(impl)
    try:
        _attention_fwd_bwd(Tq, Tk, causal, masked)
    finally:
        ops.attention_impl(0)


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(256, 12, 197, 197, False), (256, 16, 64, 64, True), (256, 16, 64, 50, False)])
def test_attention_full_size_matches_torch(B, H, Tq, Tk, causal):
    """BASELINE-size launches (3072 / 4096 (batch, head) pairs: every CTA slot of the grid is exercised) against torch's
    fp32 attention on the same bf16 inputs; forward and all three gradients, relative RMS error."""
    hd, d = 64, H * 64
    g = torch.Generator().manual_seed(7)
    q = (torch.randn(B * Tq, d, generator=g) * 1.2).to(BF16).to(DEV)
    kv = (torch.randn(B * Tk, 2 * d, generator=g) * 1.2).to(BF16).to(DEV)
    do = torch.randn(B * Tq, d, generator=g).to(BF16).to(DEV)
    out = torch.empty_like(q)
    lse = torch.empty(B, H, Tq, device=DEV)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    scale = 1 / math.sqrt(hd)
    ops.attention_fwd(q, kv[:, :d], kv[:, d:], out, lse, None, causal, B, H, Tq, Tk, scale)
    ops.attention_bwd(q, kv[:, :d], kv[:, d:], out, do, lse, None, causal, dq, dkv[:, :d], dkv[:, d:], B, H, Tq, Tk, scale)
    qf = q.float().view(B, Tq, H, hd).transpose(1, 2).requires_grad_(True)          # [B,H,T,hd]
    kf = kv[:, :d].float().reshape(B, Tk, H, hd).transpose(1, 2).requires_grad_(True)
    vf = kv[:, d:].float().reshape(B, Tk, H, hd).transpose(1, 2).requires_grad_(True)
    want = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, is_causal=causal, scale=scale)
    want.backward(do.float().view(B, Tq, H, hd).transpose(1, 2))
    torch.cuda.synchronize()

    def rms(a, b):
        return float((a.float() - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())
    assert rms(out.view(B, Tq, H, hd).transpose(1, 2), want.detach()) < 4e-3
    assert rms(dq.view(B, Tq, H, hd).transpose(1, 2), qf.grad) < 8e-3
    assert rms(dkv[:, :d].reshape(B, Tk, H, hd).transpose(1, 2), kf.grad) < 8e-3
    assert rms(dkv[:, d:].reshape(B, Tk, H, hd).transpose(1, 2), vf.grad) < 8e-3
    want_lse = torch.logsumexp(torch.einsum("bhqd,bhkd->bhqk", qf.detach() * scale, kf.detach()).masked_fill(
        ~torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device=DEV)) if causal else torch.zeros(Tq, Tk, dtype=torch.bool, device=DEV),
        float("-inf")), -1)
    assert float((lse - want_lse).abs().max()) < 1e-4


def test_attention_fully_masked_rows_are_zero():
    """A query whose keys are all padded gets a zero output and zero gradients (softmax over nothing) in both kernels."""
    B, H, hd, T = 2, 2, 64, 48
    d = H * hd
    for impl in (0, 1, 2):
        ops.attention_impl(impl)
        try:
            qkv = rnd(B * T, 3 * d, seed=64)
            km = torch.ones(B, T, dtype=torch.int32)
            km[1, :] = 0
            km = km.to(DEV)
            out = torch.full((B * T, d), 7.0, dtype=BF16, device=DEV)
            lse = torch.empty(B, H, T, device=DEV)
            ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, lse, km, False, B, H, T, T, 0.125)
            dq = torch.full((B * T, d), 7.0, dtype=BF16, device=DEV)
            dkv = torch.full((B * T, 2 * d), 7.0, dtype=BF16, device=DEV)
            ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, rnd(B * T, d, seed=65), lse, km, False, dq,
                              dkv[:, :d], dkv[:, d:], B, H, T, T, 0.125)
            torch.cuda.synchronize()
            assert float(out[T:].float().abs().max()) == 0.0
            assert float(dq[T:].float().abs().max()) == 0.0 and float(dkv[T:].float().abs().max()) == 0.0
            assert torch.isfinite(out.float()).all() and torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
        finally:
            ops.attention_impl(0)


def _attention_fwd_bwd(Tq, Tk, causal, masked):
    B, H, hd = 3, 2, 64
    d = H * hd
    qkv = rnd(B * Tq, 3 * d, seed=60) if Tq == Tk else None
    if qkv is not None:
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q = rnd(B * Tq, d, seed=61)
        kv = rnd(B * Tk, 2 * d, seed=62)
        k, v = kv[:, :d], kv[:, d:]
    key_mask = None
    if masked:
        km = torch.ones(B, Tk, dtype=torch.int32)
        km[0, Tk // 2:] = 0
        km[1, 3] = 0
        key_mask = km.to(DEV)
    out = torch.empty(B * Tq, d, dtype=BF16, device=DEV)
    lse = torch.empty(B, H, Tq, device=DEV)
    scale = 1 / math.sqrt(hd)
    ops.attention_fwd(q, k, v, out, lse, key_mask, causal, B, H, Tq, Tk, scale)
    qf = q.float().reshape(B, Tq, H, hd).requires_grad_(True)
    kf = k.float().reshape(B, Tk, H, hd).requires_grad_(True)
    vf = v.float().reshape(B, Tk, H, hd).requires_grad_(True)
    want = _ref_attn(qf, kf, vf, key_mask, causal, scale)
    torch.cuda.synchronize()
    close(out.view(B, Tq, H, hd), want, 2e-2, 2e-2, "attn fwd")
    do = rnd(B * Tq, d, seed=63)
    dq = torch.empty(B * Tq, d, dtype=BF16, device=DEV)
    dkv = torch.empty(B * Tk, 2 * d, dtype=BF16, device=DEV)
    ops.attention_bwd(q, k, v, out, do, lse, key_mask, causal, dq, dkv[:, :d], dkv[:, d:], B, H, Tq, Tk, scale)
    want.backward(do.float().view(B, Tq, H, hd))
    torch.cuda.synchronize()
    close(dq.view(B, Tq, H, hd), qf.grad, 3e-2, 3e-2, "dq")
    close(dkv[:, :d].reshape(B, Tk, H, hd), kf.grad, 3e-2, 3e-2, "dk")
    close(dkv[:, d:].reshape(B, Tk, H, hd), vf.grad, 3e-2, 3e-2, "dv")


def test_decode_attention_with_ancestors_and_cross():
    R, H, hd, T = 8, 2, 64, 16
    d = H * hd
    q = rnd(R, d, seed=70)
    cache = rnd(R * T, 2 * d, seed=71)                   # [row, pos, (k|v)]
    anc = torch.randint(0, R, (R, T), dtype=torch.int32).to(DEV)
    n = 11
    out = torch.empty(R, d, dtype=BF16, device=DEV)
    scale = 1 / math.sqrt(hd)
    ops.decode_attention(q, cache[:, :d], cache[:, d:], 2 * d, anc, T, n, 1, out, R, H, scale)
    torch.cuda.synchronize()
    c = cache.float().view(R, T, 2, H, hd)
    rows = anc.long()[:, :n]
    pos = torch.arange(n, device=DEV)[None].expand(R, n)
    kk, vv = c[rows, pos, 0], c[rows, pos, 1]            # [R, n, H, hd]
    w = torch.softmax(torch.einsum("rhd,rnhd->rhn", q.float().view(R, H, hd) * scale, kk), -1)
    want = torch.einsum("rhn,rnhd->rhd", w, vv).reshape(R, d)
    close(out, want, 2e-2, 2e-2, "decode self-attn")
    # cross: 2 beams per image, S = 5 visual tokens
    S, beams = 5, 2
    enc = rnd((R // beams) * S, 2 * d, seed=72)
    ops.decode_attention(q, enc[:, :d], enc[:, d:], 2 * d, None, S, S, beams, out, R, H, scale)
    torch.cuda.synchronize()
    e = enc.float().view(R // beams, S, 2, H, hd).repeat_interleave(beams, 0)
    w = torch.softmax(torch.einsum("rhd,rnhd->rhn", q.float().view(R, H, hd) * scale, e[:, :, 0]), -1)
    close(out, torch.einsum("rhn,rnhd->rhd", w, e[:, :, 1]).reshape(R, d), 2e-2, 2e-2, "decode cross-attn")


# ------------------------------------------------------------------------------------------------
def test_lm_head_search_and_merge():
    from oracle import reference_generate as rg
    V, K, R = 5003, 128, 24
    h, E = rnd(R, K, seed=80), rnd(V, K, seed=81, scale=0.3)
    bias = rnd(V, seed=82, dtype=torch.float32, scale=0.1)
    n = ops.lm_head_search_num_partials(R)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=DEV)
    ws = {"nparts": n, "pmax": f(n, R), "psum": f(n, R), "cand_val": f(n, R, 8),
          "cand_idx": torch.empty(n, R, 8, dtype=torch.int32, device=DEV), "row_lp": f(R, 8),
          "row_tok": torch.empty(R, 8, dtype=torch.int32, device=DEV), "row_ml": f(R, 2)}
    ops.lm_head_search(h, E, bias, 2, ws)
    ops.search_merge(ws, R)
    torch.cuda.synchronize()
    z = (h.float() @ E.float().t() + bias)
    z[:, 2] = float("-inf")
    lp = rg.log_softmax(z.cpu().numpy())
    val, idx = rg.top_k(lp, 8)
    got_tok = ws["row_tok"].cpu().numpy()
    got_lp = ws["row_lp"].cpu().numpy()
    # tensor-core fp32 accumulation order differs from torch: compare tokens where the gap is clear
    gap = np.abs(np.diff(np.concatenate([val, np.sort(lp, -1)[:, -9:-8]], 1), axis=1))
    clear = np.minimum.accumulate(gap > 1e-4, axis=1)
    assert (got_tok[clear] == idx[clear]).all()
    assert clear.mean() > 0.9
    np.testing.assert_allclose(got_lp[clear], val[clear], atol=2e-4)


@pytest.mark.parametrize("R,V", [(24, 5003), (200, 70001)])
def test_lm_head_search_packed_equals_unpacked(R, V):
    """Tile-image operands + bulk copies (mic_lm_head_search_packed, mic_pack_kmajor_tiles) give bit-identical
    candidates and log-probs to the TMA-box kernel; large vocabulary exercises the two-phase top-8 queue path."""
    K = 256
    h, E = rnd(R, K, seed=83), rnd(V, K, seed=84, scale=0.3)
    bias = rnd(V, seed=85, dtype=torch.float32, scale=0.1)
    n = ops.lm_head_search_num_partials(R)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=DEV)
    def ws_():
        return {"nparts": n, "pmax": f(n, R), "psum": f(n, R), "cand_val": f(n, R, 8),
                "cand_idx": torch.empty(n, R, 8, dtype=torch.int32, device=DEV), "row_lp": f(R, 8),
                "row_tok": torch.empty(R, 8, dtype=torch.int32, device=DEV), "row_ml": f(R, 2)}
    def aligned(nbytes):
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=DEV)
        off = (-raw.data_ptr()) % 1024
        return raw[off:off + nbytes]
    a, b = ws_(), ws_()
    ops.lm_head_search(h, E, bias, 7, a)
    ops.search_merge(a, R)
    ht, et = aligned(ops.pack_kmajor_tiles_bytes(R, K, 128)), aligned(ops.pack_kmajor_tiles_bytes(V, K, 256))
    ops.pack_kmajor_tiles(h, 128, ht)
    ops.pack_kmajor_tiles(E, 256, et)
    ops.lm_head_search_packed(ht, et, bias, 7, R, V, K, b)
    ops.search_merge(b, R)
    torch.cuda.synchronize()
    assert torch.equal(a["row_tok"], b["row_tok"])
    assert torch.equal(a["row_lp"], b["row_lp"])
    # and against torch on the big case: the 8 candidates are the fp32 top-8 wherever the 8th/9th gap is clear
    z = h.float() @ E.float().t() + bias
    z[:, 7] = float("-inf")
    lp = torch.log_softmax(z, -1)
    val, idx = torch.sort(lp, dim=-1, descending=True, stable=True)
    clear = (val[:, 7] - val[:, 8]) > 1e-3
    assert clear.float().mean() > 0.5
    got = torch.sort(b["row_tok"][clear].long(), dim=-1).values
    want = torch.sort(idx[clear][:, :8], dim=-1).values
    assert torch.equal(got, want)


@pytest.mark.parametrize("split_k", [0, 3, 8])
def test_gemm_split_k_reduce_add(split_k):
    """wgrad-shaped GEMM (few output tiles, long K) with K slices combined by TMA reduce-add."""
    tokens, din, dout = 4096, 256, 384
    x, dy = rnd(tokens, din, seed=90), rnd(tokens, dout, seed=91)
    out = torch.full((din, dout), 3.0, dtype=torch.float32, device=DEV)      # must be overwritten, not added to
    ops.gemm(x, dy, a_mn=True, b_mn=True, out=out, split_k=split_k)
    torch.cuda.synchronize()
    want = x.float().t() @ dy.float()
    close(out, want, 0.5, 1e-2, f"split_k={split_k}")
    ops.gemm(x, dy, a_mn=True, b_mn=True, out=out, split_k=split_k, accumulate=True)
    torch.cuda.synchronize()
    close(out, 2 * want, 1.0, 1e-2, f"split_k={split_k} accumulate")


def test_dropout_mask_is_shared_by_forward_and_backward():
    M, N, K, p = 512, 256, 128, 0.1
    a, w = rnd(M, K, seed=95, scale=0.5), rnd(K, N, seed=96, scale=0.2)
    res = rnd(M, N, seed=97)
    seed = torch.tensor([1234], dtype=torch.int32, device=DEV)
    drop = (seed, 77, p)
    out = ops.gemm(a, w, b_mn=True, residual=res, dropout=drop)
    plain = ops.gemm(a, w, b_mn=True)
    torch.cuda.synchronize()
    z = out.float() - res.float()
    kept = z.abs() > 1e-3 * plain.float().abs().clamp_min(1e-3)
    frac = 1.0 - kept.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    scale = 65536.0 / (65536.0 - round(p * 65536))
    close(z[kept], plain.float()[kept] * scale, 3e-2, 2e-2, "kept values scaled by 1/(1-p)")
    # backward: dz = dy * mask * scale with the SAME mask
    dy = rnd(M, N, seed=98)
    du = torch.empty_like(dy)
    dbias = torch.empty(N, device=DEV)
    ws = torch.empty(ops.colsum_workspace_floats(M, N), device=DEV)
    ops.act_bwd_colsum(dy, None, "none", du, dbias, ws, dropout=drop)
    torch.cuda.synchronize()
    want = torch.where(kept, dy.float() * scale, torch.zeros_like(dy.float()))
    strong = plain.float().abs() > 0.05                      # ignore entries whose forward value was ~0 (mask unobservable)
    close(du.float()[strong], want[strong], 1e-2, 1e-2, "masked dy")
    # a different site or seed gives a different mask
    out2 = ops.gemm(a, w, b_mn=True, residual=res, dropout=(seed, 78, p))
    torch.cuda.synchronize()
    assert (out2 != out).float().mean().item() > 0.1


@pytest.mark.parametrize("V,M,eps", [(1003, 200, 0.1), (5000, 300, 0.0)])
def test_ce_backward_from_stored_logits(V, M, eps):
    """Forward stores bf16 logits; backward rewrites them in place as dlogits and emits the bias gradient."""
    K = 128
    h = rnd(M, K, seed=11)
    E = rnd(V, K, seed=12, scale=0.2)
    bias = rnd(V, seed=13, dtype=torch.float32, scale=0.1)
    g = torch.Generator().manual_seed(14)
    labels = torch.randint(0, V, (M,), generator=g, dtype=torch.int32).to(DEV)
    mask = (torch.rand(M, generator=g) > 0.3).to(torch.int32).to(DEV)
    mask[0] = 1
    ws = _ce_ws(M, V)
    Vp = (V + 255) // 256 * 256
    buf = torch.full((M, Vp), float("nan"), dtype=BF16, device=DEV)
    ops.lm_head_ce_stats(h, E, bias, labels, ws, logits_out=buf)
    ops.ce_finalize(ws, mask, M, V, eps)
    torch.cuda.synchronize()
    z = (h.float() @ E.float().t() + bias)
    close(buf[:, :V], z, 3e-2, 1e-2, "stored logits")
    conf, low = 1.0 - eps, eps / (V - 1)
    dbias = torch.full((V,), 9.0, device=DEV)
    wsp = torch.zeros(ops.ce_softmax_bwd_workspace_floats(M, Vp), device=DEV)
    ops.ce_softmax_bwd(buf, labels, ws, conf, low, V, dbias, wsp)
    torch.cuda.synchronize()
    zz = z.clone().requires_grad_(True)
    lsm = torch.log_softmax(zz, -1)
    soft = torch.full_like(zz, low).scatter_(1, labels.long()[:, None], conf)
    loss = ((-(soft * lsm).sum(-1)) * mask.float()).sum() / mask.float().sum()
    loss.backward()
    close(buf[:, :V], zz.grad, 3e-5, 6e-2, "dlogits from stored logits")
    assert float(buf[:, V:].float().abs().max()) == 0.0
    close(dbias, buf[:, :V].float().sum(0), 2e-4, 1e-3, "final_logits_bias gradient")
