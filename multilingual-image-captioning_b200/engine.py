"""Execution engine: sequences the sm_100a kernels (ops.py -> libmic_b200.so) for the training
forward/backward of the captioning model and for the cached decode step.

Reference call stack replaced: `train_step` main.py:684-707 -> `__call__`
modeling_clip_vision_mbart.py:447-510 -> FlaxCLIPVisionModule / visual_projection / FlaxMBartDecoder /
tied lm_head (:79-102,170-178) -> `loss_fn` main.py:658-680 -> `jax.value_and_grad` main.py:696.
Backward is a hand-scheduled tape (no autograd): each Dense = dgrad GEMM + wgrad GEMM + bias column-sum.

All activations are bf16 [tokens, features]; statistics and gradients of parameters are fp32.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .params import ParamStore

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


class _Bufs:
    """Lazily allocated device buffers, addressed by name.  `t[name]` is the buffer last requested under that name;
    every (name, shape, dtype) ever requested stays allocated at a stable address, because captured CUDA graphs
    (train step, generate loop) hold raw pointers to them: re-requesting a name with another shape must not free
    the tensor an older graph still replays on."""

    def __init__(self, device):
        self.device = device
        self.t = {}
        self.pool = {}

    def get(self, name, shape, dtype=BF16):
        shape = tuple(shape)
        t = self.t.get(name)
        if t is not None and tuple(t.shape) == shape and t.dtype == dtype:
            return t
        key = (name, shape, dtype)
        t = self.pool.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.pool[key] = t
        self.t[name] = t
        return t

    def zeros(self, name, shape, dtype=F32):
        """Like get(), but zero-filled when first allocated (accumulators that their kernels hand back zeroed)."""
        key = (name, tuple(shape), dtype)
        if key not in self.pool:
            self.pool[key] = torch.zeros(tuple(shape), dtype=dtype, device=self.device)
        self.t[name] = self.pool[key]
        return self.pool[key]

    def nbytes(self):
        return sum(x.numel() * x.element_size() for x in self.pool.values())


class CaptionEngine:
    def __init__(self, config, store: ParamStore):
        self.cfg = config
        self.c = config.clip_vision_config
        self.t = config.mbart_config
        self.ps = store
        self.dev = store.device
        assert self.c.head_dim == 64 and self.t.head_dim == 64, "attention kernels are specialised for head_dim 64"
        self.bufs = _Bufs(self.dev)
        self.Vp = (self.t.vocab_size + 255) // 256 * 256
        self.emb_scale = math.sqrt(self.t.d_model) if self.t.scale_embedding else 1.0
        self._ws = None
        import os
        # fc2 dgrad GEMM applying act'(u) of fc1 in its epilogue (mic_gemm_bf16 act < 0): parity-tested, but measured
        # 1 % SLOWER than the separate activation-backward pass (74.3 vs 73.4 ms/step: the 16-warp epilogue is
        # register-starved) -> off by default
        self.fuse_act_bwd = os.environ.get("MIC_FUSE_ACT_BWD", "0") != "0"
        # decoder dropout (flax.linen.Dropout, rate = mbart_config.dropout) — active only inside train steps
        self.dropout_p = 0.0
        self.drop_seed = torch.zeros(1, dtype=I32, device=self.dev)
        # weight-gradient GEMMs run on a side stream (a parallel branch of the captured graph): their
        # prologue / last partial wave overlaps the data-gradient GEMM that follows on the main stream
        self.overlap_wgrad = True
        self._side = None
        self._side_bufs = set()      # data_ptr of every buffer outstanding side-stream work reads or writes

    def _fork_side(self, fn, *buffers):
        if not self.overlap_wgrad:
            fn()
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            fn()
        for t in buffers:
            self._side_bufs.add(t.data_ptr())

    def _before_write(self, *tensors):
        """Main-stream kernels that overwrite a buffer still used by side-stream work must wait for it."""
        if self._side_bufs and any(t is not None and t.data_ptr() in self._side_bufs for t in tensors):
            self._join_side()

    def _join_side(self):
        if self._side is not None and self._side_bufs:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_bufs.clear()

    def _drop(self, site):
        """(seed tensor, site key, p) for dropout site `site` of this step, or None when dropout is off."""
        if self.dropout_p <= 0.0:
            return None
        return (self.drop_seed, (site * 0x9E3779B9) & 0x7FFFFFFF, self.dropout_p)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, floats, side=False):
        """fp32 scratch for the column-sum kernels (one per stream).  Outgrown workspaces stay allocated: a captured
        graph may still replay on them."""
        k = "_ws_side" if side else "_ws"
        cur = getattr(self, k, None)
        if cur is None or cur.numel() < floats:
            if cur is not None:
                self.__dict__.setdefault("_ws_old", []).append(cur)
            cur = torch.empty(int(floats), dtype=F32, device=self.dev)
            setattr(self, k, cur)
        return cur

    def _ln_fwd(self, x, name, eps, out, stats=None):
        ps = self.ps
        mean, rstd = (stats if stats is not None else (None, None))
        return ops.layernorm_fwd(x, ps.f(name + ".scale"), ps.f(name + ".bias"), eps, out=out, mean=mean, rstd=rstd)

    def _ln_bwd(self, dy, x, name, stats, dres, dx):
        ps = self.ps
        d = x.shape[1]
        self._before_write(dx)
        ws = self._workspace(max(ops.ln_bwd_workspace_floats(x.shape[0], d), 1))
        ops.layernorm_bwd(dy, x, ps.f(name + ".scale"), stats[0], stats[1], dres, dx, ps.g(name + ".scale"),
                          ps.g(name + ".bias"), ws)
        return dx

    def _dense_bwd(self, x, dy, wname, dx_out, bias=True, w_view=None, gw_view=None, gb_view=None, act=None, u=None,
                   du=None, dropout=None, dgrad_residual=None, dgrad_act_bwd=None):
        """Backward of y = act(x @ W + b).  x: [M,K] input, dy: [M,N] grad of the output (post-act).
        Returns dx (written into dx_out) — or None if dx_out is None."""
        ps = self.ps
        M, N = dy.shape
        w = w_view if w_view is not None else ps.w(wname + ".w")
        gw = gw_view if gw_view is not None else ps.g(wname + ".w")
        gb = gb_view if gb_view is not None else (ps.g(wname + ".b") if bias else None)
        ws = self._workspace(ops.colsum_workspace_floats(M, N))
        if dropout is not None:
            # y = residual + dropout(x W + b): the gradient reaching this Dense is dy * mask / (1-p)
            du = self.bufs.get("tr.dmask", tuple(dy.shape))
            self._before_write(du)
            ops.act_bwd_colsum(dy, None, "none", du, gb, ws, dropout=dropout)
            dy = du
        elif act is not None and act != "none":
            self._before_write(du)
            ops.act_bwd_colsum(dy, u, act, du, gb, ws)
            dy = du
        elif gb is not None:
            # (moving this plain column sum to the side stream next to the wgrad GEMM was measured: no gain, the
            #  step is power-capped rather than stream-serialisation bound)
            ops.act_bwd_colsum(dy, None, "none", None, gb, ws)
        dyv = dy
        self._fork_side(lambda: ops.gemm(x, dyv, a_mn=True, b_mn=True, out=gw), x, dyv, gw)   # dW[K,N] = x^T dy
        if dx_out is not None:
            self._before_write(dx_out)
            # dx[M,K] = dy W^T (+ skip grad), or fused with the activation backward of the layer that produced x:
            # dgrad_act_bwd = (act, U): dx = (dy W^T) o act'(U)
            ops.gemm(dy, w, a_mn=False, b_mn=False, out=dx_out, residual=dgrad_residual, act_bwd=dgrad_act_bwd)
        return dx_out

    # ------------------------------------------------------------------------------------------
    # vision encoder + projection (forward).  save=True keeps what backward needs.
    # ------------------------------------------------------------------------------------------
    def encode(self, pixel_values, trunc_int=False, save=False, tag="enc"):
        c, ps, b = self.c, self.ps, self.bufs
        B = pixel_values.shape[0]
        S, npatch, dv = c.num_tokens, c.num_patches, c.hidden_size
        Mv = B * S
        # uint8 pixels = the input hand-off (x/255 and Normalize fused into the patch kernel); anything else is the
        # reference's contract: normalised float pixels (cast to f32, modeling_clip_vision_mbart.py:501)
        px = pixel_values.to(self.dev).contiguous() if pixel_values.dtype == torch.uint8 else \
            pixel_values.to(self.dev, F32).contiguous()
        patches = ops.patchify(px, b.get(tag + ".patches", (B * npatch, c.patch_size ** 2 * 3)), B, c.image_size,
                               c.patch_size, channel_first=c.channel_first_input, trunc_int=trunc_int,
                               mean=c.image_mean, std=c.image_std)
        patch_out = ops.gemm(patches, ps.w("v.patch.w"), b_mn=True, out=b.get(tag + ".patch_out", (B * npatch, dv)))
        emb = b.get(tag + ".emb", (Mv, dv))
        st_pre = (b.get(tag + ".pre.mean", (Mv,), F32), b.get(tag + ".pre.rstd", (Mv,), F32))
        x = b.get(tag + ".x0", (Mv, dv))
        ops.vit_embed_ln_fwd(patch_out, ps.f("v.patch.b") if c.patch_bias else None, ps.w("v.cls"), ps.w("v.pos"),
                             ps.f("v.pre_ln.scale"), ps.f("v.pre_ln.bias"), c.layer_norm_eps, c.pre_layernorm, emb, x,
                             st_pre[0], st_pre[1], B, S)
        H = c.num_attention_heads
        scale = 1.0 / math.sqrt(c.head_dim)
        for l in range(c.num_hidden_layers):
            n = f"v.{l}"
            sfx = f".{l}" if save else ""
            st1 = (b.get(tag + ".ln1.mean" + sfx, (Mv,), F32), b.get(tag + ".ln1.rstd" + sfx, (Mv,), F32))
            a = self._ln_fwd(x, n + ".ln1", c.layer_norm_eps, b.get(tag + ".ln1" + sfx, (Mv, dv)), st1)
            qkv = ops.gemm(a, ps.w(n + ".qkv.w"), b_mn=True, bias=ps.f(n + ".qkv.b"), out=b.get(tag + ".qkv" + sfx, (Mv, 3 * dv)))
            att = b.get(tag + ".att" + sfx, (Mv, dv))
            lse = b.get(tag + ".lse" + sfx, (B, H, S), F32)
            ops.attention_fwd(qkv[:, :dv], qkv[:, dv:2 * dv], qkv[:, 2 * dv:], att, lse, None, False, B, H, S, S, scale)
            xm = ops.gemm(att, ps.w(n + ".o.w"), b_mn=True, bias=ps.f(n + ".o.b"), residual=x,
                          out=b.get(tag + ".xm" + sfx, (Mv, dv)))
            st2 = (b.get(tag + ".ln2.mean" + sfx, (Mv,), F32), b.get(tag + ".ln2.rstd" + sfx, (Mv,), F32))
            m = self._ln_fwd(xm, n + ".ln2", c.layer_norm_eps, b.get(tag + ".ln2" + sfx, (Mv, dv)), st2)
            u = b.get(tag + ".u" + sfx, (Mv, c.intermediate_size)) if save else None
            g = ops.gemm(m, ps.w(n + ".fc1.w"), b_mn=True, bias=ps.f(n + ".fc1.b"), act=c.hidden_act, pre_act_out=u,
                         out=b.get(tag + ".g" + sfx, (Mv, c.intermediate_size)))
            x = ops.gemm(g, ps.w(n + ".fc2.w"), b_mn=True, bias=ps.f(n + ".fc2.b"), residual=xm,
                         out=b.get(tag + f".x{l + 1}" if save else tag + f".xping{l & 1}", (Mv, dv)))
        if c.final_layernorm:
            stf = (b.get(tag + ".post.mean", (Mv,), F32), b.get(tag + ".post.rstd", (Mv,), F32))
            x = self._ln_fwd(x, "v.post_ln", c.layer_norm_eps, b.get(tag + ".post", (Mv, dv)), stf)
        enc = ops.gemm(x, ps.w("proj.w"), b_mn=True, bias=ps.f("proj.b"), out=b.get(tag + ".out", (Mv, self.t.d_model)))
        return enc

    def cross_kv(self, enc, tag="enc"):
        """All layers' cross-attention K/V projections of the visual tokens in one GEMM: [Mv, L*2d]."""
        ps, t = self.ps, self.t
        return ops.gemm(enc, ps.w("d.ca_kv.w"), b_mn=True, bias=ps.f("d.ca_kv.b"),
                        out=self.bufs.get(tag + ".kv", (enc.shape[0], t.decoder_layers * 2 * t.d_model)))

    # ------------------------------------------------------------------------------------------
    # full-sequence decoder forward (training / eval), returns final hidden states [B*T, d]
    # ------------------------------------------------------------------------------------------
    def decoder_forward(self, ids, key_mask, pos_ids, enc_kv, B, T, S, save=False, tag="dec", train=False):
        t, ps, b = self.t, self.ps, self.bufs
        d, M, H = t.d_model, B * T, t.decoder_attention_heads
        if not t.pre_layernorm:
            return self._decoder_forward_postln(ids, key_mask, pos_ids, enc_kv, B, T, S, save, tag, train)
        eps = t.layer_norm_eps
        scale = 1.0 / math.sqrt(t.head_dim)
        emb = b.get(tag + ".emb", (M, d))
        st = (b.get(tag + ".emb.mean", (M,), F32), b.get(tag + ".emb.rstd", (M,), F32))
        x = b.get(tag + ".x0", (M, d))
        drop = self._drop if train else (lambda site: None)
        ops.embed_ln_fwd(ids, pos_ids, T, t.position_offset, ps.w("shared"), ps.w("d.pos"), self.emb_scale,
                         ps.f("d.ln_emb.scale"), ps.f("d.ln_emb.bias"), eps, emb, x, st[0], st[1], dropout=drop(1))
        for l in range(t.decoder_layers):
            n = f"d.{l}"
            sfx = f".{l}" if save else ""
            stA = (b.get(tag + ".lnA.mean" + sfx, (M,), F32), b.get(tag + ".lnA.rstd" + sfx, (M,), F32))
            a = self._ln_fwd(x, n + ".ln_sa", eps, b.get(tag + ".lnA" + sfx, (M, d)), stA)
            qkv = ops.gemm(a, ps.w(n + ".sa_qkv.w"), b_mn=True, bias=ps.f(n + ".sa_qkv.b"),
                           out=b.get(tag + ".qkv" + sfx, (M, 3 * d)))
            sa = b.get(tag + ".sa" + sfx, (M, d))
            lse1 = b.get(tag + ".lse1" + sfx, (B, H, T), F32)
            ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], sa, lse1, key_mask, True, B, H, T, T, scale)
            x1 = ops.gemm(sa, ps.w(n + ".sa_o.w"), b_mn=True, bias=ps.f(n + ".sa_o.b"), residual=x,
                          out=b.get(tag + ".x1" + sfx, (M, d)), dropout=drop(10 + 4 * l))
            stC = (b.get(tag + ".lnC.mean" + sfx, (M,), F32), b.get(tag + ".lnC.rstd" + sfx, (M,), F32))
            cc = self._ln_fwd(x1, n + ".ln_ca", eps, b.get(tag + ".lnC" + sfx, (M, d)), stC)
            qc = ops.gemm(cc, ps.w(n + ".ca_q.w"), b_mn=True, bias=ps.f(n + ".ca_q.b"), out=b.get(tag + ".qc" + sfx, (M, d)))
            ca = b.get(tag + ".ca" + sfx, (M, d))
            lse2 = b.get(tag + ".lse2" + sfx, (B, H, T), F32)
            kl = enc_kv[:, l * 2 * d: l * 2 * d + d]
            vl = enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
            ops.attention_fwd(qc, kl, vl, ca, lse2, None, False, B, H, T, S, scale)
            x2 = ops.gemm(ca, ps.w(n + ".ca_o.w"), b_mn=True, bias=ps.f(n + ".ca_o.b"), residual=x1,
                          out=b.get(tag + ".x2" + sfx, (M, d)), dropout=drop(11 + 4 * l))
            stF = (b.get(tag + ".lnF.mean" + sfx, (M,), F32), b.get(tag + ".lnF.rstd" + sfx, (M,), F32))
            f = self._ln_fwd(x2, n + ".ln_f", eps, b.get(tag + ".lnF" + sfx, (M, d)), stF)
            u = b.get(tag + ".u" + sfx, (M, t.decoder_ffn_dim)) if save else None
            g = ops.gemm(f, ps.w(n + ".fc1.w"), b_mn=True, bias=ps.f(n + ".fc1.b"), act=t.activation_function,
                         pre_act_out=u, out=b.get(tag + ".g" + sfx, (M, t.decoder_ffn_dim)))
            x = ops.gemm(g, ps.w(n + ".fc2.w"), b_mn=True, bias=ps.f(n + ".fc2.b"), residual=x2,
                         out=b.get(tag + f".x{l + 1}" if save else tag + f".xping{l & 1}", (M, d)),
                         dropout=drop(12 + 4 * l))
        if t.final_layer_norm:
            stL = (b.get(tag + ".lnL.mean", (M,), F32), b.get(tag + ".lnL.rstd", (M,), F32))
            x = self._ln_fwd(x, "d.ln_final", eps, b.get(tag + ".hf", (M, d)), stL)
        return x

    # ------------------------------------------------------------------------------------------
    # BART decoder (flax_vit_bart variant): POST-LN blocks  h = LN(h + dropout(sublayer(h))), no final LN
    # (FlaxBartDecoderLayer; HF-PT twin modeling_bart.py:354-389)
    # ------------------------------------------------------------------------------------------
    def _decoder_forward_postln(self, ids, key_mask, pos_ids, enc_kv, B, T, S, save, tag, train):
        t, ps, b = self.t, self.ps, self.bufs
        d, M, H = t.d_model, B * T, t.decoder_attention_heads
        eps = t.layer_norm_eps
        scale = 1.0 / math.sqrt(t.head_dim)
        drop = self._drop if train else (lambda site: None)
        emb = b.get(tag + ".emb", (M, d))
        st = (b.get(tag + ".emb.mean", (M,), F32), b.get(tag + ".emb.rstd", (M,), F32))
        x = b.get(tag + ".x0", (M, d))
        ops.embed_ln_fwd(ids, pos_ids, T, t.position_offset, ps.w("shared"), ps.w("d.pos"), self.emb_scale,
                         ps.f("d.ln_emb.scale"), ps.f("d.ln_emb.bias"), eps, emb, x, st[0], st[1], dropout=drop(1))
        for l in range(t.decoder_layers):
            n = f"d.{l}"
            sfx = f".{l}" if save else ""

            def stats(nm):
                return (b.get(tag + nm + ".mean" + sfx, (M,), F32), b.get(tag + nm + ".rstd" + sfx, (M,), F32))
            qkv = ops.gemm(x, ps.w(n + ".sa_qkv.w"), b_mn=True, bias=ps.f(n + ".sa_qkv.b"),
                           out=b.get(tag + ".qkv" + sfx, (M, 3 * d)))
            sa = b.get(tag + ".sa" + sfx, (M, d))
            ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], sa, b.get(tag + ".lse1" + sfx, (B, H, T), F32),
                              key_mask, True, B, H, T, T, scale)
            y1 = ops.gemm(sa, ps.w(n + ".sa_o.w"), b_mn=True, bias=ps.f(n + ".sa_o.b"), residual=x,
                          out=b.get(tag + ".y1" + sfx, (M, d)), dropout=drop(10 + 4 * l))
            x1 = self._ln_fwd(y1, n + ".ln_sa", eps, b.get(tag + ".x1" + sfx, (M, d)), stats(".lnA"))
            qc = ops.gemm(x1, ps.w(n + ".ca_q.w"), b_mn=True, bias=ps.f(n + ".ca_q.b"), out=b.get(tag + ".qc" + sfx, (M, d)))
            ca = b.get(tag + ".ca" + sfx, (M, d))
            kl = enc_kv[:, l * 2 * d: l * 2 * d + d]
            vl = enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
            ops.attention_fwd(qc, kl, vl, ca, b.get(tag + ".lse2" + sfx, (B, H, T), F32), None, False, B, H, T, S, scale)
            y2 = ops.gemm(ca, ps.w(n + ".ca_o.w"), b_mn=True, bias=ps.f(n + ".ca_o.b"), residual=x1,
                          out=b.get(tag + ".y2" + sfx, (M, d)), dropout=drop(11 + 4 * l))
            x2 = self._ln_fwd(y2, n + ".ln_ca", eps, b.get(tag + ".x2" + sfx, (M, d)), stats(".lnC"))
            u = b.get(tag + ".u" + sfx, (M, t.decoder_ffn_dim)) if save else None
            g = ops.gemm(x2, ps.w(n + ".fc1.w"), b_mn=True, bias=ps.f(n + ".fc1.b"), act=t.activation_function,
                         pre_act_out=u, out=b.get(tag + ".g" + sfx, (M, t.decoder_ffn_dim)))
            y3 = ops.gemm(g, ps.w(n + ".fc2.w"), b_mn=True, bias=ps.f(n + ".fc2.b"), residual=x2,
                          out=b.get(tag + ".y3" + sfx, (M, d)), dropout=drop(12 + 4 * l))
            x = self._ln_fwd(y3, n + ".ln_f", eps, b.get(tag + f".x{l + 1}" if save else tag + f".xping{l & 1}", (M, d)),
                             stats(".lnF"))
        if t.final_layer_norm:
            stL = (b.get(tag + ".lnL.mean", (M,), F32), b.get(tag + ".lnL.rstd", (M,), F32))
            x = self._ln_fwd(x, "d.ln_final", eps, b.get(tag + ".hf", (M, d)), stL)
        return x

    def _decoder_backward_preln(self, dx, km, enc_kv, d_enc_kv, B, T, S, tg):
        t, b = self.t, self.bufs
        d, M = t.d_model, B * T
        H = t.decoder_attention_heads
        scale = 1.0 / math.sqrt(t.head_dim)
        dg = b.get("tr.dg", (M, t.decoder_ffn_dim))
        du = b.get("tr.du", (M, t.decoder_ffn_dim))
        dtmp = b.get("tr.dtmp", (M, d))
        dqkv = b.get("tr.dqkv", (M, 3 * d))
        dqc = b.get("tr.dqc", (M, d))
        for l in reversed(range(t.decoder_layers)):
            n = f"d.{l}"
            sfx = f".{l}"
            g_, u_, lnF = b.t[tg + ".g" + sfx], b.t[tg + ".u" + sfx], b.t[tg + ".lnF" + sfx]
            x2, x1, x0 = b.t[tg + ".x2" + sfx], b.t[tg + ".x1" + sfx], b.t[tg + f".x{l}"]
            # FFN
            if self.fuse_act_bwd:      # fc2 dgrad emits d(fc1 pre-activation) directly: dg = (dx W2^T) o act'(u)
                self._dense_bwd(g_, dx, n + ".fc2", dg, dropout=self._drop(12 + 4 * l),
                                dgrad_act_bwd=(t.activation_function, u_))
                self._dense_bwd(lnF, dg, n + ".fc1", dtmp)
            else:
                self._dense_bwd(g_, dx, n + ".fc2", dg, dropout=self._drop(12 + 4 * l))
                self._dense_bwd(lnF, dg, n + ".fc1", dtmp, act=t.activation_function, u=u_, du=du)
            self._ln_bwd(dtmp, x2, n + ".ln_f", (b.t[tg + ".lnF.mean" + sfx], b.t[tg + ".lnF.rstd" + sfx]), dx, dx)
            # cross attention
            ca, qc, lnC = b.t[tg + ".ca" + sfx], b.t[tg + ".qc" + sfx], b.t[tg + ".lnC" + sfx]
            self._dense_bwd(ca, dx, n + ".ca_o", dtmp, dropout=self._drop(11 + 4 * l))
            kl = enc_kv[:, l * 2 * d: l * 2 * d + d]
            vl = enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
            self._before_write(dqc, d_enc_kv)
            ops.attention_bwd(qc, kl, vl, ca, dtmp, b.t[tg + ".lse2" + sfx], None, False, dqc,
                              d_enc_kv[:, l * 2 * d: l * 2 * d + d], d_enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d],
                              B, H, T, S, scale)
            self._dense_bwd(lnC, dqc, n + ".ca_q", dtmp)
            self._ln_bwd(dtmp, x1, n + ".ln_ca", (b.t[tg + ".lnC.mean" + sfx], b.t[tg + ".lnC.rstd" + sfx]), dx, dx)
            # self attention
            sa, qkv, lnA = b.t[tg + ".sa" + sfx], b.t[tg + ".qkv" + sfx], b.t[tg + ".lnA" + sfx]
            self._dense_bwd(sa, dx, n + ".sa_o", dtmp, dropout=self._drop(10 + 4 * l))
            self._before_write(dqkv)
            ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], sa, dtmp, b.t[tg + ".lse1" + sfx], km, True,
                              dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H, T, T, scale)
            self._dense_bwd(lnA, dqkv, n + ".sa_qkv", dtmp)
            self._ln_bwd(dtmp, x0, n + ".ln_sa", (b.t[tg + ".lnA.mean" + sfx], b.t[tg + ".lnA.rstd" + sfx]), dx, dx)
    def _decoder_backward_postln(self, dx, km, enc_kv, d_enc_kv, B, T, S, tg):
        """dx: gradient w.r.t. the last layer's output (in place buffer).  Returns the gradient w.r.t. x0."""
        t, b = self.t, self.bufs
        d, M, H = t.d_model, B * T, t.decoder_attention_heads
        scale = 1.0 / math.sqrt(t.head_dim)
        dg = b.get("tr.dg", (M, t.decoder_ffn_dim))
        du = b.get("tr.du", (M, t.decoder_ffn_dim))
        dy = b.get("tr.dy", (M, d))
        dtmp = b.get("tr.dtmp", (M, d))
        dqkv = b.get("tr.dqkv", (M, 3 * d))
        dqc = b.get("tr.dqc", (M, d))
        for l in reversed(range(t.decoder_layers)):
            n = f"d.{l}"
            sfx = f".{l}"

            def st(nm):
                return (b.t[tg + nm + ".mean" + sfx], b.t[tg + nm + ".rstd" + sfx])
            # FFN block: x3 = LN(y3), y3 = x2 + drop(fc2(gelu(fc1(x2))))
            self._ln_bwd(dx, b.t[tg + ".y3" + sfx], n + ".ln_f", st(".lnF"), None, dy)
            if self.fuse_act_bwd:
                self._dense_bwd(b.t[tg + ".g" + sfx], dy, n + ".fc2", dg, dropout=self._drop(12 + 4 * l),
                                dgrad_act_bwd=(t.activation_function, b.t[tg + ".u" + sfx]))
                self._dense_bwd(b.t[tg + ".x2" + sfx], dg, n + ".fc1", dx, dgrad_residual=dy)   # dx2 = dy3 + fc1 dgrad
            else:
                self._dense_bwd(b.t[tg + ".g" + sfx], dy, n + ".fc2", dg, dropout=self._drop(12 + 4 * l))
                self._dense_bwd(b.t[tg + ".x2" + sfx], dg, n + ".fc1", dx, act=t.activation_function,
                                u=b.t[tg + ".u" + sfx], du=du, dgrad_residual=dy)      # dx2 = dy3 + fc1 dgrad
            # cross-attention block
            self._ln_bwd(dx, b.t[tg + ".y2" + sfx], n + ".ln_ca", st(".lnC"), None, dy)
            self._dense_bwd(b.t[tg + ".ca" + sfx], dy, n + ".ca_o", dtmp, dropout=self._drop(11 + 4 * l))
            kl = enc_kv[:, l * 2 * d: l * 2 * d + d]
            vl = enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
            self._before_write(dqc, d_enc_kv)
            ops.attention_bwd(b.t[tg + ".qc" + sfx], kl, vl, b.t[tg + ".ca" + sfx], dtmp, b.t[tg + ".lse2" + sfx], None,
                              False, dqc, d_enc_kv[:, l * 2 * d: l * 2 * d + d],
                              d_enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d], B, H, T, S, scale)
            self._dense_bwd(b.t[tg + ".x1" + sfx], dqc, n + ".ca_q", dx, dgrad_residual=dy)   # dx1 = dy2 + q dgrad
            # self-attention block
            self._ln_bwd(dx, b.t[tg + ".y1" + sfx], n + ".ln_sa", st(".lnA"), None, dy)
            self._dense_bwd(b.t[tg + ".sa" + sfx], dy, n + ".sa_o", dtmp, dropout=self._drop(10 + 4 * l))
            qkv = b.t[tg + ".qkv" + sfx]
            self._before_write(dqkv)
            ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], b.t[tg + ".sa" + sfx], dtmp,
                              b.t[tg + ".lse1" + sfx], km, True, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H,
                              T, T, scale)
            self._dense_bwd(b.t[tg + f".x{l}"], dqkv, n + ".sa_qkv", dx, dgrad_residual=dy)   # dx0 = dy1 + qkv dgrad
        return dx

    # ------------------------------------------------------------------------------------------
    # loss (fused lm_head + CE) and logits
    # ------------------------------------------------------------------------------------------
    def _ce_ws(self, M):
        b, V = self.bufs, self.t.vocab_size
        n = ops.lm_head_num_partials(V)
        f = lambda nm, *s: b.get("ce." + nm, s, F32)
        return {"nparts": n, "pmax": f("pmax", n, M), "psum": f("psum", n, M), "psumz": f("psumz", n, M),
                "zlabel": f("zlabel", M), "lse": f("lse", M), "row_loss": f("row_loss", M), "row_w": f("row_w", M),
                "out": f("out", 2)}

    def loss_forward(self, hf, labels, mask, label_smoothing, logits_out=None):
        ps, V = self.ps, self.t.vocab_size
        M = hf.shape[0]
        ws = self._ce_ws(M)
        ops.lm_head_ce_stats(hf, ps.w("shared"), ps.f("flb"), labels, ws, logits_out)
        ops.ce_finalize(ws, mask, M, V, label_smoothing)
        return ws

    def logits(self, hf):
        """Materialised fp32 logits (B*T, V) for API parity with `__call__(...).logits`."""
        ps, V = self.ps, self.t.vocab_size
        out = torch.empty((hf.shape[0], V), dtype=F32, device=self.dev)
        return ops.gemm(hf, ps.w("shared"), b_mn=False, bias=ps.f("flb"), out=out)

    # ------------------------------------------------------------------------------------------
    # training forward + backward: fills ps.grad, returns the CE workspace (ws["out"][0] = loss)
    # ------------------------------------------------------------------------------------------
    def forward_backward(self, pixel_values, decoder_input_ids, attention_mask, labels, label_smoothing=0.0,
                         position_ids=None, stage=0):
        """stage 0: the whole step.  stage 1: forward + backward down to the cross-K/V projection (every
        gradient stored before `grad_split_offset()` is final afterwards: lm_head bias, tied embedding, decoder,
        cross K/V) — stage 2: visual projection + vision encoder backward.  The two-stage form lets the
        data-parallel all-reduce of the first 84 % of the gradient overlap the vision backward."""
        c, t, ps, b = self.c, self.t, self.ps, self.bufs
        ps.ensure_grad()
        B, T = decoder_input_ids.shape
        S, dv, d = c.num_tokens, c.hidden_size, t.d_model
        M, Mv = B * T, B * S
        if stage in (2, 3):
            # data parallel: the vision backward in two segments, so that only the gradients of the last
            # `dp_vision_tail_layers` encoder layers + the embeddings are all-reduced after the backward has ended
            L = c.num_hidden_layers
            k = min(self.dp_vision_tail_layers, L)
            return self._backward_vision(B, S, Mv, dv, d, part=("head", L - 1, k) if stage == 2 else ("tail", k - 1, 0))
        ids = decoder_input_ids.to(self.dev, I32).contiguous().view(-1)
        km = attention_mask.to(self.dev, I32).contiguous()
        lab = labels.to(self.dev, I32).contiguous().view(-1)
        pos = None if position_ids is None else position_ids.to(self.dev, I32).contiguous().view(-1)
        # ---------------- forward ----------------
        enc = self.encode(pixel_values, trunc_int=False, save=True, tag="tr.enc")
        enc_kv = self.cross_kv(enc, tag="tr.enc")
        hf = self.decoder_forward(ids, km, pos, enc_kv, B, T, S, save=True, tag="tr.dec", train=True)
        # the forward CE kernel also leaves the bf16 logits (what the reference's bf16 mode materialises) in
        # the buffer that backward turns into dlogits in place: no lm_head recompute
        dlog = b.get("tr.dlogits", (M, self.Vp))
        ws = self.loss_forward(hf, lab, km.view(-1), label_smoothing, logits_out=dlog)
        # ---------------- backward: lm_head + CE ----------------
        V = t.vocab_size
        conf, low = 1.0 - label_smoothing, label_smoothing / (V - 1)
        cws = self._workspace(ops.ce_softmax_bwd_workspace_floats(M, self.Vp))
        ops.ce_softmax_bwd(dlog, lab, ws, conf, low, V, ps.g("flb"), cws)
        gshared = ps.g("shared")
        dlv = dlog[:, :V]
        self._fork_side(lambda: ops.gemm(dlv, hf, a_mn=True, b_mn=True, out=gshared), dlog, hf, gshared)   # dE = dlogits^T h
        dx = b.get("tr.dx", (M, d))
        dhf = ops.gemm(dlog[:, :V], ps.w("shared"), a_mn=False, b_mn=True, out=b.get("tr.dhf", (M, d)))
        tg = "tr.dec"
        if t.final_layer_norm:
            self._ln_bwd(dhf, b.t[tg + f".x{t.decoder_layers}"], "d.ln_final",
                         (b.t[tg + ".lnL.mean"], b.t[tg + ".lnL.rstd"]), None, dx)
        else:
            dx.copy_(dhf)
            ps.g("d.ln_final.scale").zero_()     # BART has no final LayerNorm: parameters exist in the tree, unused
            ps.g("d.ln_final.bias").zero_()
        # ---------------- backward: decoder layers ----------------
        d_enc_kv = b.get("tr.d_enc_kv", (Mv, t.decoder_layers * 2 * d))
        dtmp = b.get("tr.dtmp", (M, d))
        if not t.pre_layernorm:
            dx = self._decoder_backward_postln(dx, km, enc_kv, d_enc_kv, B, T, S, tg)
        else:
            self._decoder_backward_preln(dx, km, enc_kv, d_enc_kv, B, T, S, tg)
        # embedding
        if self._drop(1) is not None:       # dropout after layernorm_embedding
            dmask = b.get("tr.dmask", (M, d))
            self._before_write(dmask)
            ops.act_bwd_colsum(dx, None, "none", dmask, None, self._workspace(ops.colsum_workspace_floats(M, d)),
                               dropout=self._drop(1))
            dx = dmask
        demb = self._ln_bwd(dx, b.t[tg + ".emb"], "d.ln_emb", (b.t[tg + ".emb.mean"], b.t[tg + ".emb.rstd"]), None, dtmp)
        gpos = ps.g("d.pos")
        gpos.zero_()
        self._before_write(ps.g("shared"))          # the lm_head wgrad must have landed before the scatter-add
        if pos is None:
            ops.embed_bwd(ids, demb, self.emb_scale, ps.g("shared"), gpos[t.position_offset:], B, T, t.pad_token_id)
        else:
            ops.embed_bwd(ids, demb, self.emb_scale, ps.g("shared"), None, B, T, t.pad_token_id)
            gpos.index_add_(0, (pos + t.position_offset).long(), demb.float())   # rare path (explicit position ids)
        # ---------------- backward: cross K/V projection ----------------
        d_enc = b.get("tr.d_enc", (Mv, d))
        self._dense_bwd(enc, d_enc_kv, "d.ca_kv", d_enc)
        if stage == 1:
            self._join_side()
            return ws
        self._backward_vision(B, S, Mv, dv, d)
        return ws

    dp_vision_tail_layers = 2       # encoder layers whose gradients form the last (exposed) all-reduce bucket

    def grad_split_offset(self):
        """Flat-buffer offset separating the gradients finished by stage 1 from those of stage 2."""
        return self.ps.layout.storages["proj.w"][0]

    def grad_split_offsets(self):
        """Flat-buffer offsets [after stage 1, after stage 2]: stage 1 = head, tied embedding, decoder, cross K/V;
        stage 2 = visual projection + the upper vision layers; stage 3 = the rest (the buffer is laid out in backward
        order for exactly this)."""
        L = self.c.num_hidden_layers
        k = min(self.dp_vision_tail_layers, L)
        second = self.ps.layout.storages[f"v.{k - 1}.fc2.w"][0] if k > 0 else self.ps.layout.storages["v.pre_ln.scale"][0]
        return [self.ps.layout.storages["proj.w"][0], second]

    def _backward_vision(self, B, S, Mv, dv, d, part=None):
        """part = None: everything.  ("head", hi, lo): visual projection + final LN + layers hi..lo;
        ("tail", hi, lo): layers hi..lo + the embeddings (continues from the buffers the head left)."""
        c, t, ps, b = self.c, self.t, self.ps, self.bufs
        enc = b.t["tr.enc.out"]
        d_enc = b.t["tr.d_enc"]
        te = "tr.enc"
        L = c.num_hidden_layers
        do_head = part is None or part[0] == "head"
        do_tail = part is None or part[0] == "tail"
        l_hi, l_lo = (L - 1, 0) if part is None else (part[1], part[2])
        dxv = b.get("tr.dxv", (Mv, dv))
        dvt = b.get("tr.dvt", (Mv, dv))
        if c.final_layernorm:
            dxv, dvt = dvt, dxv              # the final-LN backward below leaves the stream gradient in the other buffer
        if do_head:
            if c.final_layernorm:
                dxv, dvt = dvt, dxv
            # ---------------- backward: visual projection ----------------
            x_last = b.t[te + ".post"] if c.final_layernorm else b.t[te + f".x{L}"]
            self._dense_bwd(x_last, d_enc, "proj", dxv)
            if c.final_layernorm:
                self._ln_bwd(dxv, b.t[te + f".x{L}"], "v.post_ln", (b.t[te + ".post.mean"], b.t[te + ".post.rstd"]), None, dvt)
                dxv, dvt = dvt, dxv
            else:
                ps.g("v.post_ln.scale").zero_()      # dead parameters (pooled output unused): zero gradient
                ps.g("v.post_ln.bias").zero_()
        # ---------------- backward: vision layers ----------------
        Hv = c.num_attention_heads
        vscale = 1.0 / math.sqrt(c.head_dim)
        dgv = b.get("tr.dgv", (Mv, c.intermediate_size))
        duv = b.get("tr.duv", (Mv, c.intermediate_size))
        dqkvv = b.get("tr.dqkvv", (Mv, 3 * dv))
        for l in range(l_hi, l_lo - 1, -1):
            n = f"v.{l}"
            sfx = f".{l}"
            g_, u_, ln2 = b.t[te + ".g" + sfx], b.t[te + ".u" + sfx], b.t[te + ".ln2" + sfx]
            xm, x0 = b.t[te + ".xm" + sfx], b.t[te + f".x{l}"]
            if self.fuse_act_bwd:
                self._dense_bwd(g_, dxv, n + ".fc2", dgv, dgrad_act_bwd=(c.hidden_act, u_))
                self._dense_bwd(ln2, dgv, n + ".fc1", dvt)
            else:
                self._dense_bwd(g_, dxv, n + ".fc2", dgv)
                self._dense_bwd(ln2, dgv, n + ".fc1", dvt, act=c.hidden_act, u=u_, du=duv)
            self._ln_bwd(dvt, xm, n + ".ln2", (b.t[te + ".ln2.mean" + sfx], b.t[te + ".ln2.rstd" + sfx]), dxv, dxv)
            att, qkv, ln1 = b.t[te + ".att" + sfx], b.t[te + ".qkv" + sfx], b.t[te + ".ln1" + sfx]
            self._dense_bwd(att, dxv, n + ".o", dvt)
            self._before_write(dqkvv)
            ops.attention_bwd(qkv[:, :dv], qkv[:, dv:2 * dv], qkv[:, 2 * dv:], att, dvt, b.t[te + ".lse" + sfx], None,
                              False, dqkvv[:, :dv], dqkvv[:, dv:2 * dv], dqkvv[:, 2 * dv:], B, Hv, S, S, vscale)
            self._dense_bwd(ln1, dqkvv, n + ".qkv", dvt)
            self._ln_bwd(dvt, x0, n + ".ln1", (b.t[te + ".ln1.mean" + sfx], b.t[te + ".ln1.rstd" + sfx]), dxv, dxv)
        if not do_tail:
            self._join_side()
            return None
        # embeddings: pre-LN, class / position / patch kernel
        if c.pre_layernorm:
            demb_v = self._ln_bwd(dxv, b.t[te + ".emb"], "v.pre_ln", (b.t[te + ".pre.mean"], b.t[te + ".pre.rstd"]),
                                  None, dvt)
        else:
            demb_v = dxv
            ps.g("v.pre_ln.scale").zero_()
            ps.g("v.pre_ln.bias").zero_()
        ops.batch_sum(demb_v, B, S, dv, ps.g("v.pos"), dv)
        ps.g("v.cls").copy_(ps.g("v.pos")[0])
        dpo = b.get("tr.dpo", (B * (S - 1), dv))
        self._before_write(dpo)
        ops.drop_cls_rows(demb_v, dpo, B, S)
        if c.patch_bias:
            ops.act_bwd_colsum(dpo, None, "none", None, ps.g("v.patch.b"), self._workspace(
                ops.colsum_workspace_floats(dpo.shape[0], dv)))
        ops.gemm(b.t[te + ".patches"], dpo, a_mn=True, b_mn=True, out=ps.g("v.patch.w"))
        self._join_side()
