"""Parameter store: ONE flat fp32 master buffer (+ same-layout grad / Adam m / Adam v buffers and a bf16
shadow), addressed through the reference's Flax pytree names.

HBM layout: tensors that feed one GEMM are stored fused (q|k|v kernels side by side, all layers'
cross-attention k|v kernels side by side) so the forward pass needs no concatenation; the pytree the
reference exposes (`model.params`; `modeling_clip_vision_utils.py:99-117`) is a dict of
strided VIEWS into those fused tensors.  AdamW and the gradient all-reduce run over the flat buffers.
"""
from __future__ import annotations

import numpy as np
import torch

ALIGN = 64  # elements; keeps every tensor 256-byte (fp32) / 128-byte (bf16) aligned for TMA + vector access


class Layout:
    """name -> (offset, shape) in the flat buffer, plus pytree leaf -> (storage name, index expr)."""

    def __init__(self, config):
        self.config = config
        self.storages = {}
        self.order = []
        self.size = 0
        self.leaves = {}   # path tuple -> (storage, slicer)
        self._build()

    def _add(self, name, shape):
        n = int(np.prod(shape))
        self.storages[name] = (self.size, tuple(shape))
        self.order.append(name)
        self.size += (n + ALIGN - 1) // ALIGN * ALIGN

    def _leaf(self, path, storage, slicer=None, reshape=None):
        self.leaves[tuple(path)] = (storage, slicer, reshape)

    def _ln(self, path, name, d):
        self._add(name + ".scale", (d,))
        self._add(name + ".bias", (d,))
        self._leaf(path + ["scale"], name + ".scale")
        self._leaf(path + ["bias"], name + ".bias")

    def _dense(self, path, name, din, dout, bias=True):
        self._add(name + ".w", (din, dout))
        self._leaf(path + ["kernel"], name + ".w")
        if bias:
            self._add(name + ".b", (dout,))
            self._leaf(path + ["bias"], name + ".b")

    def _build(self):
        c, t = self.config.clip_vision_config, self.config.mbart_config
        dv, d, p = c.hidden_size, t.d_model, c.patch_size
        vm = ["model", "encoder", "vision_model"]
        # ---- decoder first (backward produces these gradients first -> early all-reduce buckets) ----
        self._add("flb", (t.vocab_size,))
        self._leaf(["final_logits_bias"], "flb", None, (1, t.vocab_size))
        self._add("shared", (t.vocab_size, d))
        self._leaf(["model", "shared", "embedding"], "shared")
        dp = ["model", "decoder"]
        self._ln(dp + ["layer_norm"], "d.ln_final", d)
        L = t.decoder_layers
        for l in reversed(range(L)):
            lp = dp + ["layers", str(l)]
            n = f"d.{l}"
            self._dense(lp + ["fc2"], n + ".fc2", t.decoder_ffn_dim, d)
            self._dense(lp + ["fc1"], n + ".fc1", d, t.decoder_ffn_dim)
            self._ln(lp + ["final_layer_norm"], n + ".ln_f", d)
            self._dense(lp + ["encoder_attn", "out_proj"], n + ".ca_o", d, d)
            self._dense(lp + ["encoder_attn", "q_proj"], n + ".ca_q", d, d)
            self._ln(lp + ["encoder_attn_layer_norm"], n + ".ln_ca", d)
            self._dense(lp + ["self_attn", "out_proj"], n + ".sa_o", d, d)
            self._add(n + ".sa_qkv.w", (d, 3 * d))
            self._add(n + ".sa_qkv.b", (3 * d,))
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                self._leaf(lp + ["self_attn", nm, "kernel"], n + ".sa_qkv.w", (slice(None), slice(j * d, (j + 1) * d)))
                self._leaf(lp + ["self_attn", nm, "bias"], n + ".sa_qkv.b", (slice(j * d, (j + 1) * d),))
            self._ln(lp + ["self_attn_layer_norm"], n + ".ln_sa", d)
        self._ln(dp + ["layernorm_embedding"], "d.ln_emb", d)
        self._add("d.pos", (t.max_position_embeddings + t.position_offset, d))
        self._leaf(dp + ["embed_positions", "embedding"], "d.pos")
        self._add("d.ca_kv.w", (d, L * 2 * d))
        self._add("d.ca_kv.b", (L * 2 * d,))
        for l in range(L):
            lp = dp + ["layers", str(l), "encoder_attn"]
            for j, nm in enumerate(("k_proj", "v_proj")):
                lo = l * 2 * d + j * d
                self._leaf(lp + [nm, "kernel"], "d.ca_kv.w", (slice(None), slice(lo, lo + d)))
                self._leaf(lp + [nm, "bias"], "d.ca_kv.b", (slice(lo, lo + d),))
        self._dense(["model", "visual_projection"], "proj", dv, d)
        # ---- vision encoder ----
        self._ln(vm + ["post_layernorm"], "v.post_ln", dv)
        for l in reversed(range(c.num_hidden_layers)):
            lp = vm + ["encoder", "layers", str(l)]
            n = f"v.{l}"
            self._dense(lp + ["mlp", "fc2"], n + ".fc2", c.intermediate_size, dv)
            self._dense(lp + ["mlp", "fc1"], n + ".fc1", dv, c.intermediate_size)
            self._ln(lp + ["layer_norm2"], n + ".ln2", dv)
            self._dense(lp + ["self_attn", "out_proj"], n + ".o", dv, dv)
            self._add(n + ".qkv.w", (dv, 3 * dv))
            self._add(n + ".qkv.b", (3 * dv,))
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                self._leaf(lp + ["self_attn", nm, "kernel"], n + ".qkv.w", (slice(None), slice(j * dv, (j + 1) * dv)))
                self._leaf(lp + ["self_attn", nm, "bias"], n + ".qkv.b", (slice(j * dv, (j + 1) * dv),))
            self._ln(lp + ["layer_norm1"], n + ".ln1", dv)
        self._ln(vm + ["pre_layrnorm"], "v.pre_ln", dv)
        self._add("v.pos", (c.num_tokens, dv))
        self._leaf(vm + ["embeddings", "position_embedding", "embedding"], "v.pos")
        self._add("v.cls", (dv,))
        self._leaf(vm + ["embeddings", "class_embedding"], "v.cls")
        self._add("v.patch.w", (p * p * 3, dv))
        self._leaf(vm + ["embeddings", "patch_embedding", "kernel"], "v.patch.w", None, (p, p, 3, dv))
        if c.patch_bias:
            self._add("v.patch.b", (dv,))
            self._leaf(vm + ["embeddings", "patch_embedding", "bias"], "v.patch.b")


class ParamStore:
    def __init__(self, config, device="cuda", with_optimizer=False):
        self.config = config
        self.layout = Layout(config)
        self.device = torch.device(device)
        n = self.layout.size
        self.master = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        self.grad = None
        self.adam_m = None
        self.adam_v = None
        self.step_count = 0
        if with_optimizer:
            self.ensure_optimizer()

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.master)

    def ensure_optimizer(self):
        self.ensure_grad()
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.master)
            self.adam_v = torch.zeros_like(self.master)

    # ---- raw storage access -------------------------------------------------------------------
    def view(self, buf, name):
        off, shape = self.layout.storages[name]
        return buf[off:off + int(np.prod(shape))].view(shape)

    def w(self, name):      # bf16 shadow (GEMM / gather operands)
        return self.view(self.shadow, name)

    def f(self, name):      # fp32 master (biases, LayerNorm)
        return self.view(self.master, name)

    def g(self, name):      # fp32 gradient
        return self.view(self.grad, name)

    # ---- reference-facing pytree ----------------------------------------------------------------
    def tree(self, buf=None):
        """Nested dict with the Flax names; leaves are (possibly strided) views of `buf`."""
        buf = self.master if buf is None else buf
        out = {}
        for path, (storage, slicer, reshape) in self.layout.leaves.items():
            v = self.view(buf, storage)
            if slicer is not None:
                v = v[slicer]
            if reshape is not None:
                v = v.view(reshape)
            node = out
            for k in path[:-1]:
                node = node.setdefault(k, {})
            node[path[-1]] = v
        return out

    def load_tree(self, tree):
        """Copy a nested dict of numpy arrays / tensors (Flax names) into the master buffer."""
        mine = self.tree()
        seen = [0]

        def rec(dst, src, path):
            for k, v in dst.items():
                if k not in src:
                    raise KeyError("missing parameter " + "/".join(path + (k,)))
                if isinstance(v, dict):
                    rec(v, src[k], path + (k,))
                else:
                    s = src[k]
                    s = torch.from_numpy(np.ascontiguousarray(s)) if isinstance(s, np.ndarray) else s
                    if tuple(s.shape) != tuple(v.shape):
                        raise ValueError(f"shape mismatch at {'/'.join(path + (k,))}: {tuple(s.shape)} vs {tuple(v.shape)}")
                    v.copy_(s.to(self.device, torch.float32))
                    seen[0] += 1
        rec(mine, tree, ())
        self.refresh_shadow()
        return seen[0]

    def refresh_shadow(self):
        from . import ops
        ops.cast_f32_to_bf16(self.master, self.shadow)

    def to_numpy_tree(self, buf=None):
        def rec(t):
            return {k: (rec(v) if isinstance(v, dict) else v.detach().float().cpu().numpy().copy()) for k, v in t.items()}
        return rec(self.tree(buf))
