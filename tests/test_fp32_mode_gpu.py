"""fp32 verification mode (dtype=float32: fp32 storage + fp32 SIMT kernels, csrc/fp32_path.cu) against the oracle at
the tolerances the north star states for BASELINE configs[0]: logits within 1e-3 (relative to the largest logit),
loss within 1e-4."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_model as rm  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_RTOL = 1e-3          # north star: logits rel <= 1e-3
LOSS_ATOL = 1e-4           # north star: loss <= 1e-4


@pytest.mark.parametrize("variant", ["clip_mbart", "vit_bart"])
@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_fp32_mode_matches_oracle_tiny(variant, eps):
    cfg = mic_b200.tiny_config() if variant == "clip_mbart" else mic_b200.tiny_vit_bart_config()
    params = synthetic.make_params(cfg, seed=3, perturbed=True, std=0.08)
    batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=1, min_len=4)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, dtype="float32", _do_init=False)
    model.params = params
    logits = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"]).logits
    p = rm.to_torch_tree(params)
    ref = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg)
    ref = ref.detach().numpy()
    got = logits.cpu().numpy()
    assert got.shape == ref.shape
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    assert rel <= LOGIT_RTOL, rel
    loss = float(model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"],
                            batch["input_ids"], eps))
    ref_loss = float(rm.loss_fn(torch.from_numpy(ref), batch["input_ids"], batch["attention_mask"], eps))
    assert abs(loss - ref_loss) <= LOSS_ATOL, (loss, ref_loss)


def test_fp32_mode_reproduces_config1_golden_at_north_star_tolerance():
    """BASELINE configs[0]: full CLIP-ViT-B/32 + mBART-50, batch 8, 64 tokens, seeded random init, fp32."""
    gold = np.load(os.path.join(G, "config1_full_golden.npz"))
    cfg = mic_b200.clip_mbart_config()
    params = synthetic.make_params(cfg, seed=1, perturbed=False)
    batch = synthetic.make_batch(cfg, 8, 64, seed=0)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, dtype="float32", _do_init=False)
    model.params = params
    del params
    logits = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"]).logits
    assert tuple(logits.shape) == (8, 64, 250054)
    cols = torch.from_numpy(gold["cols"]).to(logits.device)
    sl = logits[:, :, cols].cpu().numpy()
    rel = np.abs(sl - gold["logits_slice"]).max() / float(gold["logits_absmax"])
    assert rel <= LOGIT_RTOL, rel
    lse = torch.logsumexp(logits, -1).cpu().numpy()
    np.testing.assert_allclose(lse, gold["lse"], atol=1e-4)
    am = logits.argmax(-1).cpu().numpy()
    clear = gold["top2_gap"] > 1e-3
    assert clear.mean() > 0.9
    assert (am[clear] == gold["argmax"][clear]).all()
    del logits
    for eps in (0.0, 0.1):
        loss = float(model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"],
                                batch["input_ids"], eps))
        assert abs(loss - float(gold[f"loss_eps{eps}"])) <= LOSS_ATOL, (eps, loss, float(gold[f"loss_eps{eps}"]))
