"""Image Transform hand-off on the GPU: `mic_resize_crop_u8` through the C-ABI against the oracle
(bit-exact: byte work) and against torchvision's own outputs (golden fixtures), and the end-to-end hand-off into the model."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import gen_golden_transform as gg  # noqa: E402

import mic_b200  # noqa: E402
from mic_b200 import synthetic, transforms  # noqa: E402
from oracle import reference_transform as rt  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transform_golden.npz"))


def _images(cases):
    return [gg.make_image(h, w, kind, 100 + i) for i, (h, w, s, kind) in cases]


@pytest.mark.parametrize("size", [224, 32, 16])
def test_batch_kernel_is_bit_exact_against_the_oracle_and_within_one_lsb_of_torchvision(size):
    cases = [(i, c) for i, c in enumerate(gg.CASES) if c[2] == size]
    imgs = [gg.make_image(h, w, kind, 100 + i) for i, (h, w, s, kind) in cases]
    bt = transforms.BatchTransform(size, "cuda:0")
    out = bt(imgs).cpu().numpy()
    assert out.shape == (len(imgs), size, size, 3)
    for k, (i, (h, w, s, kind)) in enumerate(cases):
        assert (out[k] == rt.resize_crop_u8(imgs[k], size)).all(), f"case {i}: kernel != oracle"
        d = np.abs(out[k].astype(np.int32) - GOLD[f"case{i}_u8"].astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() <= 1e-4
    # the staging buffers are reused: a second, differently composed batch through the same object
    out2 = bt(imgs[::-1] + [torch.from_numpy(imgs[0])]).cpu().numpy()
    assert (out2[-1] == out[0]).all() and (out2[0] == out[-1]).all()


def test_channel_first_output_and_extreme_aspect_ratios():
    rng = np.random.RandomState(3)
    imgs = [rng.randint(0, 256, (3, h, w)).astype(np.uint8) for h, w in [(5, 400), (400, 5), (1, 1), (17, 16), (2, 3)]]
    for size in (4, 17):
        nhwc = transforms.BatchTransform(size, "cuda:0")(imgs).cpu().numpy()
        nchw = transforms.BatchTransform(size, "cuda:0", channel_first=True)(imgs).cpu().numpy()
        assert (nchw.transpose(0, 2, 3, 1) == nhwc).all()
        for k, im in enumerate(imgs):
            assert (nhwc[k] == rt.resize_crop_u8(im, size)).all()


def test_full_size_batch_properties():
    """BASELINE-size batch (256 images, S = 224): identity on 224x224 inputs, constant images stay constant, and a
    checksum over the batch equals the checksum of the per-image oracle on a sample of it."""
    rng = np.random.RandomState(11)
    imgs = []
    for i in range(256):
        h, w = int(rng.randint(224, 520)), int(rng.randint(224, 700))
        if i % 16 == 0:
            imgs.append(np.full((3, h, w), i % 251, np.uint8))
        elif i % 16 == 1:
            imgs.append(rng.randint(0, 256, (3, 224, 224)).astype(np.uint8))
        else:
            imgs.append(gg.make_image(h, w, "photo", i))
    out = transforms.BatchTransform(224, "cuda:0")(imgs).cpu().numpy()
    for i in range(0, 256, 16):
        assert (out[i] == i % 251).all()                                           # weights sum to one exactly enough
        assert (out[i + 1] == imgs[i + 1].transpose(1, 2, 0)).all()                # same-size resize is the identity
    for i in (2, 77, 131, 255):
        assert (out[i] == rt.resize_crop_u8(imgs[i], 224)).all()


def test_per_image_transform_matches_the_reference_pipeline_values():
    i = 8
    h, w, s, kind = gg.CASES[i]
    img = gg.make_image(h, w, kind, 100 + i)
    got = mic_b200.Transform(s, "cuda:0")(torch.from_numpy(img)).cpu().numpy()           # float32 [3,S,S]
    want = rt.normalize(rt.resize_crop_u8(img, s), gg.MEAN, gg.STD).transpose(2, 0, 1)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_raw_images_through_the_model_equal_the_host_transformed_float_path():
    """Raw uint8 images -> BatchTransform -> model(uint8 pixel_values) == oracle Transform on the host -> model(float32)."""
    cfg = mic_b200.tiny_config()
    S = cfg.clip_vision_config.image_size
    rng = np.random.RandomState(5)
    imgs = [rng.randint(0, 256, (3, int(rng.randint(S, 3 * S)), int(rng.randint(S, 3 * S)))).astype(np.uint8) for _ in range(4)]
    batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=1, min_len=4)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    u8 = transforms.BatchTransform(S, "cuda:0")(imgs)
    a = model(u8, batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu().numpy()
    f32 = rt.transform_batch(imgs, S, cfg.clip_vision_config.image_mean, cfg.clip_vision_config.image_std)
    b = model(f32, batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu().numpy()
    assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max()
