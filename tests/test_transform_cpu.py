"""Image Transform hand-off, CPU side: the oracle restatement against torchvision's own outputs
(tests/golden/transform_golden.npz, written by tests/golden/gen_golden_transform.py) and the host-side integer formulas."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import gen_golden_transform as gg  # noqa: E402

import mic_b200  # noqa: E402
from mic_b200 import transforms  # noqa: E402
from oracle import reference_transform as rt  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transform_golden.npz"))
MAX_LSB_FRACTION = 1e-4      # bytes of a case allowed to differ from torchvision, by exactly one LSB (FMA contraction)


@pytest.mark.parametrize("i", range(len(gg.CASES)))
def test_oracle_matches_torchvision_golden(i):
    h, w, s, kind = gg.CASES[i]
    img = gg.make_image(h, w, kind, int(GOLD[f"case{i}_seed"]))
    assert tuple(GOLD[f"case{i}_resized_hw"]) == rt.resized_size(h, w, s)
    got = rt.resize_crop_u8(img, s)
    want = GOLD[f"case{i}_u8"]
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1
    assert (d > 0).mean() <= MAX_LSB_FRACTION, (int((d > 0).sum()), d.size)
    if s <= 32:    # the full Transform output (ConvertImageDtype + Normalize, NHWC) of the small cases: exact where u8 is
        f = rt.normalize(want, gg.MEAN, gg.STD)
        np.testing.assert_allclose(f, GOLD[f"case{i}_f32"], rtol=0, atol=2e-7)


def test_seven_cases_are_bit_identical_to_torchvision():
    exact = 0
    for i, (h, w, s, kind) in enumerate(gg.CASES):
        got = rt.resize_crop_u8(gg.make_image(h, w, kind, 100 + i), s)
        exact += int((got == GOLD[f"case{i}_u8"]).all())
    assert exact >= 7


def test_host_formulas_match_oracle_and_torchvision_rules():
    for h, w, s in [(480, 640, 224), (333, 500, 224), (1200, 800, 224), (225, 1000, 224), (31, 100, 16), (224, 224, 224),
                    (7, 1000, 5), (501, 500, 3)]:
        assert transforms.resized_size(h, w, s) == rt.resized_size(h, w, s)
        nh, nw = rt.resized_size(h, w, s)
        assert min(nh, nw) == s and (nh >= s and nw >= s)
        assert transforms.crop_offsets(nh, nw, s) == rt.crop_offsets(nh, nw, s)
    # Python round() is half-to-even: an odd margin of 77 -> 38, of 79 -> 40 (torchvision center_crop)
    assert transforms.crop_offsets(224, 301, 224) == (0, 38)
    assert transforms.crop_offsets(224, 303, 224) == (0, 40)


def test_descriptor_table_layout():
    bt = transforms.BatchTransform.__new__(transforms.BatchTransform)
    bt.size = 224
    desc, total = bt.describe([(480, 640), (333, 500)])
    assert total == 3 * 480 * 640 + 3 * 333 * 500
    assert desc.tolist()[0] == [0, 480, 640, 224, 298, 0, 37, 0]
    assert desc.tolist()[1] == [3 * 480 * 640, 333, 500, 224, 336, 0, 56, 0]


def test_bicubic_weights_sum_to_one_and_identity_resize():
    idx, w = rt.bicubic_taps(100, 37, np.arange(37))
    np.testing.assert_allclose(w.sum(1), 1.0, atol=1e-6)
    assert idx.min() >= 0 and idx.max() <= 99
    img = gg.make_image(32, 32, "noise", 5)
    assert (rt.resize_crop_u8(img, 32) == img.transpose(1, 2, 0)).all()       # same size: every weight is (0, 1, 0, 0)
