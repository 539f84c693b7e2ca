"""ORACLE — test infrastructure only (imported by tests/, smoke() and bench.py's CPU legs, never by the product).

CPU restatement of the reference's image `Transform` (`main.py:165-179`, identical in `evaluation.py:33-50`):

    Resize([image_size], interpolation=BICUBIC) -> CenterCrop(image_size) -> ConvertImageDtype(float) -> Normalize(mean, std)

applied to the uint8 CHW tensor that `torchvision.io.read_image(..., RGB)` returns (`main.py:224-226`), followed by the
NHWC permute of `collate_fn` (`main.py:495`).

The arithmetic lives in a THIRD-PARTY dependency that is not vendored under /root/reference and is not pinned in its
requirements.txt: torchvision (the July-2021 release the project was written against is 0.10.0) on top of ATen's
`upsample_bicubic2d`.  Restated here from their published algorithms:

  * torchvision `functional.resize` with a one-element size: shorter edge -> S, longer edge -> int(S * long / short)
    (`_compute_resized_output_size`); tensor inputs take `functional_tensor.resize`, which in 0.10 never antialiases
    (`antialias=None` -> False for tensors): cast uint8 -> float32, `torch.nn.functional.interpolate(mode="bicubic",
    align_corners=False)`, clamp to [0, 255], `torch.round` (half to even) back to uint8;
  * ATen bicubic (`UpSample.h`): scale = in / out; source x = scale * (dst + 0.5) - 0.5; taps floor(x) - 1 .. + 2 with
    indices clamped to the image; Keys cubic convolution coefficients with A = -0.75; the CPU kernel interpolates the
    innermost (x) axis first, then combines the four rows (`UpSampleKernel.cpp`, `interpolate<2>`);
  * torchvision `center_crop`: top = int(round((H' - S) / 2.0)), left = int(round((W' - S) / 2.0)) with Python's
    round-half-to-even;
  * `ConvertImageDtype(float)`: x / 255; `Normalize`: (x - mean[c]) / std[c].

PINNING: checked against torchvision 0.26 (`Resize(..., antialias=False)` + `CenterCrop`) run in the build container —
`tests/golden/gen_golden_transform.py` wrote `tests/golden/transform_golden.npz`; `tests/test_transform_cpu.py` holds
the comparison.  ATen's CPU kernel is built with -mfma and the compiler contracts its multiply-adds; this restatement
(and the CUDA kernel, with explicit fmaf) contracts the same expressions.  Measured against the goldens: 7 of the 12
cases identical, 13 bytes of 1.06 M off by one LSB in the rest (an fp32 value within an ulp of a .5 boundary; without
the contraction it is 34 bytes) - the test allows <= 1 LSB on <= 1e-4 of the bytes of a case.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
CUBIC_A = F32(-0.75)


def resized_size(h: int, w: int, size: int):
    """torchvision `_compute_resized_output_size` for `Resize([size])`: (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def crop_offsets(h: int, w: int, size: int):
    """torchvision `center_crop` offsets (top, left) for an image already at least `size` on both edges."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def _fma(a, b, c):
    """fp32 fused multiply-add: the product of two float32 is exact in float64; the float64 sum rounds once more before
    the final rounding to float32, which differs from a true FMA only when that sum is an exact float32 midpoint."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def _cubic1(x):   # |x| <= 1:  ((A + 2) x - (A + 3)) x x + 1, contracted the way -mfma builds of ATen contract it
    p = _fma(CUBIC_A + F32(2), x, -(CUBIC_A + F32(3)))
    return _fma((p * x).astype(F32), x, F32(1))


def _cubic2(x):   # 1 < |x| < 2:  ((A x - 5A) x + 8A) x - 4A
    p = _fma(CUBIC_A, x, -F32(5) * CUBIC_A)
    p = _fma(p, x, F32(8) * CUBIC_A)
    return _fma(p, x, -F32(4) * CUBIC_A)


def bicubic_taps(in_size: int, out_size: int, dst: np.ndarray):
    """ATen `area_pixel_compute_source_index(cubic=True, align_corners=False)` + `get_cubic_upsample_coefficients`:
    for every destination index -> 4 clamped source indices [n,4] and 4 float32 weights [n,4]."""
    scale = F32(in_size) / F32(out_size)
    real = _fma(scale, dst.astype(F32) + F32(0.5), F32(-0.5))
    fl = np.floor(real)
    t = (real - fl).astype(F32)
    base = fl.astype(np.int64)
    idx = np.clip(base[:, None] + np.arange(-1, 3)[None, :], 0, in_size - 1)
    t2 = (F32(1) - t).astype(F32)                       # ATen: x2 = 1 - t; coeffs[3] = conv2(x2 + 1)
    w = np.stack([_cubic2(t + F32(1)), _cubic1(t), _cubic1(t2), _cubic2(t2 + F32(1))], axis=1).astype(F32)
    return idx, w


def resize_crop_u8(img_chw: np.ndarray, size: int) -> np.ndarray:
    """uint8 [3,H,W] -> uint8 [size,size,3] (NHWC row of the batch): Resize([size], BICUBIC, no antialias) + CenterCrop."""
    assert img_chw.dtype == np.uint8 and img_chw.ndim == 3
    C, H, W = img_chw.shape
    nh, nw = resized_size(H, W, size)
    top, left = crop_offsets(nh, nw, size)
    yi, wy = bicubic_taps(H, nh, np.arange(top, top + size))
    xi, wx = bicubic_taps(W, nw, np.arange(left, left + size))
    src = img_chw.astype(F32)
    out = np.empty((size, size, C), np.uint8)
    for c in range(C):
        rows = src[c][yi]                                  # [S, 4, W]
        # x axis first (`Interpolate<2>::eval`): t = s0*w0; t = fma(s1, w1, t); ...
        acc = (rows[:, :, xi[:, 0]] * wx[None, None, :, 0]).astype(F32)           # [S, 4, S]
        for j in range(1, 4):
            acc = _fma(rows[:, :, xi[:, j]], wx[None, None, :, j], acc)
        val = (acc[:, 0, :] * wy[:, 0, None]).astype(F32)
        for i in range(1, 4):
            val = _fma(acc[:, i, :], wy[:, i, None], val)
        out[:, :, c] = np.rint(np.clip(val, F32(0), F32(255))).astype(np.uint8)   # rint = half to even
    return out


def normalize(u8_nhwc: np.ndarray, mean, std) -> np.ndarray:
    """ConvertImageDtype(float) + Normalize on NHWC uint8 -> float32 NHWC (what `collate_fn` hands the model)."""
    x = u8_nhwc.astype(F32) / F32(255)
    return ((x - np.asarray(mean, F32)) / np.asarray(std, F32)).astype(F32)


def transform_batch(images, size: int, mean, std) -> np.ndarray:
    """The reference's per-image Transform + collate permute for a list of uint8 CHW images -> float32 [n,S,S,3]."""
    return np.stack([normalize(resize_crop_u8(im, size), mean, std) for im in images])
