"""Writes tests/golden/transform_golden.npz: outputs of the reference's image Transform (main.py:165-179) computed by
torchvision ITSELF in the build container (torchvision 0.26, `antialias=False` = the tensor path of the 2021 release the
reference ran), on seeded synthetic uint8 images of awkward sizes.

    python tests/golden/gen_golden_transform.py
"""
import os

import numpy as np
import torch
from torchvision.transforms import CenterCrop, ConvertImageDtype, Normalize, Resize
from torchvision.transforms.functional import InterpolationMode

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)
CASES = [  # (H, W, S, kind)
    (37, 53, 32, "noise"), (53, 37, 32, "smooth"), (32, 32, 32, "noise"), (100, 31, 16, "smooth"), (31, 100, 16, "noise"),
    (480, 640, 224, "smooth"), (333, 500, 224, "noise"), (224, 224, 224, "smooth"), (500, 375, 224, "photo"),
    (225, 1000, 224, "photo"), (1200, 800, 224, "photo"), (150, 120, 224, "smooth"),
]


def make_image(h, w, kind, seed):
    rng = np.random.RandomState(seed)
    if kind == "noise":
        return rng.randint(0, 256, (3, h, w)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    chans = []
    for c in range(3):
        f = rng.uniform(0.01, 0.2, 4)
        v = 127.5 + 80 * np.sin(f[0] * xx + f[1] * yy + c) + 47 * np.cos(f[2] * xx - f[3] * yy)
        if kind == "photo":   # smooth + edges + saturated regions + a little noise
            v += 60 * ((xx // 23 + yy // 17) % 2) + rng.normal(0, 6, (h, w))
        chans.append(np.clip(np.rint(v), 0, 255))
    return np.stack(chans).astype(np.uint8)


def main():
    out = {}
    for i, (h, w, s, kind) in enumerate(CASES):
        img = make_image(h, w, kind, 100 + i)
        t = torch.from_numpy(img)
        with torch.no_grad():
            r = Resize([s], interpolation=InterpolationMode.BICUBIC, antialias=False)(t)
            c = CenterCrop(s)(r)
            f = Normalize(MEAN, STD)(ConvertImageDtype(torch.float)(c))
        out[f"case{i}_shape"] = np.array([h, w, s])
        out[f"case{i}_kind"] = np.array(kind)
        out[f"case{i}_seed"] = np.array(100 + i)
        out[f"case{i}_resized_hw"] = np.array(r.shape[1:])
        out[f"case{i}_u8"] = c.permute(1, 2, 0).numpy()                    # uint8 [S,S,3] after Resize + CenterCrop
        if s <= 32:
            out[f"case{i}_f32"] = f.permute(1, 2, 0).numpy()               # float32 NHWC row (small cases only)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transform_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
