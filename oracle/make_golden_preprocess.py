"""Generate tests/golden/preprocess.npz from the REAL libraries (Pillow + torchvision, present in the build container):
the image transform the reference applies in front of the hot path (data.py:898-902 -> open_clip image_transform).
Test infrastructure; run with ``python -m oracle.make_golden_preprocess``.  Small frames carry their full fp32 output,
CALVIN-sized frames (200x200 static, 84x84 gripper -> 224) carry a SHA-256 of the fp32 bytes."""
import hashlib
import os

import numpy as np


def frames(seed, h, w):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = ((np.sin(yy / 7.0)[..., None] * 60 + np.cos(xx / 5.0)[..., None] * 60 + 128) + np.arange(3) * 20).clip(0, 255)
    return np.where(rng.random((h, w, 1)) < 0.5, base, smooth.astype(np.uint8)).astype(np.uint8)


SMALL = [(50, 50, 28), (21, 21, 28), (30, 75, 28), (120, 48, 28), (28, 28, 28), (28, 61, 28), (57, 56, 56)]
LARGE = [(200, 200, 224), (84, 84, 224), (480, 640, 224), (224, 224, 224)]


def main():
    import torch
    from PIL import Image
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode

    def real(img, size):
        tf = T.Compose([T.Resize(size, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(size), T.ToTensor(),
                        T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
        return tf(Image.fromarray(img)).numpy()

    out = {}
    for i, (h, w, s) in enumerate(SMALL):
        out[f"small{i}_out"] = real(frames(100 + i, h, w), s)
    for i, (h, w, s) in enumerate(LARGE):
        out[f"large{i}_sha"] = np.frombuffer(hashlib.sha256(real(frames(200 + i, h, w), s).tobytes()).digest(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
