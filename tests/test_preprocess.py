"""Image transform in front of the hot path: the Pillow-exact restatement is pinned against fixtures made
with the real Pillow/torchvision (and against the live libraries when importable); the CUDA kernel must then match the
restatement BIT-exactly."""
import hashlib
import os

import numpy as np
import pytest

from oracle import pil_resize as R
from oracle.make_golden_preprocess import LARGE, SMALL, frames

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))


def test_restatement_matches_golden_small():
    for i, (h, w, s) in enumerate(SMALL):
        got = R.clip_transform(frames(100 + i, h, w), s)
        assert got.dtype == np.float32 and got.shape == (3, s, s)
        assert np.array_equal(got, GOLD[f"small{i}_out"]), (h, w, s)


def test_restatement_matches_golden_calvin_sizes():
    for i, (h, w, s) in enumerate(LARGE):
        got = R.clip_transform(frames(200 + i, h, w), s)
        assert hashlib.sha256(got.tobytes()).digest() == GOLD[f"large{i}_sha"].tobytes(), (h, w, s)


def test_restatement_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    for seed, (h, w, oh, ow) in enumerate([(200, 200, 224, 224), (84, 84, 224, 224), (300, 120, 77, 31), (64, 64, 64, 200)]):
        img = frames(seed, h, w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(R.resize_bicubic_u8(img, oh, ow), ref)


def test_product_host_transform_is_the_reference_transform():
    """deer_vla_b200.eval_utils.clip_image_transform (Pillow path) == the restatement == torchvision's composition."""
    pytest.importorskip("PIL.Image")
    from deer_vla_b200.eval_utils import clip_image_transform
    for seed, (h, w, s) in enumerate([(200, 200, 224), (84, 84, 224), (120, 300, 56), (300, 120, 56)]):
        img = frames(seed + 50, h, w)
        assert np.array_equal(clip_image_transform(s)(img).numpy(), R.clip_transform(img, s)), (h, w, s)


def test_coefficients_are_normalised():
    for a, b in [(200, 224), (84, 224), (500, 224), (224, 224)]:
        bounds, kk, ksize = R.precompute_coeffs(a, b)
        assert kk.shape == (b, ksize) and (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= a).all()
        assert np.abs(kk.sum(1) - (1 << R.PRECISION_BITS)).max() <= ksize      # fixed-point rows sum to ~1.0


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(200, 200), (84, 84), (56, 56), (120, 300), (300, 120), (31, 57), (480, 640)])
def test_gpu_transform_bit_exact(h, w):
    import torch
    from deer_vla_b200 import DeerConfig
    from deer_vla_b200.engine import DeerEngine
    from oracle import deer_oracle as O
    ocfg = O.config_tiny()
    eng = DeerEngine(DeerConfig(**ocfg.to_dict()), max_batch=4, max_seq=8, precision="fp32")
    imgs = np.stack([frames(7 + i, h, w) for i in range(3)])
    got = eng.preprocess(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], R.clip_transform(imgs[i], ocfg.image_size)), (h, w, i)
    eng.close()


@pytest.mark.gpu
def test_gpu_transform_224_and_step_u8():
    import torch
    from deer_vla_b200 import DeerConfig
    from deer_vla_b200.engine import DeerEngine
    from oracle import deer_oracle as O
    ocfg = O.config_tiny(image_size=224)
    W = O.synth_weights(ocfg, 0)
    eng = DeerEngine(DeerConfig(**ocfg.to_dict()), max_batch=2, max_seq=16, precision="fp32")
    eng.load_state_dict(W)
    eng.set_thresholds([0.3, 0.25, 1e8])
    rgb = np.stack([frames(1, 200, 200), frames(2, 200, 200)])
    grip = np.stack([frames(3, 84, 84), frames(4, 84, 84)])
    a = eng.preprocess(torch.from_numpy(rgb).cuda())
    b = eng.preprocess(torch.from_numpy(grip).cuda())
    for i in range(2):
        assert np.array_equal(a[i].cpu().numpy(), R.clip_transform(rgb[i], 224))
        assert np.array_equal(b[i].cpu().numpy(), R.clip_transform(grip[i], 224))
    ids = torch.stack([O.synth_tokens(ocfg, 5, 12), O.synth_tokens(ocfg, 6, 12)])
    eng.reset()
    d = eng.step(a, b, ids.cuda())
    eng.reset()
    u = eng.step_u8(torch.from_numpy(rgb).pin_memory(), torch.from_numpy(grip).pin_memory(), ids)
    assert torch.equal(d["exit_layer"].cpu(), u["exit_layer"])
    assert torch.equal(d["action"].cpu(), u["action"])          # same kernels, same inputs -> identical
    eng.close()
