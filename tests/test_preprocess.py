"""n3 — pre-processing in front of the path (generate_reports_for_images.py:129-147 `get_image_tensor`).

CPU: the oracle (real cv2.resize + restated albumentations 1.1.0 transforms) and the restated INTER_AREA arithmetic
against cv2, bit for bit.  GPU (-m gpu): `rgrg_preprocess` against the oracle, bit for bit, on 3056 x 2544 inputs
(the size of MIMIC-CXR frontal images) and on the integer-scale / no-resize special cases."""
import numpy as np
import pytest
import torch

import preprocess_oracle as P

SIZES = [(3056, 2544), (2544, 3056), (2048, 2048), (1024, 1024), (1500, 1000), (2021, 2021), (512, 512), (512, 300), (4280, 3520),
         (1025, 513)]


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    base[: h // 3] = (np.arange(w) % 256).astype(np.uint8)[None, :]  # smooth gradient: many exact .5 rounding cases
    return base


@pytest.mark.parametrize("h,w", SIZES[:7])
def test_inter_area_restatement_matches_cv2(h, w):
    cv2 = pytest.importorskip("cv2")
    img = _image(h, w, 1)
    nh, nw = P.target_size(h, w)
    if (nh, nw) == (h, w):
        return
    ref = cv2.resize(img, dsize=(nw, nh), interpolation=cv2.INTER_AREA)
    assert np.array_equal(ref, P.resize_area_restated(img, nh, nw))


def test_reference_transform_shapes_and_padding():
    pytest.importorskip("cv2")
    out = P.preprocess_reference(_image(3056, 2544, 2))
    assert out.shape == (1, 512, 512) and out.dtype == np.float32
    nh, nw = P.target_size(3056, 2544)
    assert (nh, nw) == (512, 426) and P.pad_offsets(nh, nw) == (0, 43)
    mean, denom = P.norm_constants()
    pad_val = (np.float32(0) - mean) * denom
    assert np.all(out[0, :, :43] == pad_val) and np.all(out[0, :, 43 + 426:] == pad_val)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", SIZES)
def test_gpu_preprocess_bit_exact(h, w):
    pytest.importorskip("cv2")
    from rgrg_b200 import Engine

    eng = Engine(0)
    imgs = [_image(h, w, 3), _image(h, w, 4)]
    out = eng.preprocess(imgs).cpu().numpy()
    assert out.shape == (2, 1, 512, 512)
    for i, im in enumerate(imgs):
        ref = P.preprocess_reference(im)
        assert np.array_equal(out[i], ref), "max |d| = %g" % np.abs(out[i] - ref).max()
    # device-resident source image
    out2 = eng.preprocess([torch.from_numpy(imgs[0]).cuda()]).cpu().numpy()
    assert np.array_equal(out2[0], out[0])
    eng.close()
