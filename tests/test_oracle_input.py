"""Pins oracle/input_oracle.py (the input transform of SURVEY 8 f-4) against the real third-party library: OpenCV is
installed in the build container, so the fixed-point bilinear restatement is compared with cv2.resize bit for bit."""
import numpy as np
import pytest

from oracle.input_oracle import network_input, normalize_to_chw, resize_linear_u8

cv2 = pytest.importorskip("cv2")

CASES = [((37, 53), (256, 384)), ((480, 640), (256, 384)), ((256, 384), (256, 384)), ((1080, 1920), (256, 384)),
         ((100, 100), (64, 48)), ((8, 8), (24, 40)), ((301, 517), (256, 384)), ((256, 384), (512, 768)),
         ((3, 5), (16, 24)), ((1, 7), (8, 8)), ((600, 400), (256, 384))]


@pytest.mark.parametrize("src_hw,dst_hw", CASES)
def test_resize_matches_cv2_bit_exactly(src_hw, dst_hw):
    rng = np.random.default_rng(src_hw[0] * 7919 + dst_hw[1])
    img = rng.integers(0, 256, (src_hw[0], src_hw[1], 3), dtype=np.uint8)
    want = cv2.resize(img, (dst_hw[1], dst_hw[0]), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(resize_linear_u8(img, dst_hw), want)


def test_normalize_and_layout():
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3) * 14
    out = normalize_to_chw(img)
    assert out.dtype == np.float32 and out.shape == (3, 2, 3)
    assert np.array_equal(out[1], img[:, :, 1].astype(np.float32) * np.float32(1.0 / 255.0))
    assert network_input(np.full((10, 10, 3), 255, np.uint8), (8, 16)).max() == np.float32(255) * np.float32(1.0 / 255.0)
