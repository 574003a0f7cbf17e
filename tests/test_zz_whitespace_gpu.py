"""detect_content on the GPU (whitespace_codes_kernel + the host window walk) against the oracle.

The host half is also verified on the CPU (tests/test_whitespace_product.py).  First run on a GPU: the driver's round-1
GPU test pass (it reported XPASS under the provisional xfail marker this file carried then; the marker is gone)."""
import numpy as np
import pytest

import oracle
from tests.test_whitespace_product import _images

pytestmark = pytest.mark.gpu


def test_detect_content_matches_oracle():
    import torch
    import imageflow_b200 as ifb
    assert ifb.device_count() > 0 and torch.cuda.is_available()
    b = ifb.Batch(0)
    cases = list(_images(120, 23))
    big = np.zeros((2160, 3840, 4), np.uint8)
    big[200:1900, 300:3500] = np.random.default_rng(1).integers(0, 256, (1700, 3200, 4), dtype=np.uint8)
    cases += [(big, 1, True), (big, 40, False), (np.zeros((600, 800, 4), np.uint8), 1, True)]
    for a, thr, am in cases:
        want = oracle.detect_content(a, thr, am)[0]
        assert ifb.detect_content(ifb.BitmapWindow.from_numpy(a, alpha_meaningful=am), thr) == want, ("host", a.shape, thr, am)
        h, w = a.shape[:2]
        pitch = (w * 4 + 63) // 64 * 64
        t = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
        v = t.as_strided((h, w, 4), (pitch, 4, 1))
        v.copy_(torch.from_numpy(a))
        got = b.detect_content(ifb.BitmapWindow.from_torch(v, alpha_meaningful=am), thr, stream=torch.cuda.current_stream().cuda_stream)
        assert got == want, ("device", a.shape, thr, am)
        if h >= 3 and w >= 3:       # the code map itself, byte for byte (the reference returns early below 3x3: no map)
            dc = torch.empty((h, w), dtype=torch.uint8, device="cuda")
            b.whitespace_codes(ifb.BitmapWindow.from_torch(v, alpha_meaningful=am), dc.data_ptr(), thr, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(dc.cpu().numpy(), oracle.whitespace_codes(a, thr, am)), ("codes", a.shape, thr, am)
    b.close()
