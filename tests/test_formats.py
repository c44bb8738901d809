"""format helpers against vectors produced by the reference's own utils.py
(tests/golden/make_golden_formats.py).  Bit-exact: these are index / fp32 reductions with the reference's operation order."""
import os

import numpy as np
import torch

import sovits_b200  # noqa: F401
from sovits_b200 import formats

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_formats.npz"))


def test_repeat_expand_2d_matches_reference_all_modes():
    for i, (s, t) in enumerate(GOLD["re_cases"]):
        x = torch.from_numpy(GOLD[f"re_in_{i}"])
        assert x.shape[-1] == s
        for mode in ("left", "nearest", "linear"):
            got = formats.repeat_expand_2d(x, int(t), mode)
            want = torch.from_numpy(GOLD[f"re_{mode}_{i}"])
            assert got.shape == want.shape and torch.equal(got, want), (i, mode)


def test_repeat_expand_left_index_is_monotone_and_covers_upsampling():
    idx = formats._left_index(431, 862)
    assert idx[0] == 0 and idx[-1] == 430 and (np.diff(idx) >= 0).all() and (np.diff(idx) <= 1).all()
    assert set(idx.tolist()) == set(range(431))          # upsampling never skips a source frame


def test_volume_extractor_matches_reference():
    for i in range(3):
        a = torch.from_numpy(GOLD[f"vol_in_{i}"])
        got = formats.Volume_Extractor(512).extract(a)
        want = torch.from_numpy(GOLD[f"vol_out_{i}"])
        assert got.shape == want.shape and torch.equal(got, want)
    assert formats.Volume_Extractor(512).extract(np.zeros((1, 2048), dtype=np.float32)).abs().max() == 0
