"""The operand view proposed for the narrow stages (DESIGN.md K2 "what next", tools/s2d_model.py): class-major rows +
space-to-depth + block-Toeplitz weights compute the reference's dilated Conv1d exactly, and the measured MMA cost law says
what the view is worth.  Host-side model only - the shipped kernels use the per-tap mapping."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import s2d_model as M  # noqa: E402


@pytest.mark.parametrize("C,J", [(16, 4), (32, 2)])
@pytest.mark.parametrize("k", [3, 7, 11])
@pytest.mark.parametrize("d", [1, 3, 5])
def test_space_to_depth_view_is_the_same_convolution(C, J, k, d):
    rng = np.random.default_rng(1000 * C + 10 * k + d)
    L = 257                                   # not a multiple of d * J: ragged classes and a ragged last virtual row
    x = rng.standard_normal((C, L))
    w = rng.standard_normal((C, C, k)) / np.sqrt(C * k)
    ref = M.conv_direct(x, w, d)
    got = M.conv_s2d(x, w, d, J)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("k", [3, 7, 11])
def test_only_the_non_zero_toeplitz_blocks_are_issued(k):
    # J + k - 1 K steps per virtual row instead of J * k per J rows: the useful share of the wide instructions is k / (J + k - 1)
    J, C = 4, 16
    taps, steps = M.toeplitz_weights(np.ones((C, C, k)), J)
    assert len(steps) == J + k - 1
    dense = sum(int(np.count_nonzero(b)) for b in taps.values())
    assert dense == k * J * C * C             # every weight appears once per output phase


def test_predicted_gain_on_the_measured_cost_law():
    # C = 16: 36 clk per N = 16 instruction today, 48 clk per N = 64 instruction in the view
    gains = {}
    for k in (3, 7, 11):
        today, s2d = M.mma_cycles(16, k)
        gains[k] = today / s2d
    assert gains[3] == pytest.approx(1.5, abs=0.01) and gains[7] == pytest.approx(2.1, abs=0.01) and gains[11] == pytest.approx(2.36, abs=0.01)
    # C = 32 (J = 2): 40 clk per N = 32 instruction today; the view needs (2 + k - 1) x 2 K steps of N = 64 per 256 time steps
    today, s2d = M.mma_cycles(32, 11)
    assert today / s2d == pytest.approx(1.53, abs=0.01)
