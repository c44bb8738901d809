"""Host-side model of the operand view that would lift the narrow-stage MMA cost law ("what next").

Today a narrow ResBlock conv (C = 16 / 32 channels, k taps, dilation d) runs as k shifted GEMMs with M = 128 time steps,
N = C, K = C: one tcgen05.mma per tap and K step, and for N <= 64 such an instruction costs (4096 + 32 N) / 128 cycles - the
fetch of its 4 KB A operand - whatever N is (36 clk at N = 16 for 1/8 of the math of an N = 128 instruction).

The view modelled here keeps M = time but widens N to 64:
  * class-major rows: a conv with dilation d only ever combines samples of one residue class t mod d, so the tile is stored
    class by class (class r = x[:, r::d]); inside a class the conv has dilation 1;
  * space-to-depth by J = 128 B / (2 C) along the class (J = 4 for C = 16): J consecutive samples x C channels are ONE
    128-byte operand row = J*C "virtual channels"; the conv becomes a k' = ceil-ish((k - 1 + J) / J)-tap conv between virtual
    channels with block-Toeplitz weights  W'[(j, co), (i, ci), D] = W[co, ci, J*D + i - j + h]  (zero outside 0..k-1);
  * per virtual tap only the K steps (one per i) with a non-zero block are issued: J + k - 1 instructions of N = J*C = 64 per
    J*128 time steps instead of J*k instructions of N = C.

`conv_s2d` evaluates exactly that data flow with numpy (fp64, so the comparison with the direct convolution is tight);
`mma_cycles` applies the measured cost law to both mappings.  Nothing here is on the product path.
"""
from __future__ import annotations

import numpy as np


def conv_direct(x: np.ndarray, w: np.ndarray, d: int) -> np.ndarray:
    """y[co, t] = sum_{ci, tap} w[co, ci, tap] x[ci, t + (tap - h) d], zero padding (the reference's Conv1d, 'same')."""
    C, L = x.shape
    Co, Ci, k = w.shape
    h = (k - 1) // 2
    xp = np.zeros((C, L + 2 * h * d), dtype=np.float64)
    xp[:, h * d:h * d + L] = x
    y = np.zeros((Co, L), dtype=np.float64)
    for tap in range(k):
        y += w[:, :, tap].astype(np.float64) @ xp[:, tap * d:tap * d + L]
    return y


def toeplitz_weights(w: np.ndarray, J: int):
    """W'[D][(j, co), (i, ci)] for the virtual taps D = dmin .. dmax, plus the list of (D, i) K steps that are not all zero."""
    Co, Ci, k = w.shape
    h = (k - 1) // 2
    dmin = -((h + J - 1) // J)
    dmax = (h + J - 1) // J
    taps = {}
    steps = []
    for D in range(dmin, dmax + 1):
        blk = np.zeros((J * Co, J * Ci), dtype=np.float64)
        for i in range(J):
            used = False
            for j in range(J):
                tap = J * D + i - j + h
                if 0 <= tap < k:
                    blk[j * Co:(j + 1) * Co, i * Ci:(i + 1) * Ci] = w[:, :, tap]
                    used = True
            if used:
                steps.append((D, i))
        taps[D] = blk
    return taps, steps


def conv_s2d(x: np.ndarray, w: np.ndarray, d: int, J: int) -> np.ndarray:
    """The same convolution through class-major rows + space-to-depth by J + block-Toeplitz weights."""
    C, L = x.shape
    Co, Ci, k = w.shape
    taps, _ = toeplitz_weights(w, J)
    dmin, dmax = min(taps), max(taps)
    y = np.zeros((Co, L), dtype=np.float64)
    for r in range(d):                                   # one residue class at a time (dilation 1 inside it)
        xr = x[:, r::d]
        U = xr.shape[1]
        if U == 0:
            continue
        Tv = (U + J - 1) // J                            # virtual rows of this class
        pad_lo, pad_hi = -dmin, dmax
        xv = np.zeros((J * Ci, Tv + pad_lo + pad_hi), dtype=np.float64)    # [(i, ci), t'] with zero rows either side
        for i in range(J):
            seg = xr[:, i::J]
            xv[i * Ci:(i + 1) * Ci, pad_lo:pad_lo + seg.shape[1]] = seg
        yv = np.zeros((J * Co, Tv), dtype=np.float64)
        for D, blk in taps.items():
            yv += blk @ xv[:, pad_lo + D:pad_lo + D + Tv]
        for j in range(J):                               # (j, co) at virtual row t' is class sample u = J t' + j
            n = len(range(j, U, J))
            y[:, r + d * j::d * J][:, :n] = yv[j * Co:(j + 1) * Co, :n]
    return y


def mma_clk(N: int) -> float:
    """Measured cost of one tcgen05.mma.kind::f16, M = 128, K = 16, on B200 (csrc/bench_mma.cu)."""
    return max(N / 2.0, (4096 + 32 * N) / 128.0)


def mma_cycles(C: int, k: int, rows: int = 1024):
    """(today, space-to-depth) tensor-pipe cycles for one conv over `rows` time steps of a C-channel stage."""
    J = max(1, 64 // C)
    today = (rows / 128) * k * (C / 16) * mma_clk(C)
    _, steps = toeplitz_weights(np.ones((C, C, k)), J)
    s2d = (rows / (128 * J)) * len(steps) * (C / 16) * mma_clk(J * C)
    return today, s2d
