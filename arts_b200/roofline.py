"""Algorithmic work counts of the path — the numerators of the roofline fractions.

Stage 1 is FP64-pipe bound: FLOPs per (line, frequency, level) evaluation are fixed by the
*reference's* own region map of Faddeeva::w (3rdparty/Faddeeva/Faddeeva.cc:689-741,786,890),
never by the instructions our kernels execute:

    R1  x+y > 1e7            17 flop     (w = i/sqrt(pi)/z, :708-720, + z, s*w, accumulate)
    R2  4000 < x+y <= 1e7    28 flop     (nu = 2 closed form, :721-725)
    R3  continued fraction   8 nu + 10   (nu = floor(3.9 + 11.398/(0.08254x+0.1421y+0.2023)), :729-741)
    R4  series, x < 10       350 flop    (:786-889)
    R5  x >= 10, tiny y      310 flop    (:890-931)

Stage 2 is HBM bound: 56 B of K read per (frequency, level) step + 32 B of spectral_rad written
per frequency (+ 88 B per Jacobian target and step).
"""
from __future__ import annotations

import numpy as np

FLOP_R1, FLOP_R2, FLOP_R4, FLOP_R5 = 17.0, 28.0, 350.0, 310.0
REGIONS = ("R1_x+y>1e7", "R2_far_wing", "R3_continued_fraction", "R4_series", "R5_x>=10")


def flops_per_eval(hist) -> tuple[float, dict]:
    """``hist`` = the 8 numbers of ab200_path_region_histogram.  Returns (mean algorithmic FLOP per
    evaluated (line, frequency, level) pair, {region: fraction})."""
    h = np.asarray(hist, dtype=float)
    n = h[:5].sum()
    if n <= 0:
        return 0.0, {r: 0.0 for r in REGIONS}
    total = FLOP_R1 * h[0] + FLOP_R2 * h[1] + (8.0 * h[5] + 10.0 * h[2]) + FLOP_R4 * h[3] + FLOP_R5 * h[4]
    return float(total / n), {r: float(h[i] / n) for i, r in enumerate(REGIONS)}


def stokes_bytes_per_step(np_: int, nq: int = 0) -> float:
    """Algorithmic HBM bytes per (frequency, level) step of the Stokes chain."""
    return 56.0 + 32.0 / max(np_, 1) + 88.0 * nq
