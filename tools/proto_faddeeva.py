"""numpy prototype of the device w(z) (arts_b200/csrc/faddeeva.cuh) used to choose the
window width and check accuracy against scipy.special.wofz (= the reference's Faddeeva
package) before writing CUDA.  Not part of the product or the tests."""
import numpy as np
from scipy.special import wofz, erfcx

A = 0.518321480430085929872
A2 = 0.268657157075235951582
CC = 0.329973702884629072537  # 2a/pi
ISPI = 0.56418958354775628694807945156


def cf_region(x, y):
    return (y > 7) | ((x > 6) & ((y > 0.1) | ((x > 8) & (y > 1e-10)) | (x > 28)))


def e1_of_y(y, nterm=13):
    n = np.arange(1, nterm + 1)[:, None]
    s1 = (np.exp(-A2 * n * n) / (A2 * n * n + y[None, :] ** 2)).sum(0)
    return erfcx(y) - CC * y * s1


def w_series(x, y, W=13):
    """x >= 0, y >= 0"""
    e1 = e1_of_y(y)
    expx2 = np.exp(-x * x)
    th = x * y
    s, c = np.sin(th), np.cos(th)
    s2 = 2 * s * c
    c2 = 1 - 2 * s * s
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(th < 1e-4, 1 - th * th / 6, s / th)
    re = expx2 * (e1 * c2 + CC * x * s * sinc)
    im = expx2 * (CC * x * c * sinc - e1 * s2)
    n0 = np.floor(x / A + 0.5)
    delta = x - A * n0
    g0 = np.exp(-delta * delta)
    p = np.exp(2 * A * delta)
    m = 1 / p
    y2 = y * y
    sr = np.zeros_like(x)
    si = np.zeros_like(x)
    # centre
    t = A * n0
    valid = n0 != 0
    den = np.where(valid, t * t + y2, 1.0)
    v = np.where(valid, g0 / den, 0.0)
    sr += v
    si += t * v
    pk = np.ones_like(x)
    mk = np.ones_like(x)
    for k in range(1, W + 1):
        tk = np.exp(-A2 * k * k)
        pk = pk * p
        mk = mk * m
        ep = g0 * tk * pk
        em = g0 * tk * mk
        tp = A * (n0 + k)
        tm = A * (n0 - k)
        v = ep / (tp * tp + y2)
        sr += v
        si += tp * v
        valid = (n0 - k) != 0
        den = np.where(valid, tm * tm + y2, 1.0)
        v = np.where(valid, em / den, 0.0)
        sr += v
        si += tm * v
    re = re + 0.5 * CC * y * sr
    im = im + 0.5 * CC * si
    return re + 1j * im


def w_cf(x, y):
    nu = np.floor(3.9 + 11.398 / (0.08254 * x + 0.1421 * y + 0.2023))
    wr, wi = x.copy(), y.copy()
    k = 0.5 * (nu - 1)
    while True:
        act = k > 0.4
        if not act.any():
            break
        den = np.where(act, k / (wr * wr + wi * wi), 0.0)
        wr = np.where(act, x - wr * den, wr)
        wi = np.where(act, y + wi * den, wi)
        k = k - 0.5
    den = ISPI / (wr * wr + wi * wi)
    return den * wi + 1j * den * wr


def w_far(x, y):
    q = x * x
    c1 = y * y + 0.5
    c3 = y * y - 0.5
    dr = q - c1
    d2 = dr * dr + 4 * y * y * q
    return ISPI * (y * (q + c1) / d2 + 1j * x * (q + c3) / d2)


def w_dev(z, W=13):
    x = np.abs(z.real)
    y = z.imag
    out = np.empty_like(z)
    far = x + y > 4000
    cf = cf_region(x, y) & ~far
    ser = ~far & ~cf
    out[far] = w_far(x[far], y[far])
    out[cf] = w_cf(x[cf], y[cf])
    out[ser] = w_series(x[ser], y[ser], W)
    return np.where(z.real < 0, np.conj(out), out), far, cf, ser


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    n = 400000
    x = np.concatenate([rng.uniform(0, 30, n), 10 ** rng.uniform(-8, 4.2, n)])
    y = np.concatenate([10 ** rng.uniform(-14, 1, n), 10 ** rng.uniform(-6, 4, n)])
    y[:1000] = 0.0
    z = x + 1j * y
    ref = wofz(z)
    for W in (11, 12, 13, 14):
        got, far, cf, ser = w_dev(z, W)
        for name, msk in (("far", far), ("cf", cf), ("series", ser)):
            er = np.abs(got.real[msk] - ref.real[msk]) / np.abs(ref.real[msk])
            ei = np.abs(got.imag[msk] - ref.imag[msk]) / np.maximum(np.abs(ref.imag[msk]), 1e-300)
            ea = np.abs(got[msk] - ref[msk]) / np.abs(ref[msk])
            print(f"W={W} {name:6s} n={msk.sum():7d} max rel re {er.max():.2e} im {ei.max():.2e} |w| {ea.max():.2e}")
            if name == "series":
                i = np.argmax(ei)
                print("   worst im at", z[msk][i], got[msk][i], ref[msk][i])
                i = np.argmax(er)
                print("   worst re at", z[msk][i], got[msk][i], ref[msk][i])
