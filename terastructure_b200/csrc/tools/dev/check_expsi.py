"""Host model of f_expsi (ts_persist.cuh) against mpmath: max relative error over a log grid,
with the reciprocal u = 1/xs perturbed by the worst case a single Newton step can leave."""
import numpy as np, mpmath as mp
mp.mp.dps = 40
C = [float.fromhex(h) for h in (
    "0x1.5555555555555p-5", "0x1.5555555555555p-6", "0x1.05b05b05b05b0p-8", "-0x1.2222222222222p-8",
    "-0x1.c7f80db9bf2a3p-9", "0x1.1a4cc13ddafa2p-9", "0x1.05f536517fa45p-8", "-0x1.1e6ee98a17aecp-9",
    "-0x1.e5f884ccda9f9p-8", "0x1.fe414efb9852ap-9", "0x1.54c7f9f55e0ebp-6", "-0x1.5f836e8779d89p-7",
    "-0x1.51ea52a4cdfabp-4", "0x1.59488e35cad4dp-5", "0x1.c276c25d1fbddp-2")]

def f_model(x, du):
    small = x < 8.0
    xs = np.where(small, x + 8.0, x)
    u = (1.0 / xs) * (1.0 + du)
    q = np.zeros_like(u)
    for c in C[::-1]:
        q = q * u + c
    f = u * q + (xs - 0.5)
    x2 = x * x
    pe = (((x2 + 322.0) * x2 + 6769.0) * x2 + 13068.0) * x2
    po = ((28.0 * x2 + 1960.0) * x2 + 13132.0) * x2 + 5040.0
    de = ((196.0 * x2 + 9800.0) * x2 + 39396.0) * x2 + 5040.0
    do = ((8.0 * x2 + 1932.0) * x2 + 27076.0) * x2 + 26136.0
    r = (do * x + de) / (po * x + pe)
    return np.where(small, f * np.exp(-r), f)

xs = np.exp(np.linspace(np.log(0.03), np.log(1e7), 4001))
ref = np.array([float(mp.exp(mp.digamma(mp.mpf(float(x))))) for x in xs])
for du in (0.0, 2.0 ** -38, -2.0 ** -38, 2.0 ** -34):
    err = np.abs(f_model(xs, du) / ref - 1.0)
    print(f"du={du:.2e}: max rel err {err.max():.2e} at x={xs[err.argmax()]:.3g}")
