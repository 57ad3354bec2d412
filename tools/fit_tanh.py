"""Fits the odd polynomial used by scone_tanh (csrc/scone_slab.cuh) on |x| < 0.55:
tanh(x) ~= x + x^3 * P(x^2), P of degree 4, least squares on Chebyshev nodes (near-minimax); prints the coefficients
and the float32-evaluated error."""
import numpy as np
from numpy.polynomial import chebyshev as ch

lim = 0.55
umax = lim * lim
k = np.arange(4000)
u = 0.5 * umax * (1 + np.cos(np.pi * (k + 0.5) / len(k)))
x = np.sqrt(u)
f = (np.tanh(x) - x) / x ** 3
deg = 4
c = np.polynomial.polynomial.polyfit(u, f, deg, w=x ** 3)     # weight: absolute error of the final result
print('coefficients (low -> high):', ['%.10e' % v for v in c])
xs = np.linspace(-lim, lim, 2000001).astype(np.float32)
x2 = xs * xs
p = np.float32(c[4])
for j in (3, 2, 1, 0):
    p = (p * x2 + np.float32(c[j])).astype(np.float32)
y = ((xs * x2) * p + xs).astype(np.float32)
t = np.tanh(xs.astype(np.float64))
err = np.abs(y - t)
nz = np.abs(t) > 0
print('max abs err %.3g, max rel err %.3g' % (err.max(), (err[nz] / np.abs(t[nz])).max()))
