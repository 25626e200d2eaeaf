"""Generates the coefficients of airice_atan_q (airice_math.cuh): q(u) with atan(r) = r + r u q(u), u = r^2,
|r| <= tan(pi/8), by interpolation at Chebyshev nodes in 50-digit arithmetic (near-minimax); prints the C array and
the maximum relative error of the double-rounded polynomial."""
import mpmath as mp
mp.mp.dps = 50
B = mp.tan(mp.pi / 8) ** 2
N = 10


def q(u):
    if u == 0:
        return -mp.mpf(1) / 3
    s = mp.sqrt(u)
    return (mp.atan(s) / s - 1) / u


nodes = [B / 2 + B / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (N + 1))) for k in range(N + 1)]
A = mp.matrix(N + 1, N + 1)
y = mp.matrix(N + 1, 1)
for i, x in enumerate(nodes):
    for j in range(N + 1):
        A[i, j] = x ** j
    y[i] = q(x)
c = [float(v) for v in mp.lu_solve(A, y)]
err = 0
for i in range(4001):
    u = B * i / 4000
    p = mp.mpf(0)
    for k in reversed(c):
        p = p * u + mp.mpf(k)
    r = mp.sqrt(u)
    if r > 0:
        err = max(err, abs((r + r * u * p) - mp.atan(r)) / mp.atan(r))
print("// max relative error of atan(r) with these (double-rounded) coefficients: %s" % mp.nstr(err, 3))
print(", ".join(float.hex(v) for v in c))
