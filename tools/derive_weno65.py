"""Derive the difference form of the WENO65 smoothness indicators used by the production arithmetic
(loki_b200/csrc/lk_device.cuh).  The reference's Maple-generated quadratic forms bl, br
(KineticSpeciesF.f:936-950) vanish on constants, so each is a quadratic form in the four first
differences of its five cells; an exact (rational) LDL^T factorisation writes it as a sum of four squares:
    bl = sum_k dk * (D_k + sum_{j>k} l_kj D_j)^2 .
br is the mirror image of bl about the face, so the same constants serve both.  Prints the constants as
C literals and checks the identity against the original formula in exact arithmetic."""
from fractions import Fraction as F
import random

names = ["um3", "um2", "um1", "u0", "up1"]
idx = {n: i for i, n in enumerate(names)}
M = [[F(0)] * 5 for _ in range(5)]


def sq(n, c):
    M[idx[n]][idx[n]] += c


def cr(a, b, c):  # c * a * b, symmetrised
    M[idx[a]][idx[b]] += c / 2
    M[idx[b]][idx[a]] += c / 2


k = F(1, 30240)
sq("um1", F(5489, 105))
cr("u0", "um1", -2242428 * k); cr("um2", "um1", -1887108 * k); cr("um3", "um1", 410226 * k); cr("up1", "um1", 557646 * k)
sq("um2", F(75329, 3780))
cr("u0", "um2", 1259696 * k); cr("um3", "um2", -275318 * k); cr("up1", "um2", -302534 * k)
sq("um3", F(33727, 30240))
cr("u0", "um3", -264314 * k); cr("up1", "um3", 61952 * k)
sq("u0", F(106409, 3780))
cr("u0", "up1", -F(227749, 15120))
sq("up1", F(69217, 30240))

# translation invariance: M * ones = 0
assert all(sum(row) == 0 for row in M), [sum(r) for r in M]
# u = T D with u[0] = 0: u_k = sum_{m<k} D_m
T = [[F(1) if m < kk else F(0) for m in range(4)] for kk in range(5)]
N = [[sum(T[a][i] * M[a][b] * T[b][j] for a in range(5) for b in range(5)) for j in range(4)] for i in range(4)]
# LDL^T, unit lower L (here used as rows of an upper factor: s_k = D_k + sum_{j>k} l_kj D_j)
n = 4
A = [row[:] for row in N]
d = [F(0)] * n
L = [[F(0)] * n for _ in range(n)]
for kk in range(n):
    d[kk] = A[kk][kk]
    assert d[kk] > 0
    for j in range(kk + 1, n):
        L[kk][j] = A[kk][j] / d[kk]
    for i in range(kk + 1, n):
        for j in range(kk + 1, n):
            A[i][j] -= L[kk][i] * L[kk][j] * d[kk]


def q_orig(u):
    return sum(M[i][j] * u[i] * u[j] for i in range(5) for j in range(5))


def q_new(u):
    D = [u[i + 1] - u[i] for i in range(4)]
    tot = F(0)
    for kk in range(4):
        s = D[kk] + sum(L[kk][j] * D[j] for j in range(kk + 1, 4))
        tot += d[kk] * s * s
    return tot


random.seed(1)
for _ in range(20):
    u = [F(random.randint(-1000, 1000), random.randint(1, 97)) for _ in range(5)]
    assert q_orig(u) == q_new(u)
print("// bl = d0*(s0^2 + r1*s1^2 + r2*s2^2 + r3*s3^2),  s_k = D_k + sum_{j>k} l_kj D_j ; D_k = u_{k+1}-u_k over (um3..up1)")
print("// exact rationals:")
print("//   d0 =", d[0])
for kk in range(4):
    for j in range(kk + 1, 4):
        print("//   l%d%d = %s" % (kk, j, L[kk][j]))
for kk in range(1, 4):
    print("//   r%d = %s" % (kk, d[kk] / d[0]))
print("constexpr double W65_D0 = %.17g;" % float(d[0]))
for kk in range(4):
    for j in range(kk + 1, 4):
        print("constexpr double W65_L%d%d = %.17g;" % (kk, j, float(L[kk][j])))
for kk in range(1, 4):
    print("constexpr double W65_R%d = %.17g;" % (kk, float(d[kk] / d[0])))
print("constexpr double W65_EPS = %.17g;  // 1e-10 / d0" % (1e-10 / float(d[0])))
