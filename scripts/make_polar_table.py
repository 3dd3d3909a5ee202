#!/usr/bin/env python
"""The 256-entry table of clode_b200/csrc/device/fast_polar.cuh: for the mantissa interval j of q (top 8 mantissa bits;
the upper half of a binade is folded down, m = q 2^-e' in [0.75, 1.5)) the reciprocal of the interval's centre, rounded to
double, and -2 log of THAT rounded reciprocal as hi + lo (mpmath, 300 bits).  Entry 255 (m in [1 - 1/512, 1)) is {1, 0, 0}:
next to 1 the logarithm is the polynomial alone, with full relative accuracy.
usage: python scripts/make_polar_table.py > /tmp/table.txt"""
import mpmath as mp

mp.mp.prec = 300
rows = []
for j in range(256):
    c = mp.mpf(1) + (mp.mpf(j) + mp.mpf(1) / 2) / 256
    if j >= 128:
        c /= 2
    rc = float(1 / c)
    if j == 255:
        rc, hi, lo = 1.0, 0.0, 0.0
    else:
        L = -2 * (-mp.log(mp.mpf(rc)))      # -2 * log(c_eff), c_eff = 1 / rc:  log(m) = log(c_eff) + log1p(m rc - 1)
        L = 2 * mp.log(mp.mpf(rc))          # = -2 log(c_eff)
        hi = float(L)
        lo = float(L - mp.mpf(hi))
    rows.append("{%s, %s, %s}" % (rc.hex(), hi.hex(), lo.hex()))
for k in range(0, 256, 2):
    print("    " + ", ".join(rows[k:k + 2]) + ",")
ln2 = mp.log(2)
# -2 ln2 as hi (42 significant bits: e' * hi is exact for |e'| < 2^11) + lo
m2 = -2 * ln2
hi = mp.mpf(int(m2 * mp.mpf(2) ** 41)) / mp.mpf(2) ** 41
print("// -2 ln2 hi", float(hi).hex(), "lo", float(m2 - hi).hex())
