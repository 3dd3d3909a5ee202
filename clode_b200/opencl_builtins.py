"""Python stand-ins for the OpenCL-C math functions a right-hand side may call, so that a function written for
`rhs_equation=` also runs (and type-checks) as plain Python.  Same names as the reference's clode/opencl_builtins.py;
where that module approximates (`nextafter`, `ilogb`, `rint`, `rootn` of negative numbers) these follow the OpenCL C
definitions the device code implements (OpenCL C 1.2 §6.12.2), so Python and device evaluations agree.
"""
from __future__ import annotations

import math
from math import (acos, acosh, asin, asinh, atan, atan2, atanh, ceil, copysign, cos, cosh, erf, erfc, exp, expm1,  # noqa: F401
                  fabs, floor, fmod, gamma, hypot, ldexp, lgamma, log, log1p, log2, log10, pi, pow, remainder, sin, sinh,
                  sqrt, tan, tanh, trunc)


def acospi(x: float) -> float:
    return acos(x) / pi


def asinpi(x: float) -> float:
    return asin(x) / pi


def atanpi(x: float) -> float:
    return atan(x) / pi


def atan2pi(y: float, x: float) -> float:
    return atan2(y, x) / pi


def cbrt(x: float) -> float:
    return copysign(abs(x) ** (1.0 / 3.0), x)


def cospi(x: float) -> float:
    return cos(pi * x)


def sinpi(x: float) -> float:
    return sin(pi * x)


def tanpi(x: float) -> float:
    return tan(pi * x)


def exp2(x: float) -> float:
    return 2.0 ** x


def exp10(x: float) -> float:
    return 10.0 ** x


def fdim(x: float, y: float) -> float:
    return x - y if x > y else 0.0


def heaviside(x: float) -> float:
    """clODE's own helper (clode/cpp/clODE_utilities.cl): 1 for x >= 0, else 0"""
    return 1.0 if x >= 0.0 else 0.0


def ilogb(x: float) -> int:
    """unbiased binary exponent of x"""
    return math.frexp(x)[1] - 1


def nextafter(x: float, y: float) -> float:
    return math.nextafter(x, y)


def pown(x: float, n: int) -> float:
    return x ** int(n)


def powr(x: float, y: float) -> float:
    """x >= 0"""
    return x ** y


def rint(x: float) -> float:
    """round to nearest, ties to even"""
    return float(round(x))


def rootn(x: float, n: int) -> float:
    n = int(n)
    if x < 0 and n % 2:
        return -((-x) ** (1.0 / n))
    return x ** (1.0 / n)


def rsqrt(x: float) -> float:
    return 1.0 / sqrt(x)


__all__ = ["acos", "acosh", "acospi", "asin", "asinh", "asinpi", "atan", "atan2", "atan2pi", "atanh", "atanpi", "cbrt",
           "ceil", "copysign", "cos", "cosh", "cospi", "erf", "erfc", "exp", "exp2", "exp10", "expm1", "fabs", "fdim", "floor",
           "fmod", "gamma", "heaviside", "hypot", "ilogb", "ldexp", "lgamma", "log", "log1p", "log2", "log10", "nextafter",
           "pow", "pown", "powr", "remainder", "rint", "rootn", "rsqrt", "sin", "sinh", "sinpi", "sqrt", "tan", "tanh", "tanpi",
           "trunc"]
