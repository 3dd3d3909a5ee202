"""Generate tests/golden/rhs/*.cl — OpenCL-C right-hand sides exactly as the REFERENCE's own front ends
emit them (clode/function_converter.py `OpenCLConverter`, clode/xpp_parser.py `convert_xpp_file`), for the
Python functions / XPP file its test-suite uses.  These texts are the input contract of the CUDA shim
(`cl_compat.cuh` must compile them unmodified); the transpilers themselves are out of scope.

Run in the build container (needs /root/reference):   python tests/golden/make_rhs_fixtures.py
"""
import importlib.util
import os
import shutil
import sys
import tempfile
import types

REF = os.environ.get("CLODE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "rhs")


def load(name):
    pkg = sys.modules.setdefault("clode", types.ModuleType("clode"))
    pkg.__path__ = [os.path.join(REF, "clode")]
    spec = importlib.util.spec_from_file_location("clode." + name, os.path.join(REF, "clode", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["clode." + name] = mod
    spec.loader.exec_module(mod)
    return mod


# the Python right-hand sides of the reference tests, verbatim semantics:
#   test/test_features.py:10-21, test/test_aux_values.py:10-24, test/test_vdp.py:33-47, test/test_opencl_builtins.py:65-141
FUNCS = '''
from math import cos
from typing import List
from clode.opencl_builtins import *


def sine_curve(t: float, x_: List[float], p_: List[float], dx_: List[float], aux_: List[float], w_: List[float]) -> None:
    x: float = x_[0]
    dilation: float = p_[0]
    dx: float = cos(t * dilation)
    dx_[0] = dx
    aux_[0] = x + 1.0
    aux_[1] = 1.0
    aux_[2] = -2.0


def vdp(t: float, var: List[float], par: List[float], derivatives: List[float], aux: List[float], wiener: List[float]) -> None:
    mu: float = par[0]
    x: float = var[0]
    y: float = var[1]
    dx: float = y
    dy: float = mu * (1 - x * x) * y - x
    derivatives[0] = dx
    derivatives[1] = dy


def builtins(t: float, x: List[float], p: List[float], dx: List[float], aux: List[float], w: List[float]) -> None:
    p0: int = int(p[0])
    p1: int = int(p[1])
    x0: float = x[0] * t
    x1: float = x[1] * t
    x2: float = x[2] * t
    aux[0] = acos(x0)
    aux[1] = acosh(x0 + 1)
    aux[2] = acospi(x0)
    aux[3] = asin(x0)
    aux[4] = asinh(x0 + 1)
    aux[5] = asinpi(x0)
    aux[6] = atan(x0)
    aux[7] = atan2(x0, x1)
    aux[8] = atan2pi(x0, x1)
    aux[9] = atanh(x0 / 100.0)
    aux[10] = atanpi(x0)
    aux[11] = cbrt(x0)
    aux[12] = ceil(x0)
    aux[13] = copysign(x0, x1)
    aux[14] = cos(x0)
    aux[15] = cosh(x0)
    aux[16] = cospi(x0)
    aux[17] = erf(x0)
    aux[18] = erfc(x0)
    aux[19] = exp(x0)
    aux[20] = exp2(x0)
    aux[21] = exp10(x0)
    aux[22] = expm1(x0)
    aux[23] = fabs(x0)
    aux[24] = fdim(x0, x1)
    aux[25] = floor(x1)
    aux[26] = 0.0
    aux[27] = fmod(x0, x1 + 1)
    aux[28] = heaviside(t - 0.5)
    aux[29] = gamma(x0 + 1)
    aux[30] = hypot(x0, x1)
    aux[31] = float(ilogb(x0 + 1))
    aux[32] = ldexp(x0, int(x1))
    aux[33] = lgamma(x0 + 1)
    aux[34] = log(x0 + 1)
    aux[35] = log1p(x0 + 1)
    aux[36] = log2(x0 + 1)
    aux[37] = log10(x0 + 1)
    aux[38] = 0.0
    aux[39] = 0.0
    aux[40] = 0.0
    aux[41] = nextafter(x0, x1)
    aux[42] = pow(x0, x1)
    aux[43] = pown(x0, p0)
    aux[44] = powr(x0 + 0.5, x1 + 0.2)
    aux[45] = remainder(x0, x1 + 1)
    aux[46] = rint(x0)
    aux[47] = rootn(x0, p1)
    aux[48] = rsqrt(x0 + 1)
    aux[49] = sin(x0)
    aux[50] = sinh(x0)
    aux[51] = sinpi(x0)
    aux[52] = sqrt(x0)
    aux[53] = tan(x0)
    aux[54] = tanh(x0)
    aux[55] = tanpi(x0)
    aux[56] = trunc(x0)
    dx[0] = 0.0
    dx[1] = 0.0
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    load("opencl_builtins")
    fc = load("function_converter")
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "rhs_funcs.py")
    open(path, "w").write(FUNCS)
    spec = importlib.util.spec_from_file_location("rhs_funcs", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rhs_funcs"] = mod
    spec.loader.exec_module(mod)
    for name in ("sine_curve", "vdp", "builtins"):
        conv = fc.OpenCLConverter()
        text = conv.convert_to_opencl(getattr(mod, name), mutable_args=[3, 4], function_name="getRHS")
        open(os.path.join(OUT, f"converted_{name}.cl"), "w").write(text)
        print(name, len(text), "chars")
    # expected aux values of the builtins test, evaluated by the reference's own Python fallbacks
    # (clode/opencl_builtins.py), as test/test_opencl_builtins.py:160-173 does
    import numpy as np
    expected = np.zeros((3, 57))
    for k, t in enumerate([0.0, 0.5, 1.0]):
        aux = [0.0] * 57
        mod.builtins(t, [0.3, 1.5, 2.5], [2.8, 4.55], [0.0, 0.0, 0.0], aux, [])
        expected[k] = aux
    np.save(os.path.join(OUT, "builtins_expected.npy"), expected)
    print("builtins expected", expected.shape)
    # XPP -> OpenCL (clode/xpp_parser.py:73-190; reference fixture test/xpp/van_der_pol_oscillator.xpp)
    xp = load("xpp_parser")
    xpp = os.path.join(tmp, "van_der_pol_oscillator.xpp")
    shutil.copy(os.path.join(REF, "test", "xpp", "van_der_pol_oscillator.xpp"), xpp)
    out = xp.convert_xpp_file(xpp)
    shutil.copy(out, os.path.join(OUT, "converted_xpp_van_der_pol.cl"))
    print("xpp ->", os.path.basename(out))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
