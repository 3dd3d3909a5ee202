"""clode_b200 — B200-native ensemble ODE engine behind the clODE API.

The Python-visible names are those of the reference package (clode/__init__.py:87-168) for the hot
path: `Simulator`, `FeatureSimulator`, `TrajectorySimulator`, `Stepper`, `Observer`, `ObserverOutput`,
`TrajectoryOutput`, the parameter structs and the runtime helpers.  Importing the front end needs the
compiled extension (`python -m clode_b200.build`); the low-level ctypes binding `clode_b200._rt` and
the build helpers import without it.
"""
__version__ = "0.1.0"

_FRONT_END = ["CLDeviceType", "CLVendor", "DeviceInfo", "PlatformInfo", "OpenCLResource", "initialize_runtime",
              "print_opencl", "query_opencl", "LogLevel", "get_log_level", "set_log_level", "set_log_pattern",
              "ProblemInfo", "SolverParams", "Stepper", "Simulator", "FeatureSimulator", "Observer", "ObserverParams",
              "ObserverOutput", "TrajectorySimulator", "TrajectoryOutput"]
_TEXT_FRONT_ENDS = {"OpenCLConverter": "function_converter", "OpenCLRhsEquation": "function_converter",
                    "convert_str_to_opencl": "function_converter", "convert_xpp_file": "xpp_parser",
                    "read_ode_parameters": "xpp_parser", "format_opencl_rhs": "xpp_parser"}
# Python stand-ins of the OpenCL-C builtins, so that a right-hand side written as a Python function can call
# `clode.exp`, `clode.heaviside`, `clode.pown`, ... and still run as plain Python (clode/__init__.py:13-67)
_BUILTINS = ["acos", "acosh", "acospi", "asin", "asinh", "asinpi", "atan", "atan2", "atan2pi", "atanh", "atanpi", "cbrt",
             "ceil", "copysign", "cos", "cosh", "cospi", "erf", "erfc", "exp", "exp10", "exp2", "expm1", "fabs", "fdim",
             "floor", "fmod", "gamma", "heaviside", "hypot", "ilogb", "ldexp", "lgamma", "log", "log10", "log1p", "log2",
             "nextafter", "pow", "pown", "powr", "remainder", "rint", "rootn", "rsqrt", "sin", "sinh", "sinpi", "sqrt",
             "tan", "tanh", "tanpi", "trunc"]
__all__ = list(_FRONT_END) + list(_TEXT_FRONT_ENDS) + _BUILTINS


def __getattr__(name):
    # lazy: `import clode_b200` must work before the extension is built (build.py lives in this package)
    if name in _TEXT_FRONT_ENDS:  # pure Python: Python / XPP -> OpenCL-C source (clode/__init__.py exports the same names)
        import importlib
        return getattr(importlib.import_module("." + _TEXT_FRONT_ENDS[name], __name__), name)
    if name in _BUILTINS:
        from . import opencl_builtins
        return getattr(opencl_builtins, name)
    if name in _FRONT_END:
        from . import features, runtime, solver, trajectory
        from .cpp import clode_cpp_wrapper as w

        table = {"ProblemInfo": w.ProblemInfo, "SolverParams": w.SolverParams, "ObserverParams": w.ObserverParams,
                 "Stepper": solver.Stepper, "Simulator": solver.Simulator, "FeatureSimulator": features.FeatureSimulator,
                 "Observer": features.Observer, "ObserverOutput": features.ObserverOutput,
                 "TrajectorySimulator": trajectory.TrajectorySimulator, "TrajectoryOutput": trajectory.TrajectoryOutput}
        if name in table:
            return table[name]
        return getattr(runtime, name)
    raise AttributeError(name)
