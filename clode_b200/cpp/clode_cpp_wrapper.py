"""`clode_b200.cpp.clode_cpp_wrapper` — same import path shape as the reference's
`clode.cpp.clode_cpp_wrapper` (clode/runtime.py:5-15); re-exports the compiled pybind11 module that
lives next to libclode_rt.so.  Importing fails loudly if the extension has not been built."""
import importlib.util
import os
import sys

_pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    import sysconfig

    path = os.path.join(_pkg, "clode_cpp_wrapper" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m clode_b200.build`")
    spec = importlib.util.spec_from_file_location("clode_cpp_wrapper", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_m = _load()
for _k in dir(_m):
    if not _k.startswith("__"):
        globals()[_k] = getattr(_m, _k)
sys.modules.setdefault("clode_cpp_wrapper", _m)
