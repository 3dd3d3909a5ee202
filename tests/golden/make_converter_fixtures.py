"""Generate tests/golden/converter_cases.json — what the REFERENCE's Python->OpenCL converter
(clode/function_converter.py `OpenCLConverter`) emits, or raises, for a set of Python sources.  The product's own
converter (clode_b200/function_converter.py) must reproduce the texts character for character and the error types.

Run in the build container (needs /root/reference):   python tests/golden/make_converter_fixtures.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_rhs_fixtures import load  # noqa: E402  (imports the reference modules without its compiled extension)

SIG = "t: float, x: List[float], p: List[float], dx: List[float], aux: List[float], w: List[float]"

# each case: a list of conversion calls made on ONE converter object (helpers first), the text of the last call is kept
CASES = {
    # the reference's own expectations, test/test_function_converter.py:77-181
    "ref_add_float": [dict(src="def add_float(a: float, b: float) -> float:\n    res: float = a + b\n    return res\n")],
    "ref_list_adder": [dict(src="def add_floats_in_list(lst_in: List[float], lst_out: List[float]) -> None:\n"
                                "    res: float = lst_in[0] + lst_in[1]\n    lst_out[0] = res\n")],
    "ref_all_operations": [dict(src="""def all_operations(a: float, b: float, c: int, d: int) -> float:
    res1: float = a + b * c / d
    res2: float = ((a - b) % c) ** d
    res3: float = a**b
    res4: float = a**1
    res5: float = a**2
    res6: float = a**3
    res7: float = a**4
    res8: float = a**5
    res9: float = a**0
    res10: float = a**0.5
    sum_res: float = (
        res1 + res2 + -res3 + res4 + res5 + res6 + res7 + res8 + res9 + res10
    )
    return res1 + sum_res
""")],
    "ref_second_function": [
        dict(src="def add_float(a: float, b: float) -> float:\n    res: float = a + b\n    return res\n"),
        dict(src="def get_rhs(var: List[float], derivatives: List[float]) -> None:\n"
                 "    res: float = add_float(var[0], var[1])\n    derivatives[0] = res\n")],
    # right-hand sides as clode/solver.py:239-245 converts them
    "rhs_lorenz_with_helper": [
        dict(src="def coupling(a: float, b: float, s: float) -> float:\n    return s * (b - a)\n"),
        dict(src=f"""def lorenz({SIG}) -> None:
    r: float = p[0]
    s: float = p[1]
    b: float = p[2]
    dx[0] = coupling(x[0], x[1], s)
    dx[1] = x[0] * (r - x[2]) - x[1]
    dx[2] = x[0] * x[1] - b * x[2]
    aux[0] = dx[0]
""", mutable_args=[3, 4], function_name="getRHS")],
    "rhs_module_calls_and_casts": [dict(src=f"""def f({SIG}) -> None:
    k: int = int(p[0])
    g: float = float(k) + math.exp(-x[0]) + abs(x[1]) + max(x[0], x[1]) + min(x[0], 2.5) + mod(x[0], 3.0)
    h: float
    h = gamma(x[0] + 1) ** k + (-g) ** 2.0 + g ** -1 + pown(g, k)
    n: int = k * 2 + 1
    dx[0] = g * n - h / 3
    dx[1] = float(n % 2)
    aux[0] = rootn(g, k) + ldexp(g, k)
""", mutable_args=[3, 4], function_name="getRHS")],
    "int_declared_from_int_value": [dict(src="def q(a: float) -> float:\n    one: float = 1\n    two: float = one + 1\n    return a * two\n")],
    "optional_and_int_lists": [dict(src="def q(a: float | None, n: List[int], out: List[float]) -> None:\n"
                                        "    out[0] = a * n[1]\n")],
    "mutable_args_by_name": [dict(src="def q(a: List[float], b: List[float], c: List[float]) -> None:\n    b[0] = a[0]\n",
                                  mutable_args=["c"])],
    "returns_int": [dict(src="def q(a: int, b: int) -> int:\n    return a * b - 3\n")],
    "noise_and_time": [dict(src=f"""def f({SIG}) -> None:
    dx[0] = -p[0] * x[0] + p[1] * w[0] + sin(2 * 3.141592653589793 * t)
""", mutable_args=[3, 4], function_name="getRHS")],
    # failures
    "err_unsupported_type": [dict(src="def unsupported_type(x: str) -> float:\n    return x\n")],
    "err_no_annotation": [dict(src="def no_annotation(x) -> float:\n    return x\n")],
    "err_change_type": [dict(src="def change_variable_type() -> int:\n    a: int = 1\n    a = 2.2\n    return a\n")],
    "err_no_return_type": [dict(src="def no_return_type():\n    a: int = 1\n    return a\n")],
    "err_return_str": [dict(src="def return_type_str() -> str:\n    a: int = 1\n    return a\n")],
    "err_variable_str": [dict(src="def variable_string() -> int:\n    a: str = 2\n    return 1\n")],
    "err_tuple_assign": [dict(src="def tuple_assign() -> int:\n    a, b = 1, 2\n    return a + b\n")],
    "err_redeclare": [dict(src="def redeclare_variable() -> int:\n    a: int = 1\n    a: int = 2\n    return a\n")],
    "err_unknown_variable": [dict(src="def q(a: float) -> float:\n    return a + b\n")],
    "err_unknown_function": [dict(src="def q(a: float) -> float:\n    return frobnicate(a)\n")],
    "err_unknown_module_function": [dict(src="def q(a: float) -> float:\n    return math.frobnicate(a)\n")],
    "err_builtin_arg_count": [dict(src="def q(a: float) -> float:\n    return atan2(a)\n")],
    "err_floor_division": [dict(src="def q(a: float) -> float:\n    return a // 2\n")],
    "err_unary_invert": [dict(src="def q(a: int) -> int:\n    return ~a\n")],
    "err_variable_index": [dict(src="def q(a: List[float], i: int) -> float:\n    return a[i]\n")],
    "err_if_statement": [dict(src="def q(a: float) -> float:\n    if a > 0:\n        return a\n    return -a\n")],
    "err_aug_assign": [dict(src="def q(a: float) -> float:\n    b: float = a\n    b += 1.0\n    return b\n")],
    "err_int_into_real_array": [dict(src="def q(dx: List[float]) -> None:\n    dx[0] = 1\n")],
    "err_undeclared_assign": [dict(src="def q(a: float) -> float:\n    b = a\n    return b\n")],
    "err_real_into_int": [dict(src="def q(a: float) -> int:\n    b: int = a\n    return b\n")],
    "err_string_constant": [dict(src="def q(a: float) -> float:\n    return a + 'x'\n")],
    "err_list_of_str": [dict(src="def q(a: List[str]) -> float:\n    return 1.0\n")],
    "err_dict_annotation": [dict(src="def q(a: Dict[float]) -> float:\n    return 1.0\n")],
    "err_mutable_args_mixed": [dict(src="def q(a: List[float]) -> None:\n    a[0] = 1.0\n", mutable_args=[0, "a"])],
}


def main():
    fc = load("function_converter")
    out = {}
    for name, calls in CASES.items():
        conv = fc.OpenCLConverter()
        entry = {"calls": calls}
        try:
            text = None
            for c in calls:
                text = conv.convert_to_opencl(c["src"], mutable_args=c.get("mutable_args"), function_name=c.get("function_name"))
            entry["text"] = text
        except Exception as e:  # noqa: BLE001 — the kind of failure is part of the contract
            entry["error"] = {"type": type(e).__name__, "message": str(e)}
        out[name] = entry
    path = os.path.join(HERE, "converter_cases.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    ok = sum("text" in v for v in out.values())
    print(f"wrote {path}: {ok} texts, {len(out) - ok} failures")
    # XPP front end (clode/xpp_parser.py): the reference's text for the .xpp inputs under tests/golden/xpp
    import glob
    import shutil
    import tempfile
    xp = load("xpp_parser")
    tmp = tempfile.mkdtemp()
    for src in sorted(glob.glob(os.path.join(HERE, "xpp", "*.xpp"))):
        work = os.path.join(tmp, os.path.basename(src))
        shutil.copy(src, work)
        made = xp.convert_xpp_file(work)
        shutil.copy(made, src[:-4] + "_reference.cl")
        print("xpp ->", os.path.basename(src)[:-4] + "_reference.cl")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
