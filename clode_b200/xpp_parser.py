"""XPP (.ode-style) model files -> OpenCL-C right-hand sides: the front end behind `src_file="model.xpp"`.

Same contract and, by default, the same text as the reference's clode/xpp_parser.py (`convert_xpp_file`, :169-190;
the expected text is the reference's own test/xpp/van_der_pol_oscillator_reference.cl, committed as a golden fixture).
Line kinds (xpp_parser.py:49-72): `par` / `p` parameter lists, `init` initial values (these name and order the state
variables), `aux` named auxiliary outputs, `wiener` noise terms, `x' = ...` differential equations, `@` options
(ignored), anything else is copied into the body as a statement; `%` starts a comment.

Implementation: a small model record (`XppModel`) filled by `parse_xpp` and printed by `render`, instead of the
reference's dictionaries-through-regexes pipeline.

`precise_literals=True` (not in the reference) writes floating literals as `RCONST(1.5)` — a real double
in double-precision builds — instead of the reference's forced single-precision `1.5f`, keeps exponents intact
(`2.5e-3`, which the reference turns into the invalid `2.5fe-3`) and converts real powers before integer ones
(`v^0.5`, which the reference turns into the invalid `pown(v, 0).5`).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

_KEYWORD = re.compile(r"(par|p|init|aux) ")
_NAME_EQ = re.compile(r"(\w+)(\s*=\s*)")
_DERIVATIVE = re.compile(r"\w+'\s*=\s*")


@dataclass
class XppModel:
    parameters: Dict[str, str] = field(default_factory=dict)
    auxiliaries: Dict[str, str] = field(default_factory=dict)
    initial_values: Dict[str, str] = field(default_factory=dict)
    derivatives: Dict[str, str] = field(default_factory=dict)
    noise: List[str] = field(default_factory=list)
    statements: List[str] = field(default_factory=list)


def _definitions(line: str) -> List[Tuple[str, str]]:
    """`par a=1, b = 2` -> [("a", "1,"), ("b", "2")]: a value runs up to the next `name =`"""
    rest = "=".join(part.strip() for part in _KEYWORD.sub("", line).split("="))
    pairs = []
    while rest:
        head = _NAME_EQ.search(rest)
        if not head:
            raise ValueError(f"Could not parse line '{line}'")
        following = _NAME_EQ.search(rest, head.end())
        stop = following.start() if following else len(rest)
        pairs.append((head.group(1), rest[head.end():stop]))
        rest = rest[stop:]
    return pairs


def parse_xpp(text: str) -> XppModel:
    m = XppModel()
    for line in text.splitlines():
        if line.startswith(("par ", "p ")):
            target = m.parameters
        elif line.startswith("aux "):
            target = m.auxiliaries
        elif line.startswith("init "):
            target = m.initial_values
        else:
            if line.startswith("wiener "):
                m.noise.append(line[7:])
            elif _DERIVATIVE.match(line):
                m.derivatives[line.split("'")[0]] = _DERIVATIVE.sub("", line)
            elif line.startswith("@") or not line.strip():
                pass
            else:
                m.statements.append(line)
            continue
        for name, value in _definitions(line):
            # auxiliary expressions are printed verbatim; parameter / initial values are only read back by callers
            target[name] = value.rstrip(",") if target is m.auxiliaries else value.strip().rstrip(",").strip()
    return m


def render(m: XppModel, precise_literals: bool = False) -> str:
    out = ["void getRHS(const realtype t,", "            const realtype x_[],", "            const realtype p_[],",
           "            realtype dx_[],", "            realtype aux_[],", "            const realtype w_[]) {", ""]

    def section(title, lines):
        out.append(f"    /* {title} */")
        out.extend(lines)
        out.append("")

    section("State variables", [f"    realtype {n} = x_[{i}];" for i, n in enumerate(m.initial_values)])
    section("Parameters", [f"    realtype {n} = p_[{i}];" for i, n in enumerate(m.parameters)])
    section("Noise terms", [f"    realtype {n} = w_[{i}];" for i, n in enumerate(m.noise)])
    body = []
    for line in m.statements:
        if "=" in line:  # an assignment declares its variable
            indent = len(line) - len(line.lstrip())
            line = f"{line[:indent]}realtype {line[indent:]}"
        body.append(f"    {line};")
    section("Core equations", body)
    section("Auxiliary equations", [f"    realtype {n} = {v};" for n, v in m.auxiliaries.items()])
    section("Differential equations", [f"    realtype d{n} = {v};" for n, v in m.derivatives.items()])
    section("Auxiliary outputs", [f"    aux_[{i}] = {n};" for i, n in enumerate(m.auxiliaries)])
    out.append("    /* Differential outputs */")
    out.extend(f"    dx_[{i}] = d{n};" for i, n in enumerate(m.initial_values))
    text = "\n".join(out) + "\n}"

    # `x^n`: small integer powers are written out, other integers use pown, reals pow (xpp_parser.py:143-152)
    if precise_literals:  # reals first: the reference's order turns `v^0.5` into `pown(v, 0).5`
        text = re.sub(r"(\w+)\s*\^\s*([-+]?\d*\.\d+(?:[eE][-+]?\d+)?)", r"pow(\1, \2)", text)
    for n in (2, 3, 4):
        text = re.sub(rf"(\w+)\s*\^\s*{n}", "*".join([r"\1"] * n), text)
    text = re.sub(r"(\w+)\s*\^\s*([0-9]+)", r"pown(\1, \2)", text)
    text = re.sub(r"(\w+)\s*\^\s*([-+]?(\d*\.*\d+))", r"pow(\1, \2)", text)
    if precise_literals:
        text = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?)", r"RCONST(\1)", text)
    else:
        text = re.sub(r"([-+]?(\d+\.\d*))", r"\1f", text)  # the reference forces single-precision literals
    lines = []
    for line in text.split("\n"):
        at = line.find("%")
        lines.append(line if at < 0 else f"{line[:at]}/* {line[at + 1:]} */")
    return "\n".join(lines)


# ---- the reference's function names (clode/xpp_parser.py) ----------------------------------------------------
def read_ode_parameters(xpp_string: str):
    m = parse_xpp(xpp_string)
    return m.parameters, m.auxiliaries, m.initial_values, m.derivatives, m.noise, m.statements


def format_opencl_rhs(parameters, auxiliaries, initial_values, dx, noise, statements, precise_literals: bool = False) -> str:
    return render(XppModel(dict(parameters), dict(auxiliaries), dict(initial_values), dict(dx), list(noise), list(statements)),
                  precise_literals)


def convert_xpp_file(filename: str, precise_literals: bool = False, output: str | None = None) -> str:
    """writes `<name>.cl` next to `<name>.xpp` (or to `output`) and returns its path"""
    with open(filename, "r") as f:
        model = parse_xpp(f.read())
    cl_filename = output or filename[:-4] + ".cl"
    with open(cl_filename, "w") as f:
        f.write(render(model, precise_literals))
    return cl_filename
