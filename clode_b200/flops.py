"""Algorithmic floating-point work per accepted step — the numerator of the FP64 roofline.

Convention: add/sub/mul/div/compare-select(min, max, abs) = 1 flop, FMA = 2,
each transcendental (exp, pow, log, sqrt) = 1, counted on the expressions AS WRITTEN in the
reference source.  Rejected attempts of the adaptive controller are overhead, not work.

    python -m clode_b200.flops        prints the table used by bench.py
"""
from __future__ import annotations

# right-hand sides, counted on clode_b200/models/*.cl (same expressions as the reference fixtures)
F_RHS = {
    "vanderpol": 5,          # mu*(1 - x*x)*y - x : mul sub mul mul sub
    "lorenz63": 9,           # s*(y-x): 2; r*x - y - x*z: 4; x*y - b*z: 3
    "lactotroph": 44,        # 3 gating sigmoids x 5 (sub div exp add div) + currents 19 + derivatives 10 (3 exp incl.)
    "lactotroph_noise": 46,  # + noise current (mul, add)
    "chay_keizer": 35,       # 2 sigmoids x 5 + Hill term 5 + currents 9 + derivatives 11 (2 exp incl.)
    "thompson_a1": 4,        # (w - k*y2)/m: mul sub div; aux: sub
    "sine_drive": 3,         # cos(t*p): mul + cos; aux add
}
N_VAR = {"vanderpol": 2, "lorenz63": 3, "lactotroph": 4, "lactotroph_noise": 4, "chay_keizer": 3,
         "thompson_a1": 2, "sine_drive": 1}


def controller_flops(n: int) -> int:
    """adaptive_explicit_step.clh:13-75 per accepted step: threshold 1, hmin 5, clamp 2,
    per variable (abs abs max max div abs max) 7, compare 1, growth (div pow mul min mul) 5,
    end-of-interval clip (sub min) 2, clamp 2, purified-dt bookkeeping 5  ->  7n + 23"""
    return 7 * n + 23


def flops_per_step(stepper: str, model: str) -> int:
    f, n = F_RHS[model], N_VAR[model]
    if stepper in ("euler", "seuler"):
        return 2 * n + f + 1                      # x += dt*k (2n), t += dt, RHS
    if stepper == "heun":
        return 2 * f + 6 * n + 1                  # predictor fma 2n, corrector (mul add mul add) 4n, 2 RHS
    if stepper == "rk4":
        return 4 * f + 14 * n + 3                 # stages 2n*3, update 8n, half-step / times 3, 4 RHS
    if stepper == "bs23":
        return 3 * f + 21 * n + 6 + controller_flops(n)
    if stepper == "dopri5":
        return 6 * f + 58 * n + 11 + controller_flops(n)
    raise KeyError(stepper)


def trajectory_bytes_per_point(model: str, n_aux: int) -> int:
    """8*(1 + 2*nVar + nAux) bytes per stored point per instance, write-only"""
    return 8 * (1 + 2 * N_VAR[model] + n_aux)


if __name__ == "__main__":
    for m in ("vanderpol", "lorenz63", "lactotroph", "chay_keizer"):
        print(m, {s: flops_per_step(s, m) for s in ("euler", "heun", "rk4", "bs23", "dopri5")})
