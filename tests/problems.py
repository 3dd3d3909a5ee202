"""Seeded synthetic problem sets shared by the CPU and GPU tests (inputs of the
BASELINE.json configs at reduced size; SURVEY.md §8d gives the full-size definitions)."""
from __future__ import annotations

import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from clode_b200.models import MODELS, MODELS_DIR, rhs_source  # noqa: E402,F401


def ensemble(model: str, n: int):
    """(tspan, x0 flat [nVar*n], pars flat [nPar*n]) in the variable-major API layout."""
    if model == "lorenz63":  # C2: r in [0.5, 60], s = 10, b = 8/3, x0 = (1,1,1)
        pars = np.concatenate([np.linspace(0.5, 60.0, n), np.full(n, 10.0), np.full(n, 8.0 / 3.0)])
        x0, ts = np.ones(3 * n), (0.0, 100.0)
    elif model == "vanderpol":  # C1: mu in [0.1, 10], x0 = (1,1)
        pars, x0, ts = np.linspace(0.1, 10.0, n), np.ones(2 * n), (0.0, 100.0)
    elif model == "lactotroph":  # C3: gcal x gbk grid flattened to a diagonal sweep
        pars = np.concatenate([np.linspace(0.5, 4.0, n), np.full(n, 3.0), np.linspace(0.0, 2.0, n)])
        x0 = np.concatenate([np.full(n, -60.0), np.zeros(n), np.zeros(n), np.full(n, 0.1)])
        ts = (0.0, 2000.0)
    elif model == "lactotroph_noise":  # C4: identical parameters, per-instance noise streams
        pars = np.concatenate([np.full(n, 1.5), np.full(n, 3.0), np.full(n, 1.0), np.full(n, 1.0)])
        x0 = np.concatenate([np.full(n, -60.0), np.zeros(n), np.zeros(n), np.full(n, 0.1)])
        ts = (0.0, 100.0)
    elif model == "chay_keizer":  # C5: gca x kpmca sweep, gkca = 750
        pars = np.concatenate([np.linspace(550.0, 1050.0, n), np.full(n, 750.0), np.linspace(0.095, 0.155, n)])
        x0 = np.concatenate([np.full(n, -50.0), np.full(n, 0.01), np.full(n, 0.12)])
        ts = (0.0, 1000.0)
    elif model == "sine_drive":
        pars, x0, ts = np.linspace(0.5, 2.0, n), np.zeros(n), (0.0, 40.0)
    elif model == "thompson_a1":
        pars = np.concatenate([np.full(n, 0.25), np.full(n, 8.0), np.full(n, 2.0), np.full(n, 10.0)])
        x0, ts = np.zeros(2 * n), (0.0, 5.0)
    else:
        raise KeyError(model)
    return ts, x0, pars
