"""Right-hand sides of the benchmark / test systems (OpenCL-C `getRHS` files next to this module)
and their dimensions."""
import os

MODELS_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> (nVar, nPar, nAux, nWiener); source in <name>.cl
MODELS = {
    "lorenz63": (3, 3, 1, 0),
    "vanderpol": (2, 1, 0, 0),
    "thompson_a1": (2, 4, 1, 0),
    "lactotroph": (4, 3, 1, 0),
    "lactotroph_noise": (4, 4, 1, 1),
    "chay_keizer": (3, 3, 0, 0),
    "sine_drive": (1, 1, 3, 0),
}


def rhs_path(model: str) -> str:
    return os.path.join(MODELS_DIR, model + ".cl")


def rhs_source(model: str) -> str:
    with open(rhs_path(model)) as f:
        return f.read()
