"""`Simulator` — ensemble integration that keeps only the final state.

Mirror of the reference's clode/solver.py:49-748 (`Stepper`, `Simulator`): same constructor
arguments, method names, array conventions (ensemble arrays are (ensemble_size, n) matrices, handed
to the C++ layer flattened in Fortran order, clode/solver.py:478-481) and ensemble-building rules.
The right-hand side comes from an OpenCL-C source file (`src_file`), an XPP file (`src_file="m.xpp"`,
converted by `xpp_parser`) or a typed Python function (`rhs_equation=`, with `supplementary_equations=` as
helpers, converted by `function_converter`) — clode/solver.py:220-252.
"""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, Mapping, Optional, Tuple, Union

import numpy as np

from .cpp.clode_cpp_wrapper import ProblemInfo, SimulatorBase, SolverParams
from .runtime import CLDeviceType, CLVendor, OpenCLResource, _clode_root_dir, initialize_runtime


class Stepper(Enum):
    euler = "euler"
    heun = "heun"
    rk4 = "rk4"
    bs23 = "bs23"
    dormand_prince = "dopri5"
    stochastic_euler = "seuler"


ArrayOrMap = Union[np.ndarray, Mapping[str, Union[float, List[float], np.ndarray]]]


def _as_f(a: np.ndarray) -> np.ndarray:
    """(ensemble, n) matrix -> flat variable-major vector"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).flatten(order="F"))


def _as_matrix(a: np.ndarray):
    """the (ensemble, n) float64 matrix itself when the native layer can read it in place (positive strides that are
    multiples of 8 bytes): it is then transposed into the device layout inside the runtime's staging copy instead of
    by a host-side `flatten(order="F")` (clode/solver.py:384 does the flatten; 25 ms for 2^20 x 3 values)"""
    a = np.asarray(a)
    if a.dtype != np.float64 or a.ndim != 2 or any(s <= 0 or s % 8 for s in a.strides) or a.shape[0] == 0:
        return None
    return a


class Simulator:
    _integrator: SimulatorBase
    _runtime: OpenCLResource
    _generated_files: set = set()  # temporary getRHS files written for Python right-hand sides

    def __init__(
        self,
        variables: Dict[str, float],
        parameters: Dict[str, float],
        aux: Optional[List[str]] = None,
        num_noise: int = 0,
        src_file: Optional[str] = None,
        rhs_equation=None,
        supplementary_equations=None,
        stepper: Stepper = Stepper.rk4,
        dt: float = 0.1,
        dtmax: float = 1.0,
        abstol: float = 1e-6,
        reltol: float = 1e-3,
        max_steps: int = 1000000,
        max_store: int = 1000000,
        nout: int = 1,
        solver_parameters: Optional[SolverParams] = None,
        t_span: Tuple[float, float] = (0.0, 1000.0),
        single_precision: bool = True,
        device_type: Optional[CLDeviceType] = None,
        vendor: Optional[CLVendor] = None,
        platform_id: Optional[int] = None,
        device_id: Optional[int] = None,
        device_ids: Optional[List[int]] = None,
    ) -> None:
        if src_file is not None and rhs_equation is not None:
            raise ValueError("Cannot specify both src_file and rhs_equation")
        if src_file is None and rhs_equation is None:
            raise ValueError("Must specify either src_file or rhs_equation")
        src_file = self._handle_clode_rhs_cl_file(src_file, rhs_equation, supplementary_equations)
        self._pi = ProblemInfo(src_file, list(variables.keys()), list(parameters.keys()), list(aux or []), num_noise)
        self._stepper = stepper
        self._single_precision = single_precision
        self._runtime = initialize_runtime(device_type, vendor, platform_id, device_id, device_ids)
        self._device_parameters = None
        self._device_initial_state = None
        self._device_final_state = self._device_dt = self._device_tf = None

        self._create_integrator()
        self._build_cl_program()
        if src_file in Simulator._generated_files:  # the C++ layer holds the source text now (CLODE::setProblemInfo)
            import os
            Simulator._generated_files.discard(src_file)
            try:
                os.remove(src_file)
            except OSError:
                pass

        self._sp = solver_parameters if solver_parameters is not None else SolverParams(
            dt, dtmax, abstol, reltol, max_steps, max_store, nout)
        self.set_solver_parameters()
        self.set_tspan(t_span=t_span)

        self._variable_defaults = dict(variables)
        self._parameter_defaults = dict(parameters)
        self._ensemble_size = 1
        self._ensemble_shape: Tuple = (1,)
        self._set_problem_data(np.array(list(variables.values()), dtype=np.float64, ndmin=2),
                               np.array(list(parameters.values()), dtype=np.float64, ndmin=2))

    @staticmethod
    def _handle_clode_rhs_cl_file(src_file, rhs_equation, supplementary_equations) -> str:
        """clode/solver.py:220-252: `.xpp` files and Python functions become an OpenCL-C file.  The reference writes
        `clode_rhs.cl` into the working directory; here the generated text goes to a private temporary file, so
        concurrent simulators (and read-only working directories) do not collide."""
        if src_file is not None:
            if src_file.endswith(".xpp"):
                from .xpp_parser import convert_xpp_file
                return convert_xpp_file(src_file)
            return src_file
        import os
        import tempfile

        from .function_converter import OpenCLConverter
        converter = OpenCLConverter()
        for eq in supplementary_equations or []:
            converter.convert_to_opencl(eq)
        text = converter.convert_to_opencl(rhs_equation, mutable_args=[3, 4], function_name="getRHS")
        fd, path = tempfile.mkstemp(prefix="clode_rhs_", suffix=".cl")
        with os.fdopen(fd, "w") as f:
            f.write(text)
        # ProblemInfo / setProblemInfo read the file when the integrator is created; removed at interpreter exit at the
        # latest (and by the simulator's finalizer, see __init__)
        import atexit
        atexit.register(lambda p=path: os.path.exists(p) and os.remove(p))
        Simulator._generated_files.add(path)
        return path

    # ---- properties (clode/solver.py:83-119) ---------------------------------------------------
    @property
    def variable_names(self) -> List[str]:
        return self._pi.vars

    @property
    def num_variables(self) -> int:
        return self._pi.num_var

    @property
    def parameter_names(self) -> List[str]:
        return self._pi.pars

    @property
    def num_parameters(self) -> int:
        return self._pi.num_par

    @property
    def aux_names(self) -> List[str]:
        return self._pi.aux

    @property
    def num_aux(self) -> int:
        return self._pi.num_aux

    @property
    def num_noise(self) -> int:
        return self._pi.num_noise

    # ---- construction hooks --------------------------------------------------------------------
    def _create_integrator(self) -> None:
        self._integrator = SimulatorBase(self._pi, self._stepper.value, self._single_precision, self._runtime,
                                         _clode_root_dir)

    def _build_cl_program(self):
        self._integrator.build_cl()
        self._cl_program_is_valid = True

    # ---- ensembles (clode/solver.py:254-503) ---------------------------------------------------
    def set_repeat_ensemble(self, num_repeats: int) -> None:
        x0, p = self._make_problem_data(new_size=num_repeats, new_shape=(num_repeats, 1))
        self._set_problem_data(x0, p)

    def _size_and_shape(self, spec, names, what):
        """ensemble size / shape implied by an array or a name->values mapping (scalars broadcast)"""
        if isinstance(spec, np.ndarray):
            if spec.ndim != 2 or spec.shape[1] != len(names):
                raise ValueError(f"{what} must be a matrix with {len(names)} columns")
            return spec, spec.shape[0], (spec.shape[0], 1)
        if isinstance(spec, Mapping):
            unknown = set(spec.keys()) - set(names)
            if unknown:
                raise ValueError(f"Unknown {what} name(s): {unknown}")
            spec = {k: np.array(v, dtype=np.float64) for k, v in spec.items()}
            shapes = {k: v.shape for k, v in spec.items() if v.size > 1}
            if len(set(shapes.values())) > 1:
                raise ValueError(f"Shape of arrays for {what} don't match: {shapes}")
            if shapes:
                shape = next(iter(shapes.values()))
                return spec, int(np.prod(shape)), shape
            return spec, 1, (1,)
        if spec is not None:
            raise ValueError(f"Expected np.ndarray or Mapping for {what}, but got {type(spec)}")
        return None, 1, (1,)

    def set_ensemble(self, variables: Optional[ArrayOrMap] = None, parameters: Optional[ArrayOrMap] = None) -> None:
        if variables is None and parameters is None:
            raise ValueError("initial_state and parameters cannot both be None")
        variables, var_size, var_shape = self._size_and_shape(variables, self.variable_names, "variables")
        parameters, par_size, par_shape = self._size_and_shape(parameters, self.parameter_names, "parameters")
        if var_size > 1 and par_size > 1 and var_size != par_size:
            raise ValueError("Arrays specified for parameters and initial states must have the same size")
        new_size, new_shape = (var_size, var_shape) if var_size > 1 else (par_size, par_shape)
        x0, p = self._make_problem_data(variables, parameters, new_size, new_shape)
        self._set_problem_data(x0, p)

    def _make_problem_data(self, variables=None, parameters=None, new_size=None, new_shape=None):
        if len(new_shape) == 1:
            new_shape = (new_size, 1)
        # keep the current state when the ensemble keeps its size or grows from a single instance; a side that is given
        # as a complete matrix needs no starting point at all (and no read-back of the device's current state)
        keep = self._ensemble_size in (new_size, 1)
        x0 = p = None
        if not isinstance(variables, np.ndarray):
            x0 = (np.array(self.get_initial_state(), dtype=np.float64) if keep else
                  np.array(list(self._variable_defaults.values()), dtype=np.float64, ndmin=2))
            if x0.shape[0] == 1:
                x0 = np.tile(x0, (new_size, 1))
        if not isinstance(parameters, np.ndarray):
            p = (np.array(self._device_parameters, dtype=np.float64) if keep else
                 np.array(list(self._parameter_defaults.values()), dtype=np.float64, ndmin=2))
            if p.shape[0] == 1:
                p = np.tile(p, (new_size, 1))
        for spec, target, names in ((variables, "x0", self.variable_names), (parameters, "p", self.parameter_names)):
            if isinstance(spec, np.ndarray):
                if target == "x0":
                    x0 = spec
                else:
                    p = spec
            elif isinstance(spec, Mapping):
                dest = x0 if target == "x0" else p
                for key, value in spec.items():
                    dest[:, names.index(key)] = np.repeat(value, new_size) if value.size == 1 else value.flatten()
        self._ensemble_size, self._ensemble_shape = new_size, new_shape
        return x0, p

    def _set_problem_data(self, initial_state: np.ndarray, parameters: np.ndarray) -> None:
        self._device_initial_state, self._device_parameters = initial_state, parameters
        x0m, pm = _as_matrix(initial_state), _as_matrix(parameters)
        if x0m is not None and pm is not None and hasattr(self._integrator, "set_problem_data_matrix"):
            self._integrator.set_problem_data_matrix(x0m, pm)
        else:
            self._integrator.set_problem_data(_as_f(initial_state), _as_f(parameters))

    def _set_parameters(self, parameters: np.ndarray) -> None:
        self._device_parameters = parameters
        pm = _as_matrix(parameters)
        if pm is not None and hasattr(self._integrator, "set_pars_matrix"):
            self._integrator.set_pars_matrix(pm)
        else:
            self._integrator.set_pars(_as_f(parameters))

    def _set_initial_state(self, initial_state: np.ndarray) -> None:
        self._device_initial_state = initial_state
        x0m = _as_matrix(initial_state)
        if x0m is not None and hasattr(self._integrator, "set_x0_matrix"):
            self._integrator.set_x0_matrix(x0m)
        else:
            self._integrator.set_x0(_as_f(initial_state))

    # ---- time span / solver parameters -----------------------------------------------------------
    def set_tspan(self, t_span: Tuple[float, float]) -> None:
        self._t_span = tuple(t_span)
        self._integrator.set_tspan(list(t_span))

    def get_tspan(self) -> Tuple[float, float]:
        self._t_span = tuple(self._integrator.get_tspan())
        return self._t_span

    def shift_tspan(self) -> None:
        self._integrator.shift_tspan()
        self._t_span = tuple(self._integrator.get_tspan())

    def set_solver_parameters(self, solver_parameters: Optional[SolverParams] = None, dt=None, dtmax=None, abstol=None,
                              reltol=None, max_steps=None, max_store=None, nout=None) -> None:
        if solver_parameters is not None:
            self._sp = solver_parameters
        else:
            for name, value in (("dt", dt), ("dtmax", dtmax), ("abstol", abstol), ("reltol", reltol),
                                ("max_steps", max_steps), ("max_store", max_store), ("nout", nout)):
                if value is not None:
                    setattr(self._sp, name, value)
        self._integrator.set_solver_params(self._sp)

    def get_solver_parameters(self):
        return self._integrator.get_solver_params()

    def seed_rng(self, seed: Optional[int] = None) -> None:
        if seed is None:
            self._integrator.seed_rng()
        else:
            self._integrator.seed_rng(seed)

    # ---- simulation ----------------------------------------------------------------------------
    def transient(self, t_span=None, update_x0: bool = True, fetch_results: bool = False):
        if t_span is not None:
            self.set_tspan(t_span=t_span)
        self._integrator.transient()
        self._device_final_state = self._device_dt = self._device_tf = None
        if update_x0:
            self._integrator.shift_x0()
            self._device_initial_state = None
        if fetch_results:
            return self.get_final_state()

    def _matrix(self, flat, ncol) -> np.ndarray:
        return np.asarray(flat, dtype=np.float64).reshape((self._ensemble_size, ncol), order="F")

    def get_initial_state(self) -> np.ndarray:
        if self._device_initial_state is None:
            if hasattr(self._integrator, "get_x0_matrix"):
                self._device_initial_state = self._integrator.get_x0_matrix()
            else:
                self._device_initial_state = self._matrix(self._integrator.get_x0_array(), self.num_variables)
        return self._device_initial_state

    def get_final_state(self) -> np.ndarray:
        if self._device_final_state is None:
            if hasattr(self._integrator, "get_xf_matrix"):
                self._device_final_state = self._integrator.get_xf_matrix()
            else:
                self._device_final_state = self._matrix(self._integrator.get_xf_array(), self.num_variables)
        return self._device_final_state

    def get_dt(self) -> np.ndarray:
        if self._device_dt is None:
            self._device_dt = np.asarray(self._integrator.get_dt_array()).reshape(self._ensemble_shape, order="F")
        return self._device_dt

    def get_final_time(self) -> np.ndarray:
        if self._device_tf is None:
            self._device_tf = np.asarray(self._integrator.get_tf_array()).reshape(self._ensemble_shape, order="F")
        return self._device_tf

    # ---- device / program queries --------------------------------------------------------------
    def get_max_memory_alloc_size(self, deviceID: int = 0) -> int:
        return self._runtime.get_max_memory_alloc_size(deviceID)

    def get_double_support(self, deviceID: int = 0) -> bool:
        return self._runtime.get_double_support(deviceID)

    def get_device_cl_version(self, deviceID: int = 0) -> str:
        return self._runtime.get_device_cl_version(deviceID)

    def get_available_steppers(self) -> List[str]:
        return self._integrator.get_available_steppers()

    def get_program_string(self) -> str:
        return self._integrator.get_program_string()

    def print_status(self) -> None:
        self._integrator.print_status()

    def print_devices(self) -> None:
        self._runtime.print_devices()
