"""`FeatureSimulator` / `ObserverOutput` — mirror of the reference's clode/features.py:24-534."""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, Optional, Tuple

import numpy as np
from numpy.lib import recfunctions as rfn

from .cpp.clode_cpp_wrapper import FeatureSimulatorBase, ObserverParams, SolverParams
from .runtime import _clode_root_dir
from .solver import Simulator, Stepper


class Observer(Enum):
    basic = "basic"
    basic_all_variables = "basicall"
    local_max = "localmax"
    neighbourhood_1 = "nhood1"
    neighbourhood_2 = "nhood2"
    threshold_2 = "thresh2"


class ObserverOutput:
    """Feature matrix with name-based access (structured array `F`, one field per feature)."""

    def __init__(self, observer_params, feature_array, num_features, variables, observer_type, feature_names,
                 ensemble_shape) -> None:
        self._op = observer_params
        self._num_features = num_features
        self._vars = variables
        self._observer_type = observer_type
        self._feature_names = feature_names
        self._ensemble_shape = ensemble_shape
        dtype = np.dtype({"names": feature_names, "formats": [np.float64] * len(feature_names)})
        a = np.asarray(feature_array)
        if a.ndim == 2 and a.dtype == np.float64 and a.flags.c_contiguous and a.shape[1] == len(feature_names):
            # (ensemble, features) matrix as the native layer delivers it (transposed on the GPU): the record array is
            # a VIEW of it — the reference's unstructured_to_structured copy (clode/features.py:51) costs ~100 ms at 2^20
            self.F = a.view(dtype).reshape(a.shape[0])
        else:
            self.F = rfn.unstructured_to_structured(a, dtype=dtype)

    def __repr__(self) -> str:
        return (f"ObserverOutput( ensemble size: {len(self.F[self._feature_names[0]])}, number of features: "
                f"{len(self._feature_names)}, feature_names: {self._feature_names})")

    def to_ndarray(self, **kwargs):
        return rfn.structured_to_unstructured(self.F, **kwargs)

    def get_feature_names(self) -> List[str]:
        return self._feature_names

    def _get_var(self, var: str):
        try:
            out = self.F[var].squeeze().reshape(self._ensemble_shape)
        except ValueError:
            raise NotImplementedError(f"{self._observer_type} does not track {var}!")
        return out[0] if out.size == 1 else out

    def get_var_max(self, var: str):
        return self._get_var(f"max {var}")

    def get_var_min(self, var: str):
        return self._get_var(f"min {var}")

    def get_var_mean(self, var: str):
        return self._get_var(f"mean {var}")

    def get_var_max_slope(self, var: str):
        return self.get_var_max(f"d{var}/dt")

    def get_var_min_slope(self, var: str):
        return self.get_var_min(f"d{var}/dt")

    def get_var_count(self, var: str):
        return self._get_var(f"{var} count")

    def get_event_data(self, name: str, type: Optional[str] = "time"):
        """stored event times (or event-variable values) of the events called `name`, up to the first all-zero slot"""
        if not any(name in f and "event" in f and "count" not in f for f in self._feature_names):
            raise NotImplementedError(f"{self._observer_type} does not track {name} event {type}s!")
        data = []
        for k in range(self._op.max_event_timestamps):
            col = self._get_var(f"{name} event {type} {k}")
            if np.all(col == 0):
                break
            data.append(col)
        return np.stack(data, axis=-1).squeeze()

    def get_timestamps(self, var: str = "event"):
        if f"{var} event time 0" not in self._feature_names:
            raise NotImplementedError(f"{self._observer_type} does not track {var} event times!")
        data = []
        for k in range(self._op.max_event_timestamps):
            col = self._get_var(f"{var} event time {k}")
            if np.all(col == 0):
                break
            data.append(col[np.newaxis] if np.ndim(col) == 0 else col)
        return np.stack(data, axis=1).squeeze() if data else []


class FeatureSimulator(Simulator):
    _integrator: FeatureSimulatorBase

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
        max_steps: int = 10000000,
        max_store: int = 10000000,
        nout: int = 1,
        solver_parameters: Optional[SolverParams] = None,
        t_span: Tuple[float, float] = (0.0, 1000.0),
        single_precision: bool = True,
        device_type=None,
        vendor=None,
        platform_id: Optional[int] = None,
        device_id: Optional[int] = None,
        device_ids: Optional[List[int]] = None,
        observer: Observer = Observer.basic_all_variables,
        event_var: str = "",
        feature_var: str = "",
        observer_max_event_count: int = 100,
        observer_max_event_timestamps: int = 0,
        observer_min_x_amp: float = 0.0,
        observer_min_imi: float = 0.0,
        observer_neighbourhood_radius: float = 0.05,
        observer_x_up_thresh: float = 0.3,
        observer_x_down_thresh: float = 0.2,
        observer_dx_up_thresh: float = 0,
        observer_dx_down_thresh: float = 0,
        observer_eps_dx: float = 0.0,
        observer_parameters: Optional[ObserverParams] = None,
    ) -> None:
        self._observer_type = observer
        names = list(variables.keys())
        e_ix = names.index(event_var) if event_var != "" else 0
        f_ix = names.index(feature_var) if feature_var != "" else 0
        self._op = observer_parameters if observer_parameters is not None else ObserverParams(
            e_ix, f_ix, observer_max_event_count, observer_max_event_timestamps, observer_min_x_amp, observer_min_imi,
            observer_neighbourhood_radius, observer_x_up_thresh, observer_x_down_thresh, observer_dx_up_thresh,
            observer_dx_down_thresh, observer_eps_dx)
        self._device_features = None
        self._num_features = None
        super().__init__(variables=variables, parameters=parameters, aux=aux, num_noise=num_noise, src_file=src_file,
                         rhs_equation=rhs_equation, supplementary_equations=supplementary_equations, stepper=stepper,
                         dt=dt, dtmax=dtmax, abstol=abstol, reltol=reltol, max_steps=max_steps, max_store=max_store,
                         nout=nout, solver_parameters=solver_parameters, t_span=t_span, single_precision=single_precision,
                         device_type=device_type, vendor=vendor, platform_id=platform_id, device_id=device_id,
                         device_ids=device_ids)

    def _create_integrator(self) -> None:
        self._integrator = FeatureSimulatorBase(self._pi, self._stepper.value, self._observer_type.value, self._op,
                                                self._single_precision, self._runtime, _clode_root_dir)

    def set_observer(self, observer_type: Observer):
        if observer_type != self._observer_type:
            self._integrator.set_observer(observer_type.value)
            self._observer_type = observer_type
            self._cl_program_is_valid = False

    def set_observer_parameters(self, op: Optional[ObserverParams] = None, event_var=None, feature_var=None,
                                max_event_count=None, max_event_timestamps=None, min_amp=None, min_imi=None,
                                nhood_radius=None, x_up_threshold=None, x_down_threshold=None, dx_up_threshold=None,
                                dx_down_threshold=None, eps_dx=None) -> None:
        if op is not None:
            self._op = op
        else:
            if event_var is not None:
                self._op.e_var_ix = self.variable_names.index(event_var)
            if feature_var is not None:
                self._op.f_var_ix = self.variable_names.index(feature_var)
            for name, value in (("max_event_count", max_event_count), ("max_event_timestamps", max_event_timestamps),
                                ("min_amp", min_amp), ("min_imi", min_imi), ("nhood_radius", nhood_radius),
                                ("x_up_threshold", x_up_threshold), ("x_down_threshold", x_down_threshold),
                                ("dx_up_threshold", dx_up_threshold), ("dx_down_threshold", dx_down_threshold),
                                ("eps_dx", eps_dx)):
                if value is not None:
                    setattr(self._op, name, value)
        self._integrator.set_observer_params(self._op)

    def get_observer_parameters(self):
        return self._integrator.get_observer_params()

    def get_feature_names(self) -> List[str]:
        return self._integrator.get_feature_names()

    def is_observer_initialized(self):
        return self._integrator.is_observer_initialized()

    def initialize_observer(self):
        self._integrator.initialize_observer()

    def features(self, t_span=None, initialize_observer: Optional[bool] = None, update_x0: bool = True,
                 fetch_results: bool = True) -> Optional[ObserverOutput]:
        if t_span is not None:
            self.set_tspan(t_span=t_span)
        if initialize_observer is not None:
            self._integrator.features(initialize_observer)
        else:
            self._integrator.features()
        self._device_features = None
        self._device_final_state = self._device_dt = self._device_tf = None
        if update_x0:
            self._integrator.shift_x0()
            self._device_initial_state = None
        if fetch_results:
            return self.get_observer_results()

    def get_observer_results(self) -> ObserverOutput:
        if self._device_features is None:
            self._num_features = self._integrator.get_n_features()
            if hasattr(self._integrator, "get_f_matrix"):
                self._device_features = self._integrator.get_f_matrix()
            else:
                self._device_features = self._matrix(self._integrator.get_f_array(), self._num_features)
        return ObserverOutput(self._op, self._device_features, self._num_features, self.variable_names,
                              self._observer_type, self._integrator.get_feature_names(), self._ensemble_shape)
