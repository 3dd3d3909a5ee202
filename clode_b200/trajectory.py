"""`TrajectorySimulator` / `TrajectoryOutput` — mirror of the reference's clode/trajectory.py:18-266."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
from numpy.lib import recfunctions as rfn

from .cpp.clode_cpp_wrapper import SolverParams, TrajectorySimulatorBase
from .runtime import _clode_root_dir
from .solver import Simulator, Stepper


class TrajectoryOutput:
    def __init__(self, t, x, dx, aux, variable_names: List[str], aux_names: List[str]) -> None:
        self.t = t
        x_dtype = np.dtype({"names": variable_names, "formats": [np.float64] * len(variable_names)})
        self.x = rfn.unstructured_to_structured(x, dtype=x_dtype)
        self.dx = rfn.unstructured_to_structured(dx, dtype=x_dtype)
        if aux_names:
            aux_dtype = np.dtype({"names": aux_names, "formats": [np.float64] * len(aux_names)})
            self.aux = rfn.unstructured_to_structured(aux, dtype=aux_dtype)
        self._variable_names, self._aux_names = variable_names, aux_names

    def __repr__(self) -> str:
        return (f"TrajectoryOutput( length: {len(self.t)}, variable names: {self._variable_names}, "
                f"aux variable names: {self._aux_names} )")

    def to_ndarray(self, slot: str, **kwargs):
        return rfn.structured_to_unstructured({"x": self.x, "dx": self.dx, "aux": getattr(self, "aux", None)}[slot], **kwargs)


class TrajectorySimulator(Simulator):
    _integrator: TrajectorySimulatorBase

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
        reltol: float = 1e-4,
        max_steps: int = 1000000,
        max_store: int = 1000000,
        nout: int = 1,
        solver_parameters: Optional[SolverParams] = None,
        t_span: Tuple[float, float] = (0.0, 1000.0),
        single_precision: bool = True,
        device_type=None,
        vendor=None,
        platform_id: Optional[int] = None,
        device_id: Optional[int] = None,
        device_ids: Optional[List[int]] = None,
        stream_chunk_rows: Optional[int] = None,
    ) -> None:
        # not in the reference (TODO at clode/trajectory.py:166): integrate in launches of `stream_chunk_rows` stored
        # points, copying each chunk to the host while the next one integrates; results are identical
        self._stream_chunk_rows = int(stream_chunk_rows or 0)
        super().__init__(variables=variables, parameters=parameters, aux=aux, num_noise=num_noise, src_file=src_file,
                         rhs_equation=rhs_equation, supplementary_equations=supplementary_equations, stepper=stepper,
                         dt=dt, dtmax=dtmax, abstol=abstol, reltol=reltol, max_steps=max_steps, max_store=max_store,
                         nout=nout, solver_parameters=solver_parameters, t_span=t_span, single_precision=single_precision,
                         device_type=device_type, vendor=vendor, platform_id=platform_id, device_id=device_id,
                         device_ids=device_ids)

    def _create_integrator(self) -> None:
        self._integrator = TrajectorySimulatorBase(self._pi, self._stepper.value, self._single_precision, self._runtime,
                                                   _clode_root_dir)
        self._integrator.set_stream_chunk(self._stream_chunk_rows)

    def set_stream_chunk(self, rows: Optional[int]) -> None:
        """0 / None: one launch holding all max_store points on the device (the reference's scheme); > 0: streamed"""
        self._stream_chunk_rows = int(rows or 0)
        self._integrator.set_stream_chunk(self._stream_chunk_rows)

    def trajectory(self, t_span=None, update_x0: bool = True, fetch_results: bool = True):
        if t_span is not None:
            self.set_tspan(t_span=t_span)
        self._integrator.trajectory()
        self._device_final_state = self._device_dt = self._device_tf = None
        if update_x0:
            self._integrator.shift_x0()
            self._device_initial_state = None
        if fetch_results:
            return self.get_trajectory()

    def get_trajectory(self):
        """one TrajectoryOutput per instance (a single object for an ensemble of one)"""
        n, nv, na, ms = self._ensemble_size, self.num_variables, len(self.aux_names), self._sp.max_store
        n_stored = np.asarray(self._integrator.get_n_stored())
        t = np.asarray(self._integrator.get_t_array())[: n * ms].reshape((n, ms), order="F")
        x = np.asarray(self._integrator.get_x_array())[: n * nv * ms].reshape((n, nv, ms), order="F")
        dx = np.asarray(self._integrator.get_dx_array())[: n * nv * ms].reshape((n, nv, ms), order="F")
        aux = np.asarray(self._integrator.get_aux_array())[: n * na * ms].reshape((n, na, ms), order="F")
        results = []
        for i in range(n):
            k = int(n_stored[i]) + 1
            results.append(TrajectoryOutput(t=t[i, :k].transpose(), x=x[i, :, :k].transpose(), dx=dx[i, :, :k].transpose(),
                                            aux=aux[i, :, :k].transpose(), variable_names=self.variable_names,
                                            aux_names=self.aux_names))
        return results[0] if n == 1 else results
