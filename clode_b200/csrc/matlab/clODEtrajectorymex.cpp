// clODEtrajectorymex — MATLAB entry point for CLODEtrajectory (replaces matlab/clODEtrajectorymex.cpp of the reference); see mex_gateway.hpp
#include "mex_gateway.hpp"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) { clode_mex::dispatch<CLODEtrajectory>(nlhs, plhs, nrhs, prhs); }
