// clODEmex — MATLAB entry point for CLODE (replaces matlab/clODEmex.cpp of the reference); see mex_gateway.hpp
#include "mex_gateway.hpp"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) { clode_mex::dispatch<CLODE>(nlhs, plhs, nrhs, prhs); }
