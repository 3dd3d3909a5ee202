// clODEfeaturesmex — MATLAB entry point for CLODEfeatures (replaces matlab/clODEfeaturesmex.cpp of the reference); see mex_gateway.hpp
#include "mex_gateway.hpp"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) { clode_mex::dispatch<CLODEfeatures>(nlhs, plhs, nrhs, prhs); }
