// mex.h — a stand-in for MATLAB's header, for machines without MATLAB (this repo's build container and GPU box):
// the subset of the C Matrix / MEX API the gateway uses, with the documented MathWorks signatures, so that
// csrc/matlab/*.cpp compile, link against stub/mex_stub.cpp and can be driven from a test harness.  With a real
// MATLAB the gateway is built by `mex` against MATLAB's own mex.h and this directory is not on the include path.
#pragma once

#include <cstddef>

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

#ifdef __cplusplus
extern "C" {
#endif

bool mxIsChar(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsStruct(const mxArray *a);
bool mxIsCell(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
double mxGetScalar(const mxArray *a);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
char *mxArrayToString(const mxArray *a);
void mxFree(void *p);
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name);
void mxSetField(mxArray *a, mwIndex index, const char *name, mxArray *value);
mxArray *mxGetCell(const mxArray *a, mwIndex index);
void mxSetCell(mxArray *a, mwIndex index, mxArray *value);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateLogicalScalar(bool v);
mxArray *mxCreateString(const char *s);
mxArray *mxCreateCellMatrix(mwSize m, mwSize n);
mxArray *mxCreateStructArray(mwSize ndim, const mwSize *dims, int nfields, const char **names);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
void mxDestroyArray(mxArray *a);

void mexErrMsgTxt(const char *msg);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexWarnMsgTxt(const char *msg);
int mexPrintf(const char *fmt, ...);
void mexLock(void);
void mexUnlock(void);

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

#ifdef __cplusplus
}
#endif
