// mex_stub.cpp — a small in-process implementation of the Matrix / MEX API subset declared in stub/mex.h, so the
// gateway can be exercised without MATLAB (tests/test_mex_gateway.py).  Errors become C++ exceptions (MexError)
// that the harness catches, which is also what MATLAB does with mexErrMsgIdAndTxt under the hood.
#include "mex.h"

#include "mex_stub.hpp"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct mxArray_tag {
    enum Kind { Double, Char, Cell, Struct, Logical } kind = Double;
    size_t m = 0, n = 0;
    std::vector<double> real;
    std::string text;
    std::vector<mxArray *> cells;                           // Cell: m*n entries
    std::vector<std::string> field_names;                   // Struct
    std::vector<std::map<std::string, mxArray *>> elements; // Struct: m*n maps
};

int mex_stub_lock_count = 0;
std::string mex_stub_last_warning;

extern "C" {

bool mxIsChar(const mxArray *a) { return a && a->kind == mxArray::Char; }
bool mxIsDouble(const mxArray *a) { return a && a->kind == mxArray::Double; }
bool mxIsStruct(const mxArray *a) { return a && a->kind == mxArray::Struct; }
bool mxIsCell(const mxArray *a) { return a && a->kind == mxArray::Cell; }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->kind == mxArray::Char ? a->text.size() : a->m * a->n; }
double mxGetScalar(const mxArray *a)
{
    if (!a || a->real.empty()) throw MexError("stub:scalar", "mxGetScalar on an empty or non-numeric array");
    return a->real[0];
}
double *mxGetPr(const mxArray *a) { return const_cast<double *>(a->real.data()); }
void *mxGetData(const mxArray *a) { return const_cast<double *>(a->real.data()); }
char *mxArrayToString(const mxArray *a)
{
    if (!mxIsChar(a)) return nullptr;
    char *s = (char *)std::malloc(a->text.size() + 1);
    std::memcpy(s, a->text.c_str(), a->text.size() + 1);
    return s;
}
void mxFree(void *p) { std::free(p); }
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name)
{
    if (!mxIsStruct(a) || index >= a->elements.size()) return nullptr;
    auto it = a->elements[index].find(name);
    return it == a->elements[index].end() ? nullptr : it->second;
}
void mxSetField(mxArray *a, mwIndex index, const char *name, mxArray *value)
{
    if (!mxIsStruct(a) || index >= a->elements.size()) throw MexError("stub:field", "mxSetField: bad struct or index");
    a->elements[index][name] = value;
}
mxArray *mxGetCell(const mxArray *a, mwIndex index) { return (mxIsCell(a) && index < a->cells.size()) ? a->cells[index] : nullptr; }
void mxSetCell(mxArray *a, mwIndex index, mxArray *value)
{
    if (!mxIsCell(a) || index >= a->cells.size()) throw MexError("stub:cell", "mxSetCell: bad cell array or index");
    a->cells[index] = value;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity)
{
    mxArray *a = new mxArray;
    a->m = m; a->n = n;
    a->real.assign(m * n, 0.0);
    return a;
}
mxArray *mxCreateDoubleScalar(double v)
{
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->real[0] = v;
    return a;
}
mxArray *mxCreateLogicalScalar(bool v)
{
    mxArray *a = mxCreateDoubleScalar(v ? 1.0 : 0.0);
    a->kind = mxArray::Logical;
    return a;
}
mxArray *mxCreateString(const char *s)
{
    mxArray *a = new mxArray;
    a->kind = mxArray::Char;
    a->text = s ? s : "";
    a->m = 1; a->n = a->text.size();
    return a;
}
mxArray *mxCreateCellMatrix(mwSize m, mwSize n)
{
    mxArray *a = new mxArray;
    a->kind = mxArray::Cell;
    a->m = m; a->n = n;
    a->cells.assign(m * n, nullptr);
    return a;
}
mxArray *mxCreateStructArray(mwSize ndim, const mwSize *dims, int nfields, const char **names)
{
    mxArray *a = new mxArray;
    a->kind = mxArray::Struct;
    a->m = ndim > 0 ? dims[0] : 1;
    a->n = ndim > 1 ? dims[1] : 1;
    for (int i = 0; i < nfields; ++i) a->field_names.push_back(names[i]);
    a->elements.resize(a->m * a->n);
    return a;
}
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names)
{
    const mwSize dims[2] = {m, n};
    return mxCreateStructArray(2, dims, nfields, names);
}
void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    for (mxArray *c : a->cells) mxDestroyArray(c);
    for (auto &e : a->elements)
        for (auto &kv : e) mxDestroyArray(kv.second);
    delete a;
}

void mexErrMsgTxt(const char *msg) { throw MexError("", msg ? msg : ""); }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw MexError(id ? id : "", buf);
}
void mexWarnMsgTxt(const char *msg) { mex_stub_last_warning = msg ? msg : ""; }
int mexPrintf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    const int n = std::vfprintf(stdout, fmt, ap);
    va_end(ap);
    return n;
}
void mexLock(void) { ++mex_stub_lock_count; }
void mexUnlock(void) { --mex_stub_lock_count; }
}
