/* TEST INFRASTRUCTURE: exposes the portable math header to ctypes for tests/test_pm_math.py */
#include "pm_math.h"
void probe_exp(const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] = pm_exp(x[i]); }
void probe_log(const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] = pm_log(x[i]); }
void probe_sin(const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] = pm_sin(x[i]); }
void probe_cos(const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] = pm_cos(x[i]); }
void probe_pow(const double *x, const double *e, double *y, int n) { for (int i = 0; i < n; ++i) y[i] = pm_pow(x[i], e[i]); }
