void getRHS(const realtype t,
            const realtype x[],
            const realtype p[],
            realtype dx[],
            realtype aux[],
            const realtype w[]) {
    int p0 = (int)(p[0]);
    int p1 = (int)(p[1]);
    realtype x0 = (x[0] * t);
    realtype x1 = (x[1] * t);
    realtype x2 = (x[2] * t);
    aux[0] = acos(x0);
    aux[1] = acosh((x0 + 1));
    aux[2] = acospi(x0);
    aux[3] = asin(x0);
    aux[4] = asinh((x0 + 1));
    aux[5] = asinpi(x0);
    aux[6] = atan(x0);
    aux[7] = atan2(x0, x1);
    aux[8] = atan2pi(x0, x1);
    aux[9] = atanh((x0 / RCONST(100.0)));
    aux[10] = atanpi(x0);
    aux[11] = cbrt(x0);
    aux[12] = ceil(x0);
    aux[13] = copysign(x0, x1);
    aux[14] = cos(x0);
    aux[15] = cosh(x0);
    aux[16] = cospi(x0);
    aux[17] = erf(x0);
    aux[18] = erfc(x0);
    aux[19] = exp(x0);
    aux[20] = exp2(x0);
    aux[21] = exp10(x0);
    aux[22] = expm1(x0);
    aux[23] = fabs(x0);
    aux[24] = fdim(x0, x1);
    aux[25] = floor(x1);
    aux[26] = RCONST(0.0);
    aux[27] = fmod(x0, (x1 + 1));
    aux[28] = heaviside((t - RCONST(0.5)));
    aux[29] = tgamma((x0 + 1));
    aux[30] = hypot(x0, x1);
    aux[31] = (realtype)(ilogb((x0 + 1)));
    aux[32] = ldexp(x0, (int)(x1));
    aux[33] = lgamma((x0 + 1));
    aux[34] = log((x0 + 1));
    aux[35] = log1p((x0 + 1));
    aux[36] = log2((x0 + 1));
    aux[37] = log10((x0 + 1));
    aux[38] = RCONST(0.0);
    aux[39] = RCONST(0.0);
    aux[40] = RCONST(0.0);
    aux[41] = nextafter(x0, x1);
    aux[42] = pow(x0, x1);
    aux[43] = pown(x0, p0);
    aux[44] = powr((x0 + RCONST(0.5)), (x1 + RCONST(0.2)));
    aux[45] = remainder(x0, (x1 + 1));
    aux[46] = rint(x0);
    aux[47] = rootn(x0, p1);
    aux[48] = rsqrt((x0 + 1));
    aux[49] = sin(x0);
    aux[50] = sinh(x0);
    aux[51] = sinpi(x0);
    aux[52] = sqrt(x0);
    aux[53] = tan(x0);
    aux[54] = tanh(x0);
    aux[55] = tanpi(x0);
    aux[56] = trunc(x0);
    dx[0] = RCONST(0.0);
    dx[1] = RCONST(0.0);
}

