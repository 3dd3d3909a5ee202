// Lactotroph model with an additive current-noise term (one Wiener variable).
//   params : p_[0] = gcal, p_[1] = gsk, p_[2] = gbk, p_[3] = noise amplitude
// Expression order follows the reference sample samples/lactotroph_noise.cl:2-18.
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    const realtype v = x_[0], n = x_[1], f = x_[2], c = x_[3];

    const realtype csq   = c * c;
    const realtype drive = v - RCONST(-75.0);
    const realtype m_inf = RCONST(1.0) / (RCONST(1.0) + exp((RCONST(-20.0) - v) / RCONST(12.0)));
    const realtype n_inf = RCONST(1.0) / (RCONST(1.0) + exp((RCONST(-5.0) - v) / RCONST(10.0)));
    const realtype f_inf = RCONST(1.0) / (RCONST(1.0) + exp((RCONST(-20.0) - v) / RCONST(2.0)));

    const realtype i_ca    = p_[0] * m_inf * (v - RCONST(60.0));
    const realtype i_sk    = p_[1] * csq / (csq + RCONST(0.40) * RCONST(0.40)) * drive;
    const realtype i_bk    = p_[2] * f * drive;
    const realtype i_k     = RCONST(2.0) * n * drive;
    const realtype i_leak  = RCONST(0.050) * (v - RCONST(-50.0));
    const realtype i_noise = p_[3] * w_[0];
    const realtype i_sum   = i_ca + i_sk + i_bk + i_k + i_leak + i_noise;

    dx_[0] = -i_sum / RCONST(10.0);
    dx_[1] = (n_inf - n) / RCONST(30.0);
    dx_[2] = (f_inf - f) / RCONST(8.0);
    dx_[3] = -RCONST(0.010) * (RCONST(0.00150) * i_ca + RCONST(0.20) * c);
    aux_[0] = i_ca;
}
