// Chay-Keizer style beta-cell bursting model, 3 variables.
//   state  : v (mV), n (K activation), c (cytosolic Ca)
//   params : p_[0] = gca, p_[1] = gkca, p_[2] = kpmca
// Expression order follows the reference example examples/chay_keizer.cl:31-43.
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    const realtype v = x_[0], n = x_[1], c = x_[2];
    const realtype g_ca = p_[0], g_kca = p_[1], k_pmca = p_[2];

    const realtype g_k   = RCONST(3000.0);
    const realtype v_ca  = RCONST(25.0),  v_k = RCONST(-75.0);
    const realtype c_m   = RCONST(5300.0);
    const realtype alpha = RCONST(4.5e-6), f_cyt = RCONST(0.01), k_d = RCONST(0.4);
    const realtype v_m   = RCONST(-20.0), s_m = RCONST(12.0);
    const realtype v_n   = RCONST(-16.0), s_n = RCONST(5.0), tau_n = RCONST(20.0);

    const realtype m_inf = RCONST(1.0) / (RCONST(1.0) + exp((v_m - v) / s_m));
    const realtype n_inf = RCONST(1.0) / (RCONST(1.0) + exp((v_n - v) / s_n));
    const realtype w_inf = pown(c, 2) / (pown(c, 2) + pown(k_d, 2));

    const realtype i_ca  = g_ca * m_inf * (v - v_ca);
    const realtype i_k   = g_k * n * (v - v_k);
    const realtype i_kca = g_kca * w_inf * (v - v_k);

    dx_[0] = -(i_ca + i_k + i_kca) / c_m;
    dx_[1] = (n_inf - n) / tau_n;
    dx_[2] = f_cyt * (-alpha * i_ca - k_pmca * c);
}
