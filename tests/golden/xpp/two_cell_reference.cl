void getRHS(const realtype t,
            const realtype x_[],
            const realtype p_[],
            realtype dx_[],
            realtype aux_[],
            const realtype w_[]) {

    /* State variables */
    realtype x1 = x_[0];
    realtype x2 = x_[1];
    realtype y1 = x_[2];
    realtype y2 = x_[3];

    /* Parameters */
    realtype g = p_[0];
    realtype vth = p_[1];
    realtype k = p_[2];
    realtype tau1 = p_[3];
    realtype tau2 = p_[4];

    /* Noise terms */

    /* Core equations */
    realtype s1 = 1.0f / (1.0f + exp(-(x1 - vth) / k));
    realtype s2 = 1.0f / (1.0f + exp(-(x2 - vth) / k));

    /* Auxiliary equations */
    realtype syn1 = s1;

    /* Differential equations */
    realtype dx1 = -x1 + g * s2 * (2.0f - x1);
    realtype dx2 = -x2 + g * s1 * (2.0f - x2);
    realtype dy1 = (s1 - y1) / tau1;
    realtype dy2 = (s2 - y2) / tau2;

    /* Auxiliary outputs */
    aux_[0] = syn1;

    /* Differential outputs */
    dx_[0] = dx1;
    dx_[1] = dx2;
    dx_[2] = dy1;
    dx_[3] = dy2;
}