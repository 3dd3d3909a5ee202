void getRHS(const realtype t,
            const realtype x_[],
            const realtype p_[],
            realtype dx_[],
            realtype aux_[],
            const realtype w_[]) {

    /* State variables */
    realtype v = x_[0];
    realtype w = x_[1];

    /* Parameters */
    realtype a = p_[0];
    realtype b = p_[1];
    realtype eps = p_[2];
    realtype iapp = p_[3];
    realtype sigma = p_[4];

    /* Noise terms */
    realtype xi = w_[0];

    /* Core equations */
    /*  FitzHugh-Nagumo with additive noise on v: exercises every line kind of the XPP front end; */
    /*  cubic nullcline terms; */
    realtype v3 = v*v*v;
    realtype drive = iapp + sigma * xi;
    realtype slow = eps * (v + a - b * w);
    realtype sat = v*v / (1.0f + v*v*v*v) + pown(w, 7) + pown(v, 0).5 + pow(w, -1.5f);

    /* Auxiliary equations */
    realtype cubic = v3;
    realtype power = sat;

    /* Differential equations */
    realtype dv = v - v3 / 3.0f - w + drive;
    realtype dw = slow;

    /* Auxiliary outputs */
    aux_[0] = cubic;
    aux_[1] = power;

    /* Differential outputs */
    dx_[0] = dv;
    dx_[1] = dw;
}