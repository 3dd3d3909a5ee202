void getRHS(const realtype t,
            const realtype x_[],
            const realtype p_[],
            realtype dx_[],
            realtype aux_[],
            const realtype w_[]) {
    realtype x = x_[0];
    realtype dilation = p_[0];
    realtype dx = cos((t * dilation));
    dx_[0] = dx;
    aux_[0] = (x + RCONST(1.0));
    aux_[1] = RCONST(1.0);
    aux_[2] = (-RCONST(2.0));
}

