// Non-autonomous test system x' = cos(dilation * t) with three aux outputs.
// It is the system the reference's observer tests integrate
// (test/test_features.py:10-21, test/test_aux_values.py:10-24).
//   state : x    params : p_[0] = dilation    aux : x + 1, +1, -2
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    dx_[0] = cos(t * p_[0]);
    aux_[0] = x_[0] + RCONST(1.0);
    aux_[1] = RCONST(1.0);
    aux_[2] = -RCONST(2.0);
}
