// Van der Pol oscillator, x'' - mu (1 - x^2) x' + x = 0 as a first-order system.
//   state  : x, y (= x')     params : p_[0] = mu     aux : none
// Arithmetic order follows the reference fixture test/van_der_pol_oscillator.cl:11-15.
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    const realtype mu = p_[0];
    const realtype pos = x_[0], vel = x_[1];

    dx_[0] = vel;
    dx_[1] = mu * (1 - pos * pos) * vel - pos;
}
