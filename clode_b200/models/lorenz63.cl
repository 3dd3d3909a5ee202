// Lorenz-63 convection model.
//   state  : X, Y, Z
//   params : p_[0] = rho (Rayleigh), p_[1] = sigma (Prandtl), p_[2] = beta
//   aux    : aux_[0] = dX/dt
// Same system (and the same left-to-right arithmetic, so results are
// bit-comparable) as the reference fixture test/lorenz.cl:20-29.
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    const realtype X = x_[0], Y = x_[1], Z = x_[2];
    const realtype rho = p_[0], sigma = p_[1], beta = p_[2];

    dx_[0] = sigma * (Y - X);
    dx_[1] = rho * X - Y - X * Z;
    dx_[2] = X * Y - beta * Z;

    aux_[0] = dx_[0];
}
