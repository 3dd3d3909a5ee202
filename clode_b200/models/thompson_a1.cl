// ORNL/Thompson test problem A1 (https://www.osti.gov/biblio/6111421):
// falling body with linear drag; exact solution
//   y1 = 4 (t + exp(-8 t)/8 - 1/8),  y2 = 4 (1 - exp(-8 t))   for m=1/4, w=8, k=2.
//   state : y1, y2    params : m, w, k, H    aux : g1 = y1 - H
// Arithmetic order follows the reference fixture test/ornl_thompson_a1.cl:18-23.
void getRHS(const realtype t, const realtype x_[], const realtype p_[],
            realtype dx_[], realtype aux_[], const realtype w_[])
{
    const realtype mass = p_[0], weight = p_[1], drag = p_[2], height = p_[3];

    dx_[0] = x_[1];
    dx_[1] = (weight - drag * x_[1]) / mass;
    aux_[0] = x_[0] - height;
}
