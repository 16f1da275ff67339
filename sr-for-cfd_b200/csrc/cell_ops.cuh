// cell_ops.cuh -- one relaxation update of one cell, shared by every sweep order.
//
// Each function takes the stencil VALUES (the caller decides which of them are already updated in
// this sweep -- that is what distinguishes Gauss-Seidel, Jacobi and red-black), and returns the new
// cell value; R (the residual that feeds the rms norm) comes back by reference.
// Names: c = (i,j); ip/im = (i+1,j)/(i-1,j); jp/jm = (i,j+1)/(i,j-1); ip2.. = second neighbours.
// Face fluxes: fE = Ff[0] (i+1/2), fN = Ff[1] (j+1/2), fW = Ff[2] (i-1/2), fS = Ff[3] (j-1/2).
#pragma once
#include "common.cuh"

namespace srcfd {

// LDC.py:232-237 diffusive_flux
__device__ __forceinline__ double diffusive_flux(double c, double ip, double im, double jp, double jm,
                                                 const Consts& K) {
    return K.volp * ((ip - 2.0 * c + im) / K.dx2 + (jp - 2.0 * c + jm) / K.dy2);
}

// LDC.py:300-310 (solve_pressure body).  rhs = rho/dt*(fE+fN+fW+fS), precomputed per outer iteration.
__device__ __forceinline__ double pressure_cell(double c, double ip, double im, double jp, double jm,
                                                double rhs, const Consts& K, double& R) {
    const double Fd = diffusive_flux(c, ip, im, jp, jm, K);
    R = rhs - Fd;
    return c + R / K.ap_d;
}

__device__ __forceinline__ double momentum_finish(double c, double vold, double Fc, double ap_c, double Fd,
                                                  const Consts& K, double& R) {
    // LDC.py:260-263
    R = -(K.volp_dt * (c - vold) + Fc + K.neg_nu * Fd);
    const double ap = K.volp_dt + ap_c + K.neg_nu_ap_d;
    return c + R / ap;
}

// LDC.py:156-188 simple_upwind + LDC.py:277-286
__device__ __forceinline__ double upwind_cell(double c, double ip, double im, double jp, double jm,
                                              double vold, double fE, double fN, double fW, double fS,
                                              const Consts& K, double& R) {
    double ue, uw, un, us, sum_flux = 0.0;
    if (fE >= 0) { ue = c; sum_flux += fE; } else ue = ip;
    if (fW >= 0) { uw = c; sum_flux += fW; } else uw = im;
    if (fN >= 0) { un = c; sum_flux += fN; } else un = jp;
    if (fS >= 0) { us = c; sum_flux += fS; } else us = jm;
    const double Fc = ue * fE + uw * fW + un * fN + us * fS;
    const double ap_c = sum_flux * K.volp;
    const double Fd = diffusive_flux(c, ip, im, jp, jm, K);
    return momentum_finish(c, vold, Fc, ap_c, Fd, K, R);
}

// LDC.py:190-230 quick_scheme + LDC.py:255-264
__device__ __forceinline__ double quick_cell(double c, double ip, double im, double jp, double jm,
                                             double ip2, double im2, double jp2, double jm2,
                                             double vold, double fE, double fN, double fW, double fS,
                                             const Consts& K, double& R) {
    double ue, uw, un, us, sum_flux = 0.0;
    if (fE >= 0) { ue = 0.75 * c + 0.375 * ip - 0.125 * im; sum_flux += 0.75 * fE; }
    else         { ue = 0.75 * ip + 0.375 * c - 0.125 * ip2; sum_flux += 0.375 * fE; }
    if (fW >= 0) { uw = 0.75 * c + 0.375 * im - 0.125 * ip; sum_flux += 0.75 * fW; }
    else         { uw = 0.75 * im + 0.375 * c - 0.125 * im2; sum_flux += 0.375 * fW; }
    if (fN >= 0) { un = 0.75 * c + 0.375 * jp - 0.125 * jm; sum_flux += 0.75 * fN; }
    else         { un = 0.75 * jp + 0.375 * c - 0.125 * jp2; sum_flux += 0.375 * fN; }
    if (fS >= 0) { us = 0.75 * c + 0.375 * jm - 0.125 * jp; sum_flux += 0.75 * fS; }
    else         { us = 0.75 * jm + 0.375 * c - 0.125 * jm2; sum_flux += 0.375 * fS; }
    const double Fc = ue * fE + uw * fW + un * fN + us * fS;
    const double ap_c = sum_flux * K.volp;
    const double Fd = diffusive_flux(c, ip, im, jp, jm, K);
    return momentum_finish(c, vold, Fc, ap_c, Fd, K, R);
}

}  // namespace srcfd
