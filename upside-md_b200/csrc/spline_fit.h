// Load-time spline fitting in double precision (host).  Produces the same polynomial-form coefficient tables as
// the reference's LayeredPeriodicSpline2D / LayeredClampedSpline1D (src/spline.h:396-516, src/spline.cpp:121-292):
// interpolating cubic splines on an integer grid, periodic (2-D) or clamped with the reference's mirrored end
// condition (1-D), stored per cell as coefficients of {1,f,f^2,f^3} (x) {1,g,g^2,g^3}.
#pragma once
#include <vector>

namespace ub {

// data: (n_layer, nx, ny, ndim) row-major.  returns coeff[(((il*nx+ix)*ny+iy)*ndim+id)*16 + px*4+py]
std::vector<float> fit_periodic_spline_2d(int n_layer, int nx, int ny, int ndim, const double* data);

struct ClampedSpline1D {
    int n_layer, nx, ndim;
    std::vector<float> coeff;   // [((il*(nx-1)+ix)*ndim+id)*4 + p]
    std::vector<float> left, right;   // [il*ndim+id]
};
// data: (n_layer, nx, ndim) row-major
ClampedSpline1D fit_clamped_spline_1d(int n_layer, int nx, int ndim, const double* data);

// B-spline coefficients (length n_coeff = n_values+2) of the clamped interpolating spline, as
// solve_clamped_1d_spline_for_bsplines (spline.cpp:158-189); used by the engine_c_library spline helpers.
std::vector<double> clamped_bspline_coefficients(const std::vector<double>& values);

}  // namespace ub
