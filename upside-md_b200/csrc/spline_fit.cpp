#include "spline_fit.h"

#include <string>

namespace ub {
namespace {

// basis-to-power matrix: on [i,i+1) S(f) = sum_k beta[i-1+k] * sum_p M[k][p] f^p  (uniform cubic B-spline)
const double M[4][4] = {{1. / 6., -3. / 6., 3. / 6., -1. / 6.},
                        {4. / 6., 0., -6. / 6., 3. / 6.},
                        {1. / 6., 3. / 6., 3. / 6., -3. / 6.},
                        {0., 0., 0., 1. / 6.}};

void thomas(int n, std::vector<double>& sub, std::vector<double>& diag, std::vector<double>& sup, std::vector<double>& rhs) {
    // sub[k] couples row k to k-1 (k>=1); sup[k] couples row k to k+1
    for (int k = 1; k < n; ++k) {
        double m = sub[k] / diag[k - 1];
        diag[k] -= m * sup[k - 1];
        rhs[k] -= m * rhs[k - 1];
    }
    rhs[n - 1] /= diag[n - 1];
    for (int k = n - 2; k >= 0; --k) rhs[k] = (rhs[k] - sup[k] * rhs[k + 1]) / diag[k];
}

// periodic interpolation: (1/6) b[i-1] + (2/3) b[i] + (1/6) b[i+1] = y[i], indices mod n
std::vector<double> periodic_bspline(const std::vector<double>& y) {
    int n = (int)y.size();
    if (n < 3) throw std::string("periodic spline needs at least 3 points");
    const double a = 1. / 6., b = 2. / 3.;
    // Sherman-Morrison on the cyclic system
    double gamma = -b;
    std::vector<double> sub(n, a), sup(n, a), diag(n, b), x(y), u(n, 0.);
    diag[0] = b - gamma;
    diag[n - 1] = b - a * a / gamma;
    std::vector<double> sub2(sub), sup2(sup), diag2(diag);
    thomas(n, sub, diag, sup, x);
    u[0] = gamma;
    u[n - 1] = a;
    thomas(n, sub2, diag2, sup2, u);
    double fact = (x[0] + a * x[n - 1] / gamma) / (1. + u[0] + a * u[n - 1] / gamma);
    for (int i = 0; i < n; ++i) x[i] -= fact * u[i];
    return x;
}

}  // namespace

std::vector<float> fit_periodic_spline_2d(int n_layer, int nx, int ny, int ndim, const double* data) {
    std::vector<float> coeff(size_t(n_layer) * nx * ny * ndim * 16);
    std::vector<double> beta(size_t(nx) * ny), tmp;
    for (int il = 0; il < n_layer; ++il)
        for (int id = 0; id < ndim; ++id) {
            // separable solve: along y for every row, then along x for every column
            for (int ix = 0; ix < nx; ++ix) {
                tmp.assign(ny, 0.);
                for (int iy = 0; iy < ny; ++iy) tmp[iy] = data[((size_t(il) * nx + ix) * ny + iy) * ndim + id];
                auto b = periodic_bspline(tmp);
                for (int iy = 0; iy < ny; ++iy) beta[size_t(ix) * ny + iy] = b[iy];
            }
            for (int iy = 0; iy < ny; ++iy) {
                tmp.assign(nx, 0.);
                for (int ix = 0; ix < nx; ++ix) tmp[ix] = beta[size_t(ix) * ny + iy];
                auto b = periodic_bspline(tmp);
                for (int ix = 0; ix < nx; ++ix) beta[size_t(ix) * ny + iy] = b[ix];
            }
            for (int ix = 0; ix < nx; ++ix)
                for (int iy = 0; iy < ny; ++iy) {
                    double c[4][4] = {};
                    for (int kx = 0; kx < 4; ++kx)
                        for (int ky = 0; ky < 4; ++ky) {
                            double bv = beta[size_t((ix - 1 + kx + nx) % nx) * ny + (iy - 1 + ky + ny) % ny];
                            for (int px = 0; px < 4; ++px)
                                for (int py = 0; py < 4; ++py) c[px][py] += bv * M[kx][px] * M[ky][py];
                        }
                    float* out = &coeff[(((size_t(il) * nx + ix) * ny + iy) * ndim + id) * 16];
                    for (int px = 0; px < 4; ++px)
                        for (int py = 0; py < 4; ++py) out[px * 4 + py] = (float)c[px][py];
                }
        }
    return coeff;
}

static std::vector<double> mirrored_bspline(const std::vector<double>& y) {
    // interior rows (1/6,2/3,1/6); end rows folded with beta[-1]=beta[1], beta[n]=beta[n-2]
    int n = (int)y.size();
    if (n < 3) throw std::string("clamped spline needs at least 3 points");
    std::vector<double> sub(n, 1. / 6.), sup(n, 1. / 6.), diag(n, 2. / 3.), x(y);
    sup[0] *= 2.;
    sub[n - 1] *= 2.;
    thomas(n, sub, diag, sup, x);
    return x;
}

ClampedSpline1D fit_clamped_spline_1d(int n_layer, int nx, int ndim, const double* data) {
    ClampedSpline1D s;
    s.n_layer = n_layer; s.nx = nx; s.ndim = ndim;
    s.coeff.assign(size_t(n_layer) * (nx - 1) * ndim * 4, 0.f);
    s.left.assign(size_t(n_layer) * ndim, 0.f);
    s.right.assign(size_t(n_layer) * ndim, 0.f);
    std::vector<double> y(nx);
    for (int il = 0; il < n_layer; ++il)
        for (int id = 0; id < ndim; ++id) {
            for (int ix = 0; ix < nx; ++ix) y[ix] = data[(size_t(il) * nx + ix) * ndim + id];
            s.left[il * ndim + id] = (float)y[0];
            s.right[il * ndim + id] = (float)y[nx - 1];
            auto b = mirrored_bspline(y);
            auto B = [&](int i) { return i < 0 ? b[1] : (i >= nx ? b[nx - 2] : b[i]); };
            for (int ix = 0; ix < nx - 1; ++ix) {
                double c[4] = {};
                for (int k = 0; k < 4; ++k)
                    for (int p = 0; p < 4; ++p) c[p] += B(ix - 1 + k) * M[k][p];
                for (int p = 0; p < 4; ++p) s.coeff[((size_t(il) * (nx - 1) + ix) * ndim + id) * 4 + p] = (float)c[p];
            }
        }
    return s;
}

std::vector<double> clamped_bspline_coefficients(const std::vector<double>& values) {
    int n = (int)values.size();
    auto b = mirrored_bspline(values);
    std::vector<double> out(n + 2);
    for (int i = 0; i < n; ++i) out[i + 1] = b[i];
    out[0] = out[2];
    out[n + 1] = out[n - 1];
    return out;
}

}  // namespace ub
