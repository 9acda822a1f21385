// jmb_host.h - host-side helpers shared by the translation units of libjmb200.so
#pragma once
#include <algorithm>
#include <string>

// records the message jmb_last_error() returns on this thread; always returns 1
int jmb_set_error(const std::string& message);

// One row of splines::splineDesign(knots, x, ord, outer.ok = TRUE) in band form (R/survfitJM.mvJMbayes.R:205,258;
// de Boor's local recursion): the `ord` values of the basis functions off .. off + ord - 1; rows outside
// [knots[ord-1], knots[n_knots-ord]] are zero, the basis is right-closed at the last knot.
inline void spline_band_row(const double* t, int nk, int ord, double x, int& off, double* vals) {
    const int nb = nk - ord;
    for (int r = 0; r < ord; ++r) vals[r] = 0.0;
    off = 0;
    if (!(x >= t[ord - 1] && x <= t[nb])) return;
    int j = (int)(std::upper_bound(t, t + nk, x) - t) - 1;      // t[j] <= x < t[j+1]
    j = std::min(std::max(j, ord - 1), nb - 1);
    double Bl[16], dl[16], dr[16];
    Bl[0] = 1.0;
    for (int d = 1; d < ord; ++d) {
        dr[d - 1] = t[j + d] - x;
        dl[d - 1] = x - t[j + 1 - d];
        double saved = 0.0;
        for (int r = 0; r < d; ++r) {
            const double den = dr[r] + dl[d - 1 - r];
            const double term = den != 0.0 ? Bl[r] / den : 0.0;
            Bl[r] = saved + dr[r] * term;
            saved = dl[d - 1 - r] * term;
        }
        Bl[d] = saved;
    }
    off = j - (ord - 1);
    for (int r = 0; r < ord; ++r) vals[r] = Bl[r];
}

