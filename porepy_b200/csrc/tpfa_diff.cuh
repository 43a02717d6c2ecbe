// tpfa_diff.cuh -- differentiable two-point flux: face transmissibilities and their derivative with respect to the
// cell permeabilities, one call per face (reference numerics/fv/tpfa.py:281-760
// DifferentiableTpfa + the AD expression of models/constitutive_laws.py:1544-1583 that the reference re-evaluates in
// every Newton iteration of a model with a solution-dependent permeability):
//
//   half-face h = (face f, cell c):   t_h = d_h^T K_c n_f / |d_h|^2,   d_h = x_f - x_c      (tpfa.py:617-660)
//   face:                             T_f = 1 / sum_h s_h / t_h,        s_h = cell_faces[f, c]
//   derivative:                       dT_f / dK_c[3a+b] = T_f^2 s_h / t_h^2 * d_a n_b / |d_h|^2
//
// K_c is the 3 x 3 tensor of cell c at k[9c .. 9c+8] row-major (the layout of the reference's k_c AD vector).
// Half-faces are numbered by face, then by cell (scipy.sparse.find order, as in the reference): half-face
// fc_ptr[f] + rank, rank = position of the cell among the face's cells in ascending order.
#pragma once
#include "views.hpp"

namespace pb {

// face_cells: 2 per face, (cell << 1) | (sign < 0), -1 = none, slot 0 = the smaller cell index
PB_HD void tpfa_diff_face(int64_t f, const GeoView &G, const double *k, const int32_t *face_cells,
                          const int32_t *fc_ptr, double *t_hf, double *T, double *dT_dk) {
    double th[2], sg[2], geo[2][9];
    int n = 0;
    double inv_sum = 0.0;
    for (int sd = 0; sd < 2; ++sd) {
        const int32_t enc = face_cells[2 * f + sd];
        if (enc < 0) continue;
        const int64_t c = enc >> 1;
        const double s = (enc & 1) ? -1.0 : 1.0;
        double nv[3], d[3], dist = 0.0;
        for (int i = 0; i < 3; ++i) {
            nv[i] = G.fnorm[i * G.face_cs + f * G.face_es];
            d[i] = G.fcent[i * G.face_cs + f * G.face_es] - G.ccent[i * G.cell_cs + c * G.cell_es];
            dist += d[i] * d[i];
        }
        // the reference forms  (diag(1/dist) d_vec n) @ k_c : entry (a, b) of the geometric factor is d_a n_b / dist
        double t = 0.0;
        for (int a = 0; a < 3; ++a) {
            double kn = 0.0;
            for (int b = 0; b < 3; ++b) {
                geo[n][3 * a + b] = d[a] * nv[b] / dist;
                kn += nv[b] * k[9 * c + 3 * a + b];
            }
            t += d[a] * kn;
        }
        t /= dist;
        th[n] = t;
        sg[n] = s;
        inv_sum += s / t;
        ++n;
    }
    const double Tf = 1.0 / inv_sum;
    T[f] = Tf;
    const int64_t h0 = fc_ptr[f];
    for (int j = 0; j < n; ++j) {
        t_hf[h0 + j] = th[j];
        const double w = Tf * Tf * sg[j] / (th[j] * th[j]);
        for (int e = 0; e < 9; ++e) dT_dk[(h0 + j) * 9 + e] = w * geo[j][e];
    }
}

}  // namespace pb
