// face_kernels.cuh -- per-face routines of the two-point flux approximation and of first-order
// upwinding.  One call handles one face; the CUDA kernels run one thread per
// face, the test-only host build loops over the faces.
//
//   tpfa_face   reference numerics/fv/tpfa.py:40-280 (no periodic faces)
//   upwind_face reference numerics/fv/upwind.py:150-300
#pragma once
#include "views.hpp"

namespace pb {

// boundary-condition byte of these routines: bits 0-1 effective code (0 interior, 1 Dirichlet,
// 2 Neumann incl. internal faces, 3 Robin), bit 2 raw is_dir, bit 3 raw is_neu (the reference uses
// the raw flags for the pressure-trace terms, tpfa.py:221-225, and for upwinding)
#define PB_FBC_RAW_DIR 4
#define PB_FBC_RAW_NEU 8

struct TpfaOut {
    // per (face, cell) entry of cell_faces in CSR-by-face order: position fc_ptr[f] + rank
    double *flux, *bpc;   // nnz
    double *vs, *bpvs;    // nnz * vdim, entry-major
    double *bflux_diag, *bpf_diag;  // nf (diagonals; bound_flux only on boundary faces)
};

// face_cells: 2 per face, (cell << 1) | (sign < 0), -1 = none.  fc_ptr: CSR row pointer of the
// face x cell pattern with ascending columns (scipy's cell_faces.tocsr()).
PB_HD void tpfa_face(int64_t f, const GeoView &G, const double *perm, int64_t perm_cs, int64_t perm_es,
                     const uint8_t *bc, const int32_t *face_cells, const int32_t *fc_ptr, int vdim,
                     const TpfaOut &o) {
    int cell[2];
    double sg[2], d[2][3], thalf[2];
    int ncell = 0;
    double inv_sum = 0.0;
    for (int sd = 0; sd < 2; ++sd) {
        const int32_t enc = face_cells[2 * f + sd];
        if (enc < 0) continue;
        const int64_t c = enc >> 1;
        const double s = (enc & 1) ? -1.0 : 1.0;
        double n[3] = {0.0, 0.0, 0.0}, dd[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < 3; ++i) {  // the reference sums over all three coordinates (tpfa.py:163-175)
            n[i] = G.fnorm[i * G.face_cs + f * G.face_es] * s;
            dd[i] = G.fcent[i * G.face_cs + f * G.face_es] - G.ccent[i * G.cell_cs + c * G.cell_es];
        }
        double num = 0.0, den = 0.0;
        for (int i = 0; i < 3; ++i) {
            double nk = 0.0;
            for (int j = 0; j < 3; ++j) nk += perm[(i * 3 + j) * perm_cs + c * perm_es] * n[j];
            num += nk * dd[i];
            den += dd[i] * dd[i];
        }
        thalf[ncell] = num / den;
        inv_sum += 1.0 / thalf[ncell];
        cell[ncell] = (int)c;
        sg[ncell] = s;
        for (int i = 0; i < 3; ++i) d[ncell][i] = dd[i];
        ++ncell;
    }
    const double t_full = 1.0 / inv_sum;
    const int code = bc[f] & 3;
    const bool raw_dir = (bc[f] & PB_FBC_RAW_DIR) != 0, raw_neu = (bc[f] & PB_FBC_RAW_NEU) != 0;
    const double t = (code == 2) ? 0.0 : t_full;          // tpfa.py:200
    const int32_t p0 = fc_ptr[f];
    for (int k = 0; k < ncell; ++k) {
        // position inside the row: ascending cell index
        const int rank = (ncell == 2 && cell[k] > cell[1 - k]) ? 1 : 0;
        const int64_t p = p0 + rank;
        if (o.flux) o.flux[p] = t * sg[k];
        if (o.bpc) o.bpc[p] = raw_neu ? 1.0 : 0.0;
        for (int i = 0; i < vdim; ++i) {
            if (o.vs) o.vs[p * vdim + i] = t * sg[k] * d[k][i];
            if (o.bpvs) o.bpvs[p * vdim + i] = raw_neu ? d[k][i] : 0.0;
        }
    }
    if (o.bflux_diag) {
        double tb = 0.0;                                   // tpfa.py:193-197
        if (ncell == 1) tb = (code == 1) ? -t_full * sg[0] : (code == 2 ? sg[0] : 0.0);
        o.bflux_diag[f] = tb;
    }
    if (o.bpf_diag) o.bpf_diag[f] = raw_neu ? -1.0 / t_full : (raw_dir ? 1.0 : 0.0);  // tpfa.py:221-223
}

// up_col[f]: upstream cell of the face or -1 when the face is removed from the upwind matrix
// (Neumann faces, Dirichlet inflow faces); the two boundary diagonals as in upwind.py:282-300.
PB_HD void upwind_face(int64_t f, const double *darcy_flux, const uint8_t *bc, const int32_t *face_cells,
                       int32_t *up_col, double *neu_diag, double *dir_diag) {
    int cpos = -1, cneg = -1;  // cell on the + / - side of the face (Grid.cell_faces_as_dense)
    double sgn_div = 0.0;
    for (int sd = 0; sd < 2; ++sd) {
        const int32_t enc = face_cells[2 * f + sd];
        if (enc < 0) continue;
        if (enc & 1) { cneg = enc >> 1; sgn_div -= 1.0; }
        else { cpos = enc >> 1; sgn_div += 1.0; }
    }
    const bool pos = darcy_flux[f] >= 0.0;                 // np.sign(q) >= 0: zero counts as positive
    const bool raw_dir = (bc[f] & PB_FBC_RAW_DIR) != 0, raw_neu = (bc[f] & PB_FBC_RAW_NEU) != 0;
    const bool inflow = raw_dir && ((pos && cpos < 0) || (!pos && cneg < 0));
    up_col[f] = (raw_neu || inflow) ? -1 : (pos ? cpos : cneg);
    neu_diag[f] = raw_neu ? sgn_div : 0.0;
    dir_diag[f] = inflow ? 1.0 : 0.0;
}

}  // namespace pb
