// emu.cpp -- TEST INFRASTRUCTURE ONLY: runs the *same* per-node routines the CUDA kernels
// run (porepy_b200/csrc/node_kernels.cuh) on the host with a 1-thread team, so the arithmetic
// and the plan indexing can be checked against the oracle / golden fixtures on a box without a
// GPU.  Built by tests/emu_binding.py with g++ into tests/emu/_emu.so; the product
// (porepy_b200, libporeb200.so) never builds, links or loads it.
#include <climits>
#include <cstdio>
#include <string>
#include <vector>

#include "../../porepy_b200/csrc/node_kernels.cuh"
#include "../../porepy_b200/csrc/plan_host.hpp"
#if __has_include("../../porepy_b200/csrc/mpsa_node.cuh")
#include "../../porepy_b200/csrc/mpsa_node.cuh"
#include "../../porepy_b200/csrc/face_kernels.cuh"
#include "../../porepy_b200/csrc/geometry_kernels.cuh"
#include "../../porepy_b200/csrc/tpfa_diff.cuh"
#define HAVE_MPSA 1
#endif

using namespace pb;

struct Emu {
    HostPlan P;
    std::string err;
};

static PlanView view_of(const HostPlan &P) {
    PlanView v;
    v.nd = P.nd; v.nc = P.nc; v.nf = P.nf; v.nn = P.nn;
    v.fn_indptr = P.fn_indptr.data();
    v.node_sc_ptr = P.node_sc_ptr.data(); v.sc_cell = P.sc_cell.data();
    v.node_sf_ptr = P.node_sf_ptr.data(); v.sf_face = P.sf_face.data();
    v.sf_sides = P.sf_sides.data(); v.sf_bloc = P.sf_bloc.data(); v.slot_sf = P.slot_sf.data();
    v.node_nb = P.node_nb.data(); v.sc_ncn = P.sc_ncn.data();
    v.posfc_ptr = P.posfc_ptr.data(); v.posfb_ptr = P.posfb_ptr.data();
    v.poscc_ptr = P.poscc_ptr.data(); v.poscb_ptr = P.poscb_ptr.data();
    v.pos_fc = P.pos_fc.data(); v.pos_fb = P.pos_fb.data();
    v.pos_cc = P.pos_cc.data(); v.pos_cb = P.pos_cb.data();
    v.fc_indptr = P.pat[0].indptr.data(); v.fb_indptr = P.pat[1].indptr.data();
    v.cc_indptr = P.pat[2].indptr.data(); v.cb_indptr = P.pat[3].indptr.data();
    return v;
}

extern "C" {

int emu_create(int nd, int64_t nc, int64_t nf, int64_t nn, const int32_t *cf_indptr,
               const int32_t *cf_indices, const int8_t *cf_data, const int32_t *fn_indptr,
               const int32_t *fn_indices, void **out) {
    Emu *e = new Emu;
    int rc = build_host_plan(nd, nc, nf, nn, cf_indptr, cf_indices, cf_data, fn_indptr, fn_indices,
                             e->P, e->err);
    if (rc) { fprintf(stderr, "emu_create: %s\n", e->err.c_str()); delete e; return rc; }
    *out = e;
    return 0;
}
void emu_destroy(void *h) { delete (Emu *)h; }

int emu_pattern_size(void *h, int which, int64_t *nrows, int64_t *nnz) {
    Emu *e = (Emu *)h;
    *nrows = e->P.pat[which].nrows;
    *nnz = e->P.pat[which].nnz();
    return 0;
}
int emu_pattern_get(void *h, int which, int32_t *indptr, int32_t *indices) {
    Emu *e = (Emu *)h;
    const Csr &c = e->P.pat[which];
    std::copy(c.indptr.begin(), c.indptr.end(), indptr);
    std::copy(c.indices.begin(), c.indices.end(), indices);
    return 0;
}

int emu_mpfa(void *h, const double *nodes, const double *fnorm, const double *fcent,
             const double *farea, const double *ccent, const double *cvol, const double *perm,
             const uint8_t *bc, const double *robw, double eta, double *flux, double *bflux,
             double *bpc, double *bpf, double *vs, double *bpvs) {
    Emu *e = (Emu *)h;
    PlanView P = view_of(e->P);
    GeoView G{nodes, fnorm, fcent, farea, ccent, cvol, P.nn, 1, P.nf, 1, P.nc, 1};
    MpfaParams prm{perm, bc, robw, eta, P.nc, 1};
    MpfaOut o{flux, bflux, bpc, bpf, vs, bpvs};
    int err = INT_MAX;
    int64_t need = 0;
    for (int64_t s = 0; s < P.nn; ++s) {
        int nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s], nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        need = std::max(need, mpfa_A_doubles(P.nd, nsf, nsc, P.node_nb[s]) +
                                  mpfa_rest_doubles(P.nd, nsf, nsc, P.node_nb[s]) + nsf);
    }
    std::vector<double> sm(need + 8);
    CpuTeam t;
    for (int64_t s = 0; s < P.nn; ++s) {
        int nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s], nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        double *rest = sm.data() + mpfa_A_doubles(P.nd, nsf, nsc, P.node_nb[s]);
        double *scr = rest + mpfa_rest_doubles(P.nd, nsf, nsc, P.node_nb[s]);
        if (P.nd == 3) mpfa_node<3, SmemGJ>(t, P, G, prm, o, s, sm.data(), rest, scr, &err);
        else mpfa_node<2, SmemGJ>(t, P, G, prm, o, s, sm.data(), rest, scr, &err);
    }
    return err == INT_MAX ? 0 : 2;
}

int emu_tpfa(void *h, const double *fnorm, const double *fcent, const double *ccent, const double *perm,
             const uint8_t *bc, const int32_t *fc_ptr, int vdim, double *flux, double *bpc, double *vs,
             double *bpvs, double *bflux_diag, double *bpf_diag) {
    Emu *e = (Emu *)h;
    const HostPlan &H = e->P;
    GeoView G{nullptr, fnorm, fcent, nullptr, ccent, nullptr, H.nn, 1, H.nf, 1, H.nc, 1};
    TpfaOut o{flux, bpc, vs, bpvs, bflux_diag, bpf_diag};
    for (int64_t f = 0; f < H.nf; ++f) {
        tpfa_face(f, G, perm, H.nc, 1, bc, H.face_cells.data(), fc_ptr, vdim, o);
    }
    return 0;
}

// ---- Grid.compute_geometry (3-D): the per-face and per-cell routines of geometry.cu, (3, n) row-major arrays
int emu_geometry_3d(int64_t nc, int64_t nf, int64_t nn, const int32_t *cf_ip, const int32_t *cf_ix, const int8_t *cf_da,
                    const int32_t *fn_ip, const int32_t *fn_ix, const double *nodes, double *fnorm, double *fcent,
                    double *farea, double *ccent, double *cvol) {
    GeomOut o{fnorm, fcent, farea, ccent, cvol, nf, 1, nc, 1};
    for (int64_t f = 0; f < nf; ++f) geom_face(f, fn_ip, fn_ix, nodes, nn, 1, o);
    bool ok = true;
    for (int64_t c = 0; c < nc; ++c) ok &= geom_cell(c, cf_ip, cf_ix, cf_da, fn_ip, fn_ix, nodes, nn, 1, o);
    return ok ? 0 : 1;
}

// ---- per-face schemes on a bare face grid (any dimension): the face -> cell table is built here the way
// face.cu builds it on the device (slot 0 = smaller cell index)
static std::vector<int32_t> face_cells_of(int64_t nc, int64_t nf, const int32_t *cf_ip, const int32_t *cf_ix,
                                          const int8_t *cf_da) {
    std::vector<int32_t> fc(2 * nf, -1);
    for (int64_t c = 0; c < nc; ++c)
        for (int q = cf_ip[c]; q < cf_ip[c + 1]; ++q) {
            const int32_t f = cf_ix[q];
            const int32_t enc = (int32_t)((c << 1) | (cf_da[q] < 0 ? 1 : 0));
            if (fc[2 * f] < 0) fc[2 * f] = enc; else fc[2 * f + 1] = enc;
        }
    return fc;
}

int emu_facegrid_tpfa(int64_t nc, int64_t nf, const int32_t *cf_ip, const int32_t *cf_ix, const int8_t *cf_da,
                      const double *fnorm, const double *fcent, const double *ccent, const double *perm,
                      const uint8_t *bc, const int32_t *fc_ptr, int vdim, double *flux, double *bpc, double *vs,
                      double *bpvs, double *bflux_diag, double *bpf_diag) {
    std::vector<int32_t> fc = face_cells_of(nc, nf, cf_ip, cf_ix, cf_da);
    GeoView G{nullptr, fnorm, fcent, nullptr, ccent, nullptr, 0, 1, nf, 1, nc, 1};
    TpfaOut o{flux, bpc, vs, bpvs, bflux_diag, bpf_diag};
    for (int64_t f = 0; f < nf; ++f) tpfa_face(f, G, perm, nc, 1, bc, fc.data(), fc_ptr, vdim, o);
    return 0;
}

int emu_facegrid_tpfa_diff(int64_t nc, int64_t nf, const int32_t *cf_ip, const int32_t *cf_ix, const int8_t *cf_da,
                           const double *fnorm, const double *fcent, const double *ccent, const double *k,
                           const int32_t *fc_ptr, double *t_hf, double *T, double *dT_dk) {
    std::vector<int32_t> fc = face_cells_of(nc, nf, cf_ip, cf_ix, cf_da);
    GeoView G{nullptr, fnorm, fcent, nullptr, ccent, nullptr, 0, 1, nf, 1, nc, 1};
    for (int64_t f = 0; f < nf; ++f) tpfa_diff_face(f, G, k, fc.data(), fc_ptr, t_hf, T, dT_dk);
    return 0;
}

int emu_facegrid_upwind(int64_t nc, int64_t nf, const int32_t *cf_ip, const int32_t *cf_ix, const int8_t *cf_da,
                        const double *darcy_flux, const uint8_t *bc, int32_t *up_col, double *neu_diag,
                        double *dir_diag) {
    std::vector<int32_t> fc = face_cells_of(nc, nf, cf_ip, cf_ix, cf_da);
    for (int64_t f = 0; f < nf; ++f) upwind_face(f, darcy_flux, bc, fc.data(), up_col, neu_diag, dir_diag);
    return 0;
}

int emu_upwind(void *h, const double *darcy_flux, const uint8_t *bc, int32_t *up_col, double *neu_diag,
               double *dir_diag) {
    Emu *e = (Emu *)h;
    const HostPlan &H = e->P;
    for (int64_t f = 0; f < H.nf; ++f)
        upwind_face(f, darcy_flux, bc, H.face_cells.data(), up_col, neu_diag, dir_diag);
    return 0;
}

#ifdef HAVE_MPSA
int emu_mpsa(void *h, const double *nodes, const double *fnorm, const double *fcent,
             const double *farea, const double *ccent, const double *cvol, const double *stiff,
             const uint8_t *bc, const double *robw, const double *basis, double eta, int n_alpha,
             const double *alpha, double *stress, double *bstress, double *bdc, double *bdf, double **dd, double **bdd,
             double **sg, double **cons, double **bdp) {
    Emu *e = (Emu *)h;
    PlanView P = view_of(e->P);
    GeoView G{nodes, fnorm, fcent, farea, ccent, cvol, P.nn, 1, P.nf, 1, P.nc, 1};
    MpsaParams prm{stiff, bc, robw, basis, eta, n_alpha, alpha, P.nc, 1, P.nc, 1, 9 * P.nc};
    MpsaOut o{};
    o.stress = stress; o.bstress = bstress; o.bdc = bdc; o.bdf = bdf;
    for (int a = 0; a < n_alpha; ++a) {
        o.dd[a] = dd[a]; o.bdd[a] = bdd[a]; o.sg[a] = sg[a]; o.cons[a] = cons[a]; o.bdp[a] = bdp[a];
    }
    int err = INT_MAX;
    int64_t need = 0;
    for (int64_t s = 0; s < P.nn; ++s) {
        int nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s], nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        need = std::max(need, mpsa_A_doubles(P.nd, nsf, nsc, P.node_nb[s], n_alpha) +
                                  mpsa_rest_doubles(P.nd, nsf, nsc, P.node_nb[s], n_alpha) + (int64_t)nsf * P.nd);
    }
    std::vector<double> sm(need + 8);
    CpuTeam t;
    for (int64_t s = 0; s < P.nn; ++s) {
        int nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s], nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        double *rest = sm.data() + mpsa_A_doubles(P.nd, nsf, nsc, P.node_nb[s], n_alpha);
        double *scr = rest + mpsa_rest_doubles(P.nd, nsf, nsc, P.node_nb[s], n_alpha);
        if (P.nd == 3) mpsa_node<3, SmemGJ>(t, P, G, prm, o, s, sm.data(), rest, scr, &err);
        else mpsa_node<2, SmemGJ>(t, P, G, prm, o, s, sm.data(), rest, scr, &err);
    }
    return err == INT_MAX ? 0 : 2;
}
#endif
}
