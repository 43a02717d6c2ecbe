// views.hpp -- plain-pointer views of the plan / geometry / parameters that the node routines
// index.  The same structs are filled with device pointers (api.cu) or, in the test-only kernel
// emulation harness (tests/emu), with host pointers.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

struct PlanView {
    int nd;
    int64_t nc, nf, nn;
    const int32_t *fn_indptr;
    const int32_t *node_sc_ptr, *sc_cell;
    const int32_t *node_sf_ptr, *sf_face;
    const uint32_t *sf_sides;
    const uint16_t *sf_bloc, *slot_sf;
    const int32_t *node_nb;
    const int32_t *sc_ncn;
    const int64_t *posfc_ptr, *posfb_ptr, *poscc_ptr, *poscb_ptr;
    const int32_t *pos_fc, *pos_fb, *pos_cc, *pos_cb;
    const int32_t *fc_indptr, *fb_indptr, *cc_indptr, *cb_indptr;
};

// geometry as pp.Grid stores it: (3, n) row-major -> component i of entity e at [i*n + e]
// Vector arrays are addressed as base[component * cs + entity * es]: (cs, es) = (n, 1) for the
// reference's (3, n) row-major layout (what the host emulation passes), (1, 3) after the device-side
// repack into entity-major records (one 24-byte record per point instead of three sectors).
struct GeoView {
    const double *nodes, *fnorm, *fcent, *farea, *ccent, *cvol;
    int64_t node_cs, node_es, face_cs, face_es, cell_cs, cell_es;
};

struct MpfaParams {
    const double *perm;   // (3,3,nc): entry (i,j) of cell c at [(3i+j)*perm_cs + c*perm_es]
    const uint8_t *bc;    // nf
    const double *robw;   // nf or null
    double eta;
    int64_t perm_cs, perm_es;
};

struct MpfaOut {
    double *flux, *bflux, *bpc, *bpf, *vs, *bpvs;  // any may be null
};

struct MpsaParams {
    const double *stiff;  // (9,9,nc)
    const uint8_t *bc;    // (nd,nf)
    const double *robw;   // (nd,nd,nf) or null
    const double *basis;  // (nd,nd,nf) boundary-condition basis (bc.basis), null = identity
    double eta;
    int n_alpha;
    const double *alpha;  // n_alpha x (3,3,nc)
    int64_t stiff_cs, stiff_es;  // entry (p9,q9) of cell c at [(9*p9+q9)*stiff_cs + c*stiff_es]
    int64_t alpha_cs, alpha_es, alpha_stride;  // tensor q at alpha + q*alpha_stride
};

#define PB_MAX_ALPHA 4
struct MpsaOut {
    double *stress, *bstress, *bdc, *bdf;
    double *dd[PB_MAX_ALPHA], *bdd[PB_MAX_ALPHA], *sg[PB_MAX_ALPHA], *cons[PB_MAX_ALPHA],
        *bdp[PB_MAX_ALPHA];
};

}  // namespace pb
