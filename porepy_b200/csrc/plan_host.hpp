// plan_host.hpp -- sub-cell topology ("interaction regions") and output sparsity patterns.
//
// Integer-only host construction of everything the assembly kernels index with.  One
// interaction region = one grid node s; it owns
//   * its sub-cells (K,s), ordered by ascending cell index,
//   * its sub-faces (f,s), ordered by ascending (face, position in face_nodes),
//   * per sub-cell the nd sub-faces of K meeting in s, with the sign cell_faces[f,K],
//   * per sub-face its one or two sides; side 1 is the smaller cell index, which is the side
//     the reference takes the flux / traction from (numerics/fv/_fvutils.py:143-150,163).
// What the reference builds with np.lexsort + sparse products in SubcellTopology.__init__
// (numerics/fv/_fvutils.py:51-172) is done here by a counting sort over nodes.
//
// Output patterns are structural (the reference's are value dependent because scipy's SpGEMM
// drops exact zeros): row f of FACE_CELL holds every cell sharing a node with face f, etc.
#pragma once
#include <algorithm>
#if defined(_OPENMP)
#include <omp.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace pb {

// threads of the parallel regions below (set by pb_plan_create; tests/emu keeps the default)
static int g_plan_threads = 0;
static inline int plan_threads() {
#if defined(_OPENMP)
    return g_plan_threads > 0 ? g_plan_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

struct Csr {
    int64_t nrows = 0, ncols = 0;
    std::vector<int32_t> indptr, indices;
    int64_t nnz() const { return (int64_t)indices.size(); }
};

struct HostPlan {
    int nd = 0;
    int64_t nc = 0, nf = 0, nn = 0;
    int64_t S = 0, U = 0, H = 0;  // sub-cells, sub-faces, sub-half-faces
    std::vector<int32_t> fn_indptr;            // nf+1 (nodes per face)
    std::vector<int32_t> node_sc_ptr, sc_cell; // nn+1, S   (node-major)
    std::vector<int32_t> node_sf_ptr, sf_face; // nn+1, U   (node-major)
    std::vector<uint32_t> sf_sides;            // U: side1 | side2 << 16, side = k*nd+m, 0xFFFF none
    std::vector<uint16_t> sf_bloc;             // U: index among the node's boundary sub-faces or 0xFFFF
    std::vector<uint16_t> slot_sf;             // S*nd: (local sub-face << 1) | (sign < 0)
    std::vector<int32_t> node_nb;              // nn: boundary sub-faces at the node
    std::vector<int32_t> sc_ncn;               // nc: nodes per cell (= sub-cells per cell)
    Csr pat[4];                                // PB_PAT_*
    // per-node maps local (row, col) -> position in the base pattern's data array
    std::vector<int64_t> posfc_ptr, posfb_ptr, poscc_ptr, poscb_ptr;  // nn+1 each
    std::vector<int32_t> pos_fc, pos_fb, pos_cc, pos_cb;
    std::vector<int32_t> nbf_ptr, nbf_idx;     // boundary faces of each node (local order)
    std::vector<int32_t> cn_ptr, cn_idx;       // nodes of each cell
    std::vector<int32_t> face_cells;           // 2*nf: (cell << 1) | (cell_faces sign < 0), -1 = no second cell
    int32_t max_nsf = 0, max_nsc = 0, max_nb = 0;
};

// returns 0 ok, 1 invalid, 3 unsupported cell type; message in err
inline int build_host_plan(int nd, int64_t nc, int64_t nf, int64_t nn, const int32_t *cf_indptr,
                           const int32_t *cf_indices, const int8_t *cf_data,
                           const int32_t *fn_indptr, const int32_t *fn_indices, HostPlan &P,
                           std::string &err, bool build_pos_maps = true, bool build_patterns = true) {
    const bool timing = getenv("POREB200_PLAN_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[plan] %-28s %8.1f ms\n", what,
                std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    if (nd != 2 && nd != 3) { err = "nd must be 2 or 3"; return 1; }
    if (nc <= 0 || nf <= 0 || nn <= 0) { err = "empty grid"; return 1; }
    P.nd = nd; P.nc = nc; P.nf = nf; P.nn = nn;
    P.fn_indptr.assign(fn_indptr, fn_indptr + nf + 1);
    const int64_t U = fn_indptr[nf];
    P.U = U;
    // ---- pass 1: sort faces inside each cell; count sub-half-faces per node
    std::vector<int32_t> cfaces(cf_indices, cf_indices + cf_indptr[nc]);
    std::vector<int8_t> csign(cf_data, cf_data + cf_indptr[nc]);
    int bad_input = 0;
#pragma omp parallel for schedule(static) reduction(| : bad_input) num_threads(plan_threads())
    for (int64_t c = 0; c < nc; ++c) {
        int b = cf_indptr[c], e = cf_indptr[c + 1];
        for (int i = b + 1; i < e; ++i) {  // insertion sort by face (cells have few faces)
            int32_t f = cfaces[i]; int8_t sg = csign[i]; int j = i - 1;
            while (j >= b && cfaces[j] > f) { cfaces[j + 1] = cfaces[j]; csign[j + 1] = csign[j]; --j; }
            cfaces[j + 1] = f; csign[j + 1] = sg;
        }
        for (int i = b; i < e; ++i) {
            if (cfaces[i] < 0 || cfaces[i] >= nf) bad_input |= 1;
            if (csign[i] != 1 && csign[i] != -1) bad_input |= 2;
        }
    }
    if (bad_input & 1) { err = "cell_faces index out of range"; return 1; }
    if (bad_input & 2) { err = "cell_faces data must be +-1"; return 1; }
    for (int64_t q = 0; q < U; ++q)
        if (fn_indices[q] < 0 || fn_indices[q] >= nn) { err = "face_nodes index out of range"; return 1; }
    lap("sort faces in cells");
    P.face_cells.assign(2 * nf, -1);
    for (int64_t c = 0; c < nc; ++c)
        for (int i = cf_indptr[c]; i < cf_indptr[c + 1]; ++i) {
            const int32_t f = cfaces[i];
            const int32_t enc = (int32_t)((c << 1) | (csign[i] < 0 ? 1 : 0));
            if (P.face_cells[2 * f] < 0) P.face_cells[2 * f] = enc;
            else P.face_cells[2 * f + 1] = enc;
        }
    std::vector<int64_t> hptr(nn + 1, 0);
    int64_t H = 0;
    for (int64_t c = 0; c < nc; ++c)
        for (int i = cf_indptr[c]; i < cf_indptr[c + 1]; ++i) {
            int32_t f = cfaces[i];
            for (int q = fn_indptr[f]; q < fn_indptr[f + 1]; ++q) { ++hptr[fn_indices[q] + 1]; ++H; }
        }
    P.H = H;
    if (H % nd != 0) { err = "cells must have exactly nd faces meeting in each vertex"; return 3; }
    for (int64_t s = 0; s < nn; ++s) hptr[s + 1] += hptr[s];
    // ---- pass 2: bucket by node; inside a node the order is (cell, face) ascending
    struct HF { int32_t c, f, u; int8_t sg; };
    std::vector<HF> hf(H);
    {
        std::vector<int64_t> fill(hptr.begin(), hptr.end() - 1);
        for (int64_t c = 0; c < nc; ++c)
            for (int i = cf_indptr[c]; i < cf_indptr[c + 1]; ++i) {
                int32_t f = cfaces[i];
                for (int q = fn_indptr[f]; q < fn_indptr[f + 1]; ++q)
                    hf[fill[fn_indices[q]]++] = HF{(int32_t)c, f, (int32_t)q, csign[i]};
            }
    }
    lap("bucket by node");
    // ---- sub-cells / sub-faces per node (offsets known up front -> nodes are independent)
    const int64_t S = H / nd;
    P.S = S;
    P.node_sc_ptr.assign(nn + 1, 0);
    P.node_sf_ptr.assign(nn + 1, 0);
    for (int64_t s = 0; s <= nn; ++s) {
        if (hptr[s] % nd != 0) { err = "cells must have exactly nd faces meeting in each vertex"; return 3; }
        P.node_sc_ptr[s] = (int32_t)(hptr[s] / nd);
    }
    for (int64_t q = 0; q < U; ++q) ++P.node_sf_ptr[fn_indices[q] + 1];
    for (int64_t s = 0; s < nn; ++s) P.node_sf_ptr[s + 1] += P.node_sf_ptr[s];
    P.sc_cell.resize(S);
    P.slot_sf.resize(H);
    P.sf_face.resize(U);
    P.sf_sides.assign(U, 0xFFFFFFFFu);
    P.sf_bloc.assign(U, 0xFFFF);
    P.node_nb.assign(nn, 0);
    P.sc_ncn.assign(nc, 0);
    int bad = 0, mx_sf = 0, mx_sc = 0, mx_nb = 0;
#pragma omp parallel num_threads(plan_threads())
    {
        std::vector<int32_t> us;
#pragma omp for schedule(dynamic, 512) reduction(| : bad) reduction(max : mx_sf, mx_sc, mx_nb)
        for (int64_t s = 0; s < nn; ++s) {
            const int64_t b = hptr[s], e = hptr[s + 1];
            const int64_t nh = e - b;
            const int nsc = (int)(nh / nd);
            const int64_t sc_fill = P.node_sc_ptr[s], sf_fill = P.node_sf_ptr[s];
            us.resize(nh);
            for (int64_t i = 0; i < nh; ++i) us[i] = hf[b + i].u;
            std::sort(us.begin(), us.end());
            us.erase(std::unique(us.begin(), us.end()), us.end());
            const int nsf = (int)us.size();
            if (nsf != P.node_sf_ptr[s + 1] - sf_fill) { bad |= 4; continue; }
            if (nsf > 32767 || nsc > 21000) { bad |= 8; continue; }
            int lbad = 0;
            for (int k = 0; k < nsc && !lbad; ++k) {
                const HF *h = &hf[b + (int64_t)k * nd];
                for (int m = 1; m < nd; ++m)
                    if (h[m].c != h[0].c) lbad = 3;
                if (k > 0 && h[0].c == hf[b + (int64_t)(k - 1) * nd].c) lbad = 3;
                if (lbad) break;
                P.sc_cell[sc_fill + k] = h[0].c;
                for (int m = 0; m < nd; ++m) {
                    int lu = (int)(std::lower_bound(us.begin(), us.end(), h[m].u) - us.begin());
                    P.slot_sf[(sc_fill + k) * nd + m] = (uint16_t)((lu << 1) | (h[m].sg < 0 ? 1 : 0));
                    uint32_t &sd = P.sf_sides[sf_fill + lu];
                    uint32_t side = (uint32_t)(k * nd + m);
                    if ((sd & 0xFFFFu) == 0xFFFFu) sd = (sd & 0xFFFF0000u) | side;
                    else if ((sd >> 16) == 0xFFFFu) sd = (sd & 0xFFFFu) | (side << 16);
                    else lbad = 16;
                }
            }
            if (lbad) { bad |= lbad; continue; }
            int nb = 0;
            for (int lu = 0; lu < nsf; ++lu) {
                int32_t u = us[lu];
                int32_t f = (int32_t)(std::upper_bound(fn_indptr, fn_indptr + nf + 1, u) - fn_indptr) - 1;
                P.sf_face[sf_fill + lu] = f;
                if ((P.sf_sides[sf_fill + lu] >> 16) == 0xFFFFu) P.sf_bloc[sf_fill + lu] = (uint16_t)nb++;
            }
            P.node_nb[s] = nb;
            mx_sf = std::max(mx_sf, nsf);
            mx_sc = std::max(mx_sc, nsc);
            mx_nb = std::max(mx_nb, nb);
        }
    }
    lap("per-node topology");
    if (bad & 3) { err = "cells must have exactly nd faces meeting in each vertex"; return 3; }
    if (bad & 4) { err = "face_nodes holds nodes without neighbouring cells"; return 1; }
    if (bad & 8) { err = "interaction region too large"; return 1; }
    if (bad & 16) { err = "face with more than two neighbouring cells"; return 1; }
    P.max_nsf = mx_sf; P.max_nsc = mx_sc; P.max_nb = mx_nb;
    for (int64_t q = 0; q < S; ++q) ++P.sc_ncn[P.sc_cell[q]];

    // ---- adjacency: face -> nodes is fn; cell -> nodes from the sub-cells
    std::vector<int32_t> cn_ptr(nc + 1, 0), cn_idx(S);
    for (int64_t c = 0; c < nc; ++c) cn_ptr[c + 1] = cn_ptr[c] + P.sc_ncn[c];
    {
        std::vector<int32_t> fill(cn_ptr.begin(), cn_ptr.end() - 1);
        for (int64_t s = 0; s < nn; ++s)
            for (int32_t q = P.node_sc_ptr[s]; q < P.node_sc_ptr[s + 1]; ++q)
                cn_idx[fill[P.sc_cell[q]]++] = (int32_t)s;
    }
    // boundary faces per node (faces of single-sided sub-faces), ascending by local order
    std::vector<int32_t> nbf_ptr(nn + 1, 0), nbf_idx;
    for (int64_t s = 0; s < nn; ++s) nbf_ptr[s + 1] = nbf_ptr[s] + P.node_nb[s];
    nbf_idx.resize(nbf_ptr[nn]);
#pragma omp parallel for schedule(static) num_threads(plan_threads())
    for (int64_t s = 0; s < nn; ++s)
        for (int32_t q = P.node_sf_ptr[s]; q < P.node_sf_ptr[s + 1]; ++q)
            if (P.sf_bloc[q] != 0xFFFF) nbf_idx[nbf_ptr[s] + P.sf_bloc[q]] = P.sf_face[q];

    lap("adjacency");
    // ---- patterns: union over the nodes of a row entity of the node's column entities
    auto build = [&](int64_t nrows, int64_t ncols, const int32_t *row_nodes_ptr,
                     const int32_t *row_nodes, const int32_t *col_ptr, const int32_t *col_idx,
                     Csr &out) {
        out.nrows = nrows; out.ncols = ncols;
        out.indptr.assign(nrows + 1, 0);
        for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel num_threads(plan_threads())
            {
                std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
                for (int64_t r = 0; r < nrows; ++r) {
                    tmp.clear();
                    for (int32_t q = row_nodes_ptr[r]; q < row_nodes_ptr[r + 1]; ++q) {
                        int32_t s = row_nodes[q];
                        tmp.insert(tmp.end(), col_idx + col_ptr[s], col_idx + col_ptr[s + 1]);
                    }
                    std::sort(tmp.begin(), tmp.end());
                    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                    if (pass == 0) out.indptr[r + 1] = (int32_t)tmp.size();
                    else std::copy(tmp.begin(), tmp.end(), out.indices.begin() + out.indptr[r]);
                }
            }
            if (pass == 0) {
                int64_t acc = 0;
                for (int64_t r = 0; r < nrows; ++r) {
                    acc += out.indptr[r + 1];
                    if (acc > 0x7FFFFFFFll) return 1;
                    out.indptr[r + 1] = (int32_t)acc;
                }
                out.indices.resize(acc);
            }
        }
        return 0;
    };
    P.nbf_ptr = nbf_ptr;
    P.nbf_idx = nbf_idx;
    P.cn_ptr = cn_ptr;
    P.cn_idx = cn_idx;
    int rc = 0;
    if (build_patterns) {
    rc |= build(nf, nc, fn_indptr, fn_indices, P.node_sc_ptr.data(), P.sc_cell.data(), P.pat[0]);
    rc |= build(nf, nf, fn_indptr, fn_indices, nbf_ptr.data(), nbf_idx.data(), P.pat[1]);
    rc |= build(nc, nc, cn_ptr.data(), cn_idx.data(), P.node_sc_ptr.data(), P.sc_cell.data(), P.pat[2]);
    rc |= build(nc, nf, cn_ptr.data(), cn_idx.data(), nbf_ptr.data(), nbf_idx.data(), P.pat[3]);
    }
    if (rc) { err = "pattern exceeds 2^31 entries; split the grid"; return 1; }
    lap("patterns");

    // ---- per-node position maps
    P.posfc_ptr.assign(nn + 1, 0); P.posfb_ptr.assign(nn + 1, 0);
    P.poscc_ptr.assign(nn + 1, 0); P.poscb_ptr.assign(nn + 1, 0);
    for (int64_t s = 0; s < nn; ++s) {
        int64_t nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s];
        int64_t nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        int64_t nb = P.node_nb[s];
        P.posfc_ptr[s + 1] = P.posfc_ptr[s] + nsf * nsc;
        P.posfb_ptr[s + 1] = P.posfb_ptr[s] + nsf * nb;
        P.poscc_ptr[s + 1] = P.poscc_ptr[s] + nsc * nsc;
        P.poscb_ptr[s + 1] = P.poscb_ptr[s] + nsc * nb;
    }
    if (!build_pos_maps) { lap("position map offsets"); return 0; }  // filled on the device (api.cu)
    P.pos_fc.resize(P.posfc_ptr[nn]); P.pos_fb.resize(P.posfb_ptr[nn]);
    P.pos_cc.resize(P.poscc_ptr[nn]); P.pos_cb.resize(P.poscb_ptr[nn]);
    auto find = [](const Csr &A, int32_t r, int32_t c) -> int32_t {
        const int32_t *b = A.indices.data() + A.indptr[r], *e = A.indices.data() + A.indptr[r + 1];
        return (int32_t)(std::lower_bound(b, e, c) - A.indices.data());
    };
#pragma omp parallel for schedule(dynamic, 512) num_threads(plan_threads())
    for (int64_t s = 0; s < nn; ++s) {
        const int32_t *cells = P.sc_cell.data() + P.node_sc_ptr[s];
        const int32_t *faces = P.sf_face.data() + P.node_sf_ptr[s];
        const int32_t *bfs = nbf_idx.data() + nbf_ptr[s];
        int nsc = P.node_sc_ptr[s + 1] - P.node_sc_ptr[s];
        int nsf = P.node_sf_ptr[s + 1] - P.node_sf_ptr[s];
        int nb = P.node_nb[s];
        for (int u = 0; u < nsf; ++u) {
            for (int k = 0; k < nsc; ++k)
                P.pos_fc[P.posfc_ptr[s] + (int64_t)u * nsc + k] = find(P.pat[0], faces[u], cells[k]);
            for (int b = 0; b < nb; ++b)
                P.pos_fb[P.posfb_ptr[s] + (int64_t)u * nb + b] = find(P.pat[1], faces[u], bfs[b]);
        }
        for (int k = 0; k < nsc; ++k) {
            for (int k2 = 0; k2 < nsc; ++k2)
                P.pos_cc[P.poscc_ptr[s] + (int64_t)k * nsc + k2] = find(P.pat[2], cells[k], cells[k2]);
            for (int b = 0; b < nb; ++b)
                P.pos_cb[P.poscb_ptr[s] + (int64_t)k * nb + b] = find(P.pat[3], cells[k], bfs[b]);
        }
    }
    lap("position maps");
    return 0;
}

}  // namespace pb
