// shard.cu -- one rank's share of a grid, extracted from the GLOBAL topology in one native pass (host code; the
// translation unit is a .cu only so that build.py treats all sources alike).
//
// Replaces the reference's memory-splitting scheme for this path: _fvutils.subproblems
// (numerics/fv/_fvutils.py:414-539: partition -> overlap by nodes -> extract_subgrid) and
// partition.extract_subgrid (grids/partition.py:540-640), with the rule of porepy_b200/shard.py:
//   * own cells = {part == rank}; own nodes = nodes of the faces of the own cells;
//   * shard cells = every cell with a face holding an own node (one halo layer), OWN CELLS FIRST (ascending
//     global id), then the halo cells (ascending);
//   * shard faces / nodes = faces of the shard cells / their nodes, ascending global id;
//   * kept face rows: faces of own cells, a face shared with another rank goes to the lower rank;
//   * cut faces: one cell inside the shard, two in the grid (artificial boundary of the overlap).
// All of it is masks and prefix sums over the CSC arrays of cell_faces (nf x nc) and face_nodes (nn x nf):
// O(size of the global arrays) with small constants, a few threads for the passes over all cells / faces.
#include <omp.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/poreb200.h"

int pb_fail_(int code, const std::string &msg);  // api.cu

struct pb_shard {
    int64_t nc = 0, nf = 0, nn = 0, n_own = 0;
    std::vector<int64_t> cells, faces, nodes;
    std::vector<uint8_t> own_face, cut_face, single_face, own_node;
    std::vector<int32_t> cf_indptr, cf_indices, fn_indptr, fn_indices;
    std::vector<double> cf_data;
};

static int host_threads() {
    int t = 4;
    if (const char *e = getenv("POREB200_HOST_THREADS")) t = atoi(e);
    return t < 1 ? 1 : (t > 64 ? 64 : t);
}

extern "C" void pb_shard_destroy(pb_shard *s) { delete s; }

extern "C" int pb_shard_create(int64_t nc, int64_t nf, int64_t nn, const int32_t *cf_indptr, const int32_t *cf_indices,
                               const double *cf_data, const int32_t *fn_indptr, const int32_t *fn_indices,
                               const int64_t *part, int64_t rank, pb_shard **out) {
    if (!out || !cf_indptr || !cf_indices || !cf_data || !fn_indptr || !fn_indices || !part || nc < 0 || nf < 0 || nn < 0)
        return pb_fail_(PB_EINVAL, "pb_shard_create: null pointer or negative size");
    const int nt = host_threads();
    pb_shard *s = new pb_shard;
    s->nc = nc; s->nf = nf; s->nn = nn;
    // ---- own cells -> faces of own cells -> own nodes
    std::vector<uint8_t> fmark(nf, 0);   // bit 0: face of an own cell, bit 1: face of the shard
    std::vector<uint8_t> nmark(nn, 0);   // bit 0: own node, bit 1: node of the shard
    for (int64_t c = 0; c < nc; ++c) {
        if (part[c] != rank) continue;
        for (int32_t q = cf_indptr[c]; q < cf_indptr[c + 1]; ++q) {
            const int32_t f = cf_indices[q];
            if (fmark[f] & 1) continue;
            fmark[f] |= 1;
            for (int32_t t = fn_indptr[f]; t < fn_indptr[f + 1]; ++t) nmark[fn_indices[t]] |= 1;
        }
    }
    // ---- cells touching an own node (through one of their faces)
    std::vector<uint8_t> ftouch(nf), ctouch(nc);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t f = 0; f < nf; ++f) {
        uint8_t t = 0;
        for (int32_t q = fn_indptr[f]; q < fn_indptr[f + 1]; ++q) t |= nmark[fn_indices[q]] & 1;
        ftouch[f] = t;
    }
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
        uint8_t t = 0;
        for (int32_t q = cf_indptr[c]; q < cf_indptr[c + 1]; ++q) t |= ftouch[cf_indices[q]];
        ctouch[c] = t;
    }
    int64_t n_own = 0, n_halo = 0;
    for (int64_t c = 0; c < nc; ++c) {
        if (!ctouch[c]) continue;
        if (part[c] == rank) ++n_own; else ++n_halo;
    }
    s->n_own = n_own;
    s->cells.resize(n_own + n_halo);
    {
        int64_t a = 0, b = n_own;
        for (int64_t c = 0; c < nc; ++c) {
            if (!ctouch[c]) continue;
            if (part[c] == rank) s->cells[a++] = c; else s->cells[b++] = c;
        }
    }
    const int64_t ncl = (int64_t)s->cells.size();
    // ---- faces / nodes of the shard, cell counts per face (local and global), lowest part per face
    std::vector<uint8_t> lcount(nf, 0), gcount(nf, 0);
    std::vector<int64_t> fminpart(0);
    for (int64_t q = 0, e = cf_indptr[nc]; q < e; ++q) {
        uint8_t &g = gcount[cf_indices[q]];
        if (g < 3) ++g;
    }
    std::vector<int32_t> fmin(nf, INT32_MAX);   // lowest part among the shard cells of a face (ranks fit 31 bits)
    int64_t nnz_cf = 0;
    for (int64_t i = 0; i < ncl; ++i) {
        const int64_t c = s->cells[i];
        const int32_t pc = (int32_t)part[c];
        for (int32_t q = cf_indptr[c]; q < cf_indptr[c + 1]; ++q) {
            const int32_t f = cf_indices[q];
            fmark[f] |= 2;
            if (lcount[f] < 3) ++lcount[f];
            if (pc < fmin[f]) fmin[f] = pc;
        }
        nnz_cf += cf_indptr[c + 1] - cf_indptr[c];
    }
    std::vector<int32_t> fmap(nf, -1), nmap(nn, -1);
    int64_t nfl = 0, nnz_fn = 0;
    for (int64_t f = 0; f < nf; ++f) {
        if (!(fmark[f] & 2)) continue;
        fmap[f] = (int32_t)nfl++;
        nnz_fn += fn_indptr[f + 1] - fn_indptr[f];
        for (int32_t t = fn_indptr[f]; t < fn_indptr[f + 1]; ++t) nmark[fn_indices[t]] |= 2;
    }
    int64_t nnl = 0;
    for (int64_t v = 0; v < nn; ++v)
        if (nmark[v] & 2) nmap[v] = (int32_t)nnl++;
    s->faces.resize(nfl);
    s->own_face.resize(nfl); s->cut_face.resize(nfl); s->single_face.resize(nfl);
    s->fn_indptr.assign(nfl + 1, 0);
    s->fn_indices.resize(nnz_fn);
    {
        int64_t w = 0;
        for (int64_t f = 0; f < nf; ++f) {
            const int32_t lf = fmap[f];
            if (lf < 0) continue;
            s->faces[lf] = f;
            const bool single = lcount[f] == 1;
            s->single_face[lf] = single;
            s->cut_face[lf] = single && gcount[f] != 1;
            s->own_face[lf] = (fmark[f] & 1) && fmin[f] == (int32_t)rank;
            for (int32_t t = fn_indptr[f]; t < fn_indptr[f + 1]; ++t) s->fn_indices[w++] = nmap[fn_indices[t]];
            s->fn_indptr[lf + 1] = (int32_t)w;
        }
    }
    s->nodes.resize(nnl);
    s->own_node.resize(nnl);
    for (int64_t v = 0; v < nn; ++v)
        if (nmap[v] >= 0) { s->nodes[nmap[v]] = v; s->own_node[nmap[v]] = nmark[v] & 1; }
    s->cf_indptr.assign(ncl + 1, 0);
    s->cf_indices.resize(nnz_cf);
    s->cf_data.resize(nnz_cf);
    {
        int64_t w = 0;
        for (int64_t i = 0; i < ncl; ++i) {
            const int64_t c = s->cells[i];
            for (int32_t q = cf_indptr[c]; q < cf_indptr[c + 1]; ++q) {
                s->cf_indices[w] = fmap[cf_indices[q]];
                s->cf_data[w++] = cf_data[q];
            }
            s->cf_indptr[i + 1] = (int32_t)w;
        }
    }
    *out = s;
    return PB_OK;
}

extern "C" int pb_shard_sizes(const pb_shard *s, int64_t *sizes) {
    if (!s || !sizes) return pb_fail_(PB_EINVAL, "pb_shard_sizes: null pointer");
    sizes[0] = (int64_t)s->cells.size(); sizes[1] = (int64_t)s->faces.size(); sizes[2] = (int64_t)s->nodes.size();
    sizes[3] = s->n_own; sizes[4] = (int64_t)s->cf_indices.size(); sizes[5] = (int64_t)s->fn_indices.size();
    return PB_OK;
}

template <class T>
static void copy_out(T *dst, const std::vector<T> &v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(T));
}

extern "C" int pb_shard_fill(const pb_shard *s, int64_t *cells, int64_t *faces, int64_t *nodes, uint8_t *own_face,
                             uint8_t *cut_face, uint8_t *single_face, uint8_t *own_node, int32_t *cf_indptr,
                             int32_t *cf_indices, double *cf_data, int32_t *fn_indptr, int32_t *fn_indices) {
    if (!s) return pb_fail_(PB_EINVAL, "pb_shard_fill: null shard");
    copy_out(cells, s->cells); copy_out(faces, s->faces); copy_out(nodes, s->nodes);
    copy_out(own_face, s->own_face); copy_out(cut_face, s->cut_face); copy_out(single_face, s->single_face);
    copy_out(own_node, s->own_node);
    copy_out(cf_indptr, s->cf_indptr); copy_out(cf_indices, s->cf_indices); copy_out(cf_data, s->cf_data);
    copy_out(fn_indptr, s->fn_indptr); copy_out(fn_indices, s->fn_indices);
    return PB_OK;
}

// dst[r, j] = src[r, idx[j]]  (the (3, n) geometry arrays of a sub-grid; row-major, nrows small)
extern "C" int pb_gather_columns(const double *src, int64_t nrows, int64_t ncols, const int64_t *idx, int64_t n,
                                 double *dst) {
    if (!src || !idx || !dst || nrows < 0 || ncols < 0 || n < 0) return pb_fail_(PB_EINVAL, "pb_gather_columns: bad arguments");
    const int nt = host_threads();
    for (int64_t r = 0; r < nrows; ++r) {
        const double *sr = src + r * ncols;
        double *dr = dst + r * n;
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t j = 0; j < n; ++j) dr[j] = sr[idx[j]];
    }
    return PB_OK;
}
