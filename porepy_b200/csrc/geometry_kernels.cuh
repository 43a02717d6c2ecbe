// geometry_kernels.cuh -- Grid.compute_geometry for 3-D grids of general polyhedral cells, one call per face and one
// per cell (reference grids/grid.py:572-778, _compute_geometry_3d, itself after MRST).
//
//   geom_face   a face with ordered nodes x_0 .. x_{m-1} is fanned into m triangles (x_i, x_{i+1}, xbar), xbar the mean
//               of the nodes: normal = sum of the triangles' area vectors, area = sum of their areas, centre = area-
//               weighted mean of their centroids (grid.py:590-668).
//   geom_cell   a cell is the union of the tetrahedra (triangle, cbar), cbar = edge-count-weighted mean of its faces'
//               centres: volume = sum of the signed tetrahedron volumes (outward triangle normal . (centroid - cbar) / 3),
//               centre = cbar + volume-weighted mean of 3/4 (centroid - cbar)  (grid.py:672-775).
// The sums run in the reference's order (edges of a face in node order; faces of a cell in ascending index = the
// sorted CSC column), so the results agree with the reference to rounding of the individual products.
// The CUDA kernels run one thread per face / per cell; the test-only host build loops.
#pragma once
#include <cmath>

#include "views.hpp"

namespace pb {

struct GeomOut {
    double *fnorm, *fcent, *farea, *ccent, *cvol;
    int64_t face_cs, face_es, cell_cs, cell_es;   // vector component i of entity e at [i * cs + e * es]
};

struct FaceFan {     // what a cell needs of a face it looks at: recomputed per (cell, face), nothing is stored per edge
    double xbar[3], normal[3];
};

PB_HD void geom_face_fan(int64_t f, const int32_t *fn_indptr, const int32_t *fn_indices, const double *nodes,
                         int64_t node_cs, int64_t node_es, FaceFan &fan, double &area, double (&centre)[3]) {
    const int b = fn_indptr[f], e = fn_indptr[f + 1], m = e - b;
    double xb[3] = {0.0, 0.0, 0.0};
    for (int q = b; q < e; ++q)
        for (int i = 0; i < 3; ++i) xb[i] += nodes[i * node_cs + (int64_t)fn_indices[q] * node_es];
    for (int i = 0; i < 3; ++i) xb[i] /= (double)m;
    double nsum[3] = {0.0, 0.0, 0.0}, asum = 0.0, csum[3] = {0.0, 0.0, 0.0};
    for (int q = b; q < e; ++q) {
        const int64_t v0 = fn_indices[q], v1 = fn_indices[q + 1 < e ? q + 1 : b];
        double x0[3], x1[3], along[3], f2n[3];
        for (int i = 0; i < 3; ++i) {
            x0[i] = nodes[i * node_cs + v0 * node_es];
            x1[i] = nodes[i * node_cs + v1 * node_es];
            along[i] = x1[i] - x0[i];
            f2n[i] = xb[i] - x0[i];
        }
        const double sn[3] = {(along[1] * f2n[2] - along[2] * f2n[1]) / 2, (along[2] * f2n[0] - along[0] * f2n[2]) / 2,
                              (along[0] * f2n[1] - along[1] * f2n[0]) / 2};
        const double sa = sqrt(sn[0] * sn[0] + sn[1] * sn[1] + sn[2] * sn[2]);
        asum += sa;
        for (int i = 0; i < 3; ++i) {
            nsum[i] += sn[i];
            csum[i] += sa * ((x0[i] + x1[i] + xb[i]) / 3);
        }
    }
    area = asum;
    for (int i = 0; i < 3; ++i) {
        fan.xbar[i] = xb[i];
        fan.normal[i] = nsum[i];
        centre[i] = csum[i] / asum;
    }
}

PB_HD void geom_face(int64_t f, const int32_t *fn_indptr, const int32_t *fn_indices, const double *nodes,
                     int64_t node_cs, int64_t node_es, const GeomOut &o) {
    FaceFan fan;
    double area, centre[3];
    geom_face_fan(f, fn_indptr, fn_indices, nodes, node_cs, node_es, fan, area, centre);
    o.farea[f] = area;
    for (int i = 0; i < 3; ++i) {
        o.fnorm[i * o.face_cs + f * o.face_es] = fan.normal[i];
        o.fcent[i * o.face_cs + f * o.face_es] = centre[i];
    }
}

// after geom_face of all faces (reads the face centres and normals back)
PB_HD bool geom_cell(int64_t c, const int32_t *cf_indptr, const int32_t *cf_indices, const int8_t *cf_sign,
                     const int32_t *fn_indptr, const int32_t *fn_indices, const double *nodes, int64_t node_cs,
                     int64_t node_es, const GeomOut &o) {
    const int b = cf_indptr[c], e = cf_indptr[c + 1];
    // temporary centre: every edge brings in its face's centre (grid.py:723-726)
    int nedges = 0;
    for (int q = b; q < e; ++q) nedges += fn_indptr[cf_indices[q] + 1] - fn_indptr[cf_indices[q]];
    double cb[3] = {0.0, 0.0, 0.0};
    for (int q = b; q < e; ++q) {
        const int64_t f = cf_indices[q];
        const int m = fn_indptr[f + 1] - fn_indptr[f];
        for (int i = 0; i < 3; ++i) {
            const double term = o.fcent[i * o.face_cs + f * o.face_es] / (double)nedges;
            for (int r = 0; r < m; ++r) cb[i] += term;     // one addend per edge, as np.bincount sums them
        }
    }
    double vol = 0.0, rel[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    for (int q = b; q < e; ++q) {
        const int64_t f = cf_indices[q];
        const double orient = (double)cf_sign[q];
        const int fb = fn_indptr[f], fe = fn_indptr[f + 1], m = fe - fb;
        double xb[3] = {0.0, 0.0, 0.0}, fnrm[3];
        for (int t = fb; t < fe; ++t)
            for (int i = 0; i < 3; ++i) xb[i] += nodes[i * node_cs + (int64_t)fn_indices[t] * node_es];
        for (int i = 0; i < 3; ++i) {
            xb[i] /= (double)m;
            fnrm[i] = o.fnorm[i * o.face_cs + f * o.face_es];
        }
        for (int t = fb; t < fe; ++t) {
            const int64_t v0 = fn_indices[t], v1 = fn_indices[t + 1 < fe ? t + 1 : fb];
            double x0[3], x1[3], along[3], f2n[3], dist[3];
            for (int i = 0; i < 3; ++i) {
                x0[i] = nodes[i * node_cs + v0 * node_es];
                x1[i] = nodes[i * node_cs + v1 * node_es];
                along[i] = x1[i] - x0[i];
                f2n[i] = xb[i] - x0[i];
                dist[i] = (x0[i] + x1[i] + xb[i]) / 3 - cb[i];
            }
            const double sn[3] = {(along[1] * f2n[2] - along[2] * f2n[1]) / 2, (along[2] * f2n[0] - along[0] * f2n[2]) / 2,
                                  (along[0] * f2n[1] - along[1] * f2n[0]) / 2};
            const double dot = sn[0] * fnrm[0] + sn[1] * fnrm[1] + sn[2] * fnrm[2];
            const double sgn = dot > 0.0 ? 1.0 : (dot < 0.0 ? -1.0 : 0.0);     // np.sign (grid.py:649-654)
            double tv = 0.0;
            for (int i = 0; i < 3; ++i) tv += dist[i] * (sn[i] * orient * sgn);
            tv /= 3;
            if (!(tv > -1e-12)) ok = false;     // "Some tetrahedra have negative volume" (grid.py:754-755)
            vol += tv;
            for (int i = 0; i < 3; ++i) rel[i] += tv * (3.0 / 4 * dist[i]);
        }
    }
    o.cvol[c] = vol;
    for (int i = 0; i < 3; ++i) o.ccent[i * o.cell_cs + c * o.cell_es] = cb[i] + rel[i] / vol;
    return ok;
}

}  // namespace pb
