// node_kernels.cuh -- per-interaction-region assembly routines (MPFA; MPSA/Biot in
// mpsa_node.cuh).  One TEAM of threads (a warp, or a CTA for large regions) owns one grid
// node: it gathers the sub-cell geometry and tensors, builds ONE small dense system in shared
// memory, solves it for all right-hand sides by Gauss-Jordan elimination with partial
// pivoting (FP64), and scatter-adds the sub-face rows into the face-indexed CSR value arrays.
//
// Formulation (DESIGN.md "Local systems"): the reference solves for the sub-cell gradients
// g_K (order nd*#subcells, numerics/fv/mpfa.py:926-1045).  Pressure continuity
// d_{K,f}.g_K + p_K = pbar_f holds exactly at the continuity points, and every sub-cell has
// exactly nd sub-faces at the node (_fvutils.py:735), so g_K = D_K^{-1}(pbar_{F_K} - p_K 1) and
// the only unknowns left are the continuity-point pressures pbar_f, one per sub-face: a system
// of order #subfaces (12 instead of 24 on interior Cartesian nodes, 8x fewer LU flops).  The
// flux-continuity / Neumann / Robin / Dirichlet rows and all right-hand sides are those of the
// reference (mpfa.py:997, 1080-1105, 1414-1578, 1274-1307) with g_K substituted.
//
// The routines are written against a small Team interface so that the identical source also
// compiles for the host with a 1-thread team; tests/emu uses that (test infrastructure only,
// never loaded by the product) to check the arithmetic against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>

#include "views.hpp"

namespace pb {

// ------------------------------------------------------------------------------------
// teams
// ------------------------------------------------------------------------------------
struct CpuTeam {
    PB_HD int tid() const { return 0; }
    PB_HD int size() const { return 1; }
    PB_HD int lanes() const { return 1; }
    PB_HD int warp() const { return 0; }
    PB_HD int nwarps() const { return 1; }
    PB_HD int lane() const { return 0; }
    PB_HD void sync() const {}
    PB_HD void warp_argmax(double &, int &) const {}
};

#if defined(__CUDACC__)
template <int TEAM>
struct GpuTeam {
    // TEAM == 32: several teams per CTA, one warp each (__syncwarp);
    // TEAM  > 32: the CTA is the team (__syncthreads)
    __device__ __forceinline__ int tid() const { return TEAM == 32 ? (threadIdx.x & 31) : threadIdx.x; }
    __device__ __forceinline__ int size() const { return TEAM; }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ int warp() const { return TEAM == 32 ? 0 : (threadIdx.x >> 5); }
    __device__ __forceinline__ int nwarps() const { return TEAM / 32; }
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ void sync() const {
        if (TEAM == 32) __syncwarp(); else __syncthreads();
    }
    // max of v over the 32 lanes of the calling warp, with its index; result in all lanes
    __device__ __forceinline__ void warp_argmax(double &v, int &i) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
    }
};
#endif

PB_HD void red_add(double *p, double v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);  // result unused -> RED.E.ADD.F64 to L2
#else
    *p += v;
#endif
}

PB_HD void flag_singular(int *err, int64_t node) {
#if defined(__CUDA_ARCH__)
    atomicMin(err, (int)node);
#else
    if ((int)node < *err) *err = (int)node;
#endif
}

// ------------------------------------------------------------------------------------
// Gauss-Jordan with partial pivoting on the augmented matrix A (n rows, row stride W, the
// first n columns are the system, the remaining nrhs the right-hand sides).  Rows are not
// swapped physically: rowidx[p] is the physical row holding pivot p.  On return
// X(p, c) = A[rowidx[p]*W + n + c] is the solution (already divided by the pivot).
// Replaces the dense np.linalg.inv per block of invert_diagonal_blocks
// (numerics/linalg/matrix_operations.py:1310-1371).
// ------------------------------------------------------------------------------------
template <class Team>
PB_HD bool gauss_jordan(Team &t, double *A, int n, int W, int nrhs, int *rowidx, double *ipiv) {
    bool ok = true;
    for (int p = 0; p < n; ++p) {
        if (t.warp() == 0) {
            double best = -1.0;
            int bi = p;
            for (int i = p + t.lane(); i < n; i += t.lanes()) {
                double v = fabs(A[rowidx[i] * W + p]);
                if (v > best) { best = v; bi = i; }
            }
            t.warp_argmax(best, bi);
            if (t.lane() == 0) {
                int r = rowidx[bi];
                rowidx[bi] = rowidx[p];
                rowidx[p] = r;
                // NaN-safe: !(best > 0) also catches NaN
                ipiv[p] = (best > 0.0) ? 1.0 / A[r * W + p] : 0.0;
            }
        }
        t.sync();
        const int pr = rowidx[p];
        const double inv = ipiv[p];
        if (inv == 0.0 || !(inv == inv)) ok = false;
        const int c0 = p + 1;
        const int wend = n + nrhs;
        for (int li = t.warp(); li < n; li += t.nwarps()) {
            if (li == p) continue;
            const int ri = rowidx[li];
            const double f = A[ri * W + p] * inv;
            if (f == 0.0) continue;
            for (int c = c0 + t.lane(); c < wend; c += t.lanes()) A[ri * W + c] -= f * A[pr * W + c];
        }
        t.sync();
    }
    // divide the right-hand sides by the pivots
    for (int p = t.warp(); p < n; p += t.nwarps()) {
        const int pr = rowidx[p];
        const double inv = ipiv[p];
        for (int c = n + t.lane(); c < n + nrhs; c += t.lanes()) A[pr * W + c] *= inv;
    }
    t.sync();
    return ok;
}

// ------------------------------------------------------------------------------------
// small dense inverse of the nd x nd matrix of distance vectors (rows d_m)
// ------------------------------------------------------------------------------------
template <int ND>
PB_HD bool invert_small(const double (&D)[ND][ND], double (&E)[ND][ND]);

template <>
PB_HD bool invert_small<2>(const double (&D)[2][2], double (&E)[2][2]) {
    double det = D[0][0] * D[1][1] - D[0][1] * D[1][0];
    if (det == 0.0 || !(det == det)) return false;
    double id = 1.0 / det;
    E[0][0] = D[1][1] * id; E[0][1] = -D[0][1] * id;
    E[1][0] = -D[1][0] * id; E[1][1] = D[0][0] * id;
    return true;
}

template <>
PB_HD bool invert_small<3>(const double (&D)[3][3], double (&E)[3][3]) {
    double c00 = D[1][1] * D[2][2] - D[1][2] * D[2][1];
    double c01 = D[1][2] * D[2][0] - D[1][0] * D[2][2];
    double c02 = D[1][0] * D[2][1] - D[1][1] * D[2][0];
    double det = D[0][0] * c00 + D[0][1] * c01 + D[0][2] * c02;
    if (det == 0.0 || !(det == det)) return false;
    double id = 1.0 / det;
    E[0][0] = c00 * id;
    E[1][0] = c01 * id;
    E[2][0] = c02 * id;
    E[0][1] = (D[0][2] * D[2][1] - D[0][1] * D[2][2]) * id;
    E[1][1] = (D[0][0] * D[2][2] - D[0][2] * D[2][0]) * id;
    E[2][1] = (D[0][1] * D[2][0] - D[0][0] * D[2][1]) * id;
    E[0][2] = (D[0][1] * D[1][2] - D[0][2] * D[1][1]) * id;
    E[1][2] = (D[0][2] * D[1][0] - D[0][0] * D[1][2]) * id;
    E[2][2] = (D[0][0] * D[1][1] - D[0][1] * D[1][0]) * id;
    return true;
}

// ------------------------------------------------------------------------------------
// MPFA
// ------------------------------------------------------------------------------------
// doubles of shared memory one team needs for a node with the given counts
PB_HD int64_t mpfa_smem_doubles(int nd, int nsf, int nsc, int nb) {
    int64_t W = (nsf + nsc + nb + nd * nsc) | 1;
    int64_t d = (int64_t)nsf * W + 2 * (int64_t)nsc * nd * nd + 3 * (int64_t)nsf;
    int64_t ints = nsc + 5 * (int64_t)nsf + (int64_t)nsc * nd;
    return d + (ints + 1) / 2 + 2;
}

template <int ND, class Team>
PB_HD void mpfa_node(Team &t, const PlanView &P, const GeoView &G, const MpfaParams &prm,
                     const MpfaOut &o, int64_t s, double *smd, int *err) {
    const int sc0 = P.node_sc_ptr[s], nsc = P.node_sc_ptr[s + 1] - sc0;
    const int sf0 = P.node_sf_ptr[s], nsf = P.node_sf_ptr[s + 1] - sf0;
    const int nb = P.node_nb[s];
    if (nsf == 0) return;
    const int nrhs = nsc + nb + ND * nsc;
    const int W = (nsf + nrhs) | 1;
    const int64_t nf = P.nf, nc = P.nc, nn = P.nn;
    double *A = smd;
    double *Tk = A + (int64_t)nsf * W;
    double *Rk = Tk + nsc * ND * ND;
    double *invmf = Rk + nsc * ND * ND;
    double *robw = invmf + nsf;
    double *ipiv = robw + nsf;
    int *cell = (int *)(ipiv + nsf);
    int *face = cell + nsc;
    int *sides = face + nsf;
    int *bloc = sides + nsf;
    int *bcu = bloc + nsf;
    int *rowidx = bcu + nsf;
    int *slot = rowidx + nsf;

    // ---- phase 1: stage the node's index lists, zero the system
    for (int k = t.tid(); k < nsc; k += t.size()) cell[k] = P.sc_cell[sc0 + k];
    for (int i = t.tid(); i < nsc * ND; i += t.size()) slot[i] = P.slot_sf[(int64_t)sc0 * ND + i];
    for (int u = t.tid(); u < nsf; u += t.size()) {
        const int f = P.sf_face[sf0 + u];
        face[u] = f;
        sides[u] = (int)P.sf_sides[sf0 + u];
        const int bl = P.sf_bloc[sf0 + u];
        bloc[u] = (bl == 0xFFFF) ? -1 : bl;
        const double im = 1.0 / (double)(P.fn_indptr[f + 1] - P.fn_indptr[f]);
        invmf[u] = im;
        int code = 0;
        if (bl != 0xFFFF) {
            code = prm.bc[f];
            if (code == 0) code = 2;  // boundary face without a flag: Neumann (params/bc.py:130-140)
        }
        bcu[u] = code;
        robw[u] = (code == 3 && prm.robw) ? prm.robw[f] * G.farea[f] * im : 0.0;
        rowidx[u] = u;
    }
    for (int i = t.tid(); i < nsf * W; i += t.size()) A[i] = 0.0;
    t.sync();

    // ---- phase 2: per sub-cell  D (distance rows), r = (n/m)^T K,  T = R D^{-1}
    for (int k = t.tid(); k < nsc; k += t.size()) {
        const int64_t c = cell[k];
        double xc[ND], xs[ND], K[ND][ND], D[ND][ND], R[ND][ND], E[ND][ND];
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            xc[i] = G.ccent[i * nc + c];
            xs[i] = G.nodes[i * nn + s];
#pragma unroll
            for (int j = 0; j < ND; ++j) K[i][j] = prm.perm[(i * 3 + j) * nc + c];
        }
#pragma unroll
        for (int m = 0; m < ND; ++m) {
            const int u = slot[k * ND + m] >> 1;
            const int64_t f = face[u];
            const double e = (bloc[u] >= 0) ? 0.0 : prm.eta;  // eta = 0 on boundary faces (_fvutils.py:259-263)
            double nrm[ND];
#pragma unroll
            for (int i = 0; i < ND; ++i) {
                const double xf = G.fcent[i * nf + f];
                D[m][i] = xf + e * (xs[i] - xf) - xc[i];
                nrm[i] = G.fnorm[i * nf + f] * invmf[u];
            }
#pragma unroll
            for (int j = 0; j < ND; ++j) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < ND; ++i) a += nrm[i] * K[i][j];
                R[m][j] = a;
            }
        }
        if (!invert_small<ND>(D, E)) flag_singular(err, s);
#pragma unroll
        for (int m = 0; m < ND; ++m)
#pragma unroll
            for (int m2 = 0; m2 < ND; ++m2) {
                double a = 0.0;
#pragma unroll
                for (int q = 0; q < ND; ++q) a += R[m][q] * E[q][m2];
                Tk[k * ND * ND + m * ND + m2] = a;
                Rk[k * ND * ND + m * ND + m2] = R[m][m2];
            }
    }
    t.sync();

    // ---- phase 3: one row per sub-face
    for (int u = t.tid(); u < nsf; u += t.size()) {
        double *row = A + (int64_t)u * W;
        const int code = bcu[u];
        if (code == 1) {  // Dirichlet: pbar_u = p_b   (mpfa.py:1547)
            row[u] = 1.0;
            row[nsf + nsc + bloc[u]] = 1.0;
            continue;
        }
        for (int sd = 0; sd < 2; ++sd) {
            const int side = sd == 0 ? (sides[u] & 0xFFFF) : ((sides[u] >> 16) & 0xFFFF);
            if (side == 0xFFFF) continue;
            const int k = side / ND, m = side - k * ND;
            const double sg = (slot[side] & 1) ? -1.0 : 1.0;
            double tau = 0.0;
#pragma unroll
            for (int m2 = 0; m2 < ND; ++m2) {
                const int u2 = slot[k * ND + m2] >> 1;
                const double tt = sg * Tk[k * ND * ND + m * ND + m2];
                row[u2] += tt;
                tau += tt;
            }
            row[nsf + k] += tau;
#pragma unroll
            for (int j = 0; j < ND; ++j) row[nsf + nsc + nb + k * ND + j] += sg * Rk[k * ND * ND + m * ND + j];
        }
        if (code == 2 || code == 3) row[nsf + nsc + bloc[u]] = -invmf[u];  // mpfa.py:1516-1526
        if (code == 3) row[u] -= robw[u];                                   // mpfa.py:869-887
        // row scaling (matrix_operations.py:1880-1906)
        double sum = 0.0;
        for (int c = 0; c < nsf; ++c) sum += fabs(row[c]);
        if (!(sum > 0.0)) { flag_singular(err, s); continue; }
        const double is = 1.0 / sum;
        for (int c = 0; c < nsf + nrhs; ++c) row[c] *= is;
    }
    t.sync();

    // ---- phase 4: solve for all right-hand sides
    if (!gauss_jordan(t, A, nsf, W, nrhs, rowidx, ipiv)) {
        if (t.tid() == 0) flag_singular(err, s);
        t.sync();
        return;
    }

    // ---- phase 5: sub-face rows -> CSR values.  Columns: [cells | boundary faces | (cell, j)]
    const int32_t *pfc = P.pos_fc + P.posfc_ptr[s];
    const int32_t *pfb = P.pos_fb + P.posfb_ptr[s];
    for (int u = t.warp(); u < nsf; u += t.nwarps()) {
        // The reference takes the flux from the side with the smaller cell index
        // (_fvutils.py:163).  Flux continuity makes both sides give the same number; in the
        // continuity-point formulation the side with the SMALLER transmissibilities is the
        // well-conditioned one (on the high-permeability side the gradient is a difference
        // of nearly equal pressures), so evaluate from that side.
        int side1 = sides[u] & 0xFFFF;
        {
            const int side2 = (sides[u] >> 16) & 0xFFFF;
            if (side2 != 0xFFFF) {
                double n1 = 0.0, n2 = 0.0;
#pragma unroll
                for (int m2 = 0; m2 < ND; ++m2) {
                    n1 += fabs(Tk[side1 * ND + m2]);
                    n2 += fabs(Tk[side2 * ND + m2]);
                }
                if (n2 < n1) side1 = side2;
            }
        }
        const int k1 = side1 / ND, m1 = side1 - k1 * ND;
        const double *T1 = Tk + k1 * ND * ND + m1 * ND;
        const double *R1 = Rk + k1 * ND * ND + m1 * ND;
        const double *xrow[ND];
        double tau1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < ND; ++m2) {
            xrow[m2] = A + (int64_t)rowidx[slot[k1 * ND + m2] >> 1] * W + nsf;
            tau1 += T1[m2];
        }
        const double *xu = A + (int64_t)rowidx[u] * W + nsf;
        const double im = invmf[u];
        for (int c = t.lane(); c < nrhs; c += t.lanes()) {
            double fl = 0.0;
#pragma unroll
            for (int m2 = 0; m2 < ND; ++m2) fl -= T1[m2] * xrow[m2][c];
            const double tr = xu[c] * im;
            if (c < nsc) {
                if (c == k1) fl += tau1;
                const int64_t p = pfc[u * nsc + c];
                if (o.flux) red_add(o.flux + p, fl);
                if (o.bpc) red_add(o.bpc + p, tr);
            } else if (c < nsc + nb) {
                const int64_t p = pfb[u * nb + (c - nsc)];
                if (o.bflux) red_add(o.bflux + p, fl);
                if (o.bpf) red_add(o.bpf + p, tr);
            } else {
                const int cc = c - nsc - nb;
                const int k = cc / ND, j = cc - k * ND;
                if (k == k1) fl += R1[j];
                const int64_t p = (int64_t)pfc[u * nsc + k] * ND + j;
                if (o.vs) red_add(o.vs + p, fl);
                if (o.bpvs) red_add(o.bpvs + p, tr);
            }
        }
    }
    t.sync();
}

}  // namespace pb
