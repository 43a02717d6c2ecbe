// face.cu -- per-face discretizations (two-point flux approximation, first-order upwinding) on a light
// face-indexed grid handle: the face -> cell table and the three geometry arrays these schemes read.  Works
// for grids of any dimension (the reference delegates its 1-D MPFA / MPSA to TPFA, numerics/fv/mpfa.py:
// 690-712, mpsa.py:666-697); no interaction-region plan is built.
#include "plan.hpp"
#include "tpfa_diff.cuh"

struct pb_facegrid {
    int64_t nc = 0, nf = 0;
    cudaStream_t stream = nullptr;
    DevBuf face_cells, fnorm, fcent, ccent, tmp;
    GeoView geo{};
};

// one thread per cell: claim the first free slot of each of its faces
__global__ void face_cells_kernel(int64_t nc, const int32_t *__restrict__ cf_ip, const int32_t *__restrict__ cf_ix,
                                  const int8_t *__restrict__ cf_da, int32_t *__restrict__ fc, int *bad) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x)
        for (int q = cf_ip[c]; q < cf_ip[c + 1]; ++q) {
            const int32_t f = cf_ix[q];
            const int32_t enc = (int32_t)((c << 1) | (cf_da[q] < 0 ? 1 : 0));
            if (atomicCAS(fc + 2 * (int64_t)f, -1, enc) != -1)
                if (atomicCAS(fc + 2 * (int64_t)f + 1, -1, enc) != -1) atomicExch(bad, 1);
        }
}
// slot 0 = the smaller cell index (the order of the host construction in plan_host.hpp)
__global__ void face_cells_order_kernel(int64_t nf, int32_t *__restrict__ fc) {
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = fc[2 * f], b = fc[2 * f + 1];
        if (b >= 0 && b < a) { fc[2 * f] = b; fc[2 * f + 1] = a; }
    }
}

extern "C" int pb_facegrid_create(int64_t nc, int64_t nf, const int32_t *cf_indptr, const int32_t *cf_indices,
                                  const int8_t *cf_data, const double *face_normals, const double *face_centers,
                                  const double *cell_centers, pb_facegrid **out) {
    if (!out || !cf_indptr || !cf_indices || !cf_data || !face_normals || !face_centers || !cell_centers)
        return pb_fail_(PB_EINVAL, "null pointer");
    if (nc <= 0 || nf <= 0) return pb_fail_(PB_EINVAL, "empty grid");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return pb_fail_(PB_ECUDA, "no CUDA device: libporeb200 has no CPU path");
    for (int64_t q = 0; q < cf_indptr[nc]; ++q)
        if (cf_indices[q] < 0 || cf_indices[q] >= nf) return pb_fail_(PB_EINVAL, "cell_faces index out of range");
    pb_facegrid *g = new pb_facegrid;
    g->nc = nc; g->nf = nf;
    auto bail = [&](int rc) { pb_facegrid_destroy(g); return rc; };
#define FG_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return bail(pb_fail_(PB_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_))); } while (0)
    FG_TRY(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    cudaStream_t st = g->stream;
    DevBuf ip, ix, da, bad;
    FG_TRY(ip.upload(cf_indptr, (size_t)nc + 1, st));
    FG_TRY(ix.upload(cf_indices, (size_t)cf_indptr[nc], st));
    FG_TRY(da.upload(cf_data, (size_t)cf_indptr[nc], st));
    FG_TRY(bad.ensure(sizeof(int)));
    FG_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    FG_TRY(g->face_cells.ensure((size_t)2 * nf * sizeof(int32_t)));
    FG_TRY(cudaMemsetAsync(g->face_cells.p, 0xFF, (size_t)2 * nf * sizeof(int32_t), st));
    const int block = 256;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nc + block - 1) / block, (int64_t)kSMs * 16));
    face_cells_kernel<<<grid, block, 0, st>>>(nc, ip.as<int32_t>(), ix.as<int32_t>(), da.as<int8_t>(),
                                              g->face_cells.as<int32_t>(), bad.as<int>());
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((nf + block - 1) / block, (int64_t)kSMs * 16));
    face_cells_order_kernel<<<grid, block, 0, st>>>(nf, g->face_cells.as<int32_t>());
    pb_count_launch_(); pb_count_launch_();
    FG_TRY(cudaGetLastError());
    int rc;
    if ((rc = pb_upload_repacked_(st, g->tmp, g->fnorm, face_normals, 3, nf))) return bail(rc);
    if ((rc = pb_upload_repacked_(st, g->tmp, g->fcent, face_centers, 3, nf))) return bail(rc);
    if ((rc = pb_upload_repacked_(st, g->tmp, g->ccent, cell_centers, 3, nc))) return bail(rc);
    int hbad = 0;
    FG_TRY(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FG_TRY(cudaStreamSynchronize(st));
#undef FG_TRY
    if (hbad) return bail(pb_fail_(PB_EINVAL, "face with more than two neighbouring cells"));
    g->geo = GeoView{nullptr, g->fnorm.as<double>(), g->fcent.as<double>(), nullptr, g->ccent.as<double>(), nullptr,
                     1, 3, 1, 3, 1, 3};
    *out = g;
    return PB_OK;
}

extern "C" void pb_facegrid_destroy(pb_facegrid *g) {
    if (!g) return;
    if (g->stream) { cudaStreamSynchronize(g->stream); cudaStreamDestroy(g->stream); }
    delete g;
}

__global__ void tpfa_kernel(int64_t nf, GeoView G, const double *__restrict__ perm, int64_t perm_cs,
                            int64_t perm_es, const uint8_t *__restrict__ bc,
                            const int32_t *__restrict__ face_cells, const int32_t *__restrict__ fc_ptr,
                            int vdim, TpfaOut o) {
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x)
        tpfa_face(f, G, perm, perm_cs, perm_es, bc, face_cells, fc_ptr, vdim, o);
}

__global__ void upwind_kernel(int64_t nf, const double *__restrict__ q, const uint8_t *__restrict__ bc,
                              const int32_t *__restrict__ face_cells, int32_t *__restrict__ up_col,
                              double *__restrict__ neu_diag, double *__restrict__ dir_diag) {
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x)
        upwind_face(f, q, bc, face_cells, up_col, neu_diag, dir_diag);
}

extern "C" int pb_tpfa(pb_facegrid *g, const double *permeability, const uint8_t *bc_bits, const int32_t *fc_indptr,
                       int vdim, double *flux, double *bound_pressure_cell, double *vector_source,
                       double *bound_pressure_vector_source, double *bound_flux_diag,
                       double *bound_pressure_face_diag) {
    if (!g || !permeability || !bc_bits || !fc_indptr) return pb_fail_(PB_EINVAL, "null pointer");
    if (vdim < 1 || vdim > 3) return pb_fail_(PB_EINVAL, "1 <= vdim <= 3");
    cudaStream_t st = g->stream;
    const int64_t nf = g->nf, nc = g->nc;
    const size_t nnz = (size_t)fc_indptr[nf];
    DevBuf perm, bc, ip, o_flux, o_bpc, o_vs, o_bpvs, o_bf, o_bpf;
    { int rcs = pb_upload_repacked_(st, g->tmp, perm, permeability, 9, nc); if (rcs) return rcs; }
    CUDA_TRY(bc.upload(bc_bits, (size_t)nf, st));
    CUDA_TRY(ip.upload(fc_indptr, (size_t)nf + 1, st));
    CUDA_TRY(o_flux.ensure(nnz * sizeof(double)));
    CUDA_TRY(o_bpc.ensure(nnz * sizeof(double)));
    CUDA_TRY(o_vs.ensure(nnz * vdim * sizeof(double)));
    CUDA_TRY(o_bpvs.ensure(nnz * vdim * sizeof(double)));
    CUDA_TRY(o_bf.ensure((size_t)nf * sizeof(double)));
    CUDA_TRY(o_bpf.ensure((size_t)nf * sizeof(double)));
    TpfaOut o{o_flux.as<double>(), o_bpc.as<double>(), o_vs.as<double>(), o_bpvs.as<double>(),
              o_bf.as<double>(), o_bpf.as<double>()};
    const int block = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nf + block - 1) / block, (int64_t)kSMs * 16));
    tpfa_kernel<<<grid, block, 0, st>>>(nf, g->geo, perm.as<double>(), 1, 9, bc.as<uint8_t>(),
                                        g->face_cells.as<int32_t>(), ip.as<int32_t>(), vdim, o);
    pb_count_launch_();
    CUDA_TRY(cudaGetLastError());
    auto down = [&](double *h, DevBuf &d, size_t n) -> cudaError_t {
        return h ? cudaMemcpyAsync(h, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    CUDA_TRY(down(flux, o_flux, nnz));
    CUDA_TRY(down(bound_pressure_cell, o_bpc, nnz));
    CUDA_TRY(down(vector_source, o_vs, nnz * vdim));
    CUDA_TRY(down(bound_pressure_vector_source, o_bpvs, nnz * vdim));
    CUDA_TRY(down(bound_flux_diag, o_bf, (size_t)nf));
    CUDA_TRY(down(bound_pressure_face_diag, o_bpf, (size_t)nf));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PB_OK;
}

__global__ void tpfa_diff_kernel(int64_t nf, GeoView G, const double *__restrict__ k,
                                 const int32_t *__restrict__ face_cells, const int32_t *__restrict__ fc_ptr,
                                 double *__restrict__ t_hf, double *__restrict__ T, double *__restrict__ dT_dk) {
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x)
        tpfa_diff_face(f, G, k, face_cells, fc_ptr, t_hf, T, dT_dk);
}

// Differentiable TPFA (tpfa_diff.cuh): k = 9 * nc doubles, cell-major; fc_indptr as for pb_tpfa; outputs t_hf (nhf),
// T (nf), dT_dk (nhf * 9); host pointers.
extern "C" int pb_tpfa_diff(pb_facegrid *g, const double *k, const int32_t *fc_indptr, double *t_hf, double *T,
                            double *dT_dk) {
    if (!g || !k || !fc_indptr || !t_hf || !T || !dT_dk) return pb_fail_(PB_EINVAL, "null pointer");
    cudaStream_t st = g->stream;
    const int64_t nf = g->nf, nc = g->nc;
    const size_t nhf = (size_t)fc_indptr[nf];
    DevBuf dk, ip, o_t, o_T, o_d;
    CUDA_TRY(dk.upload(k, (size_t)9 * nc, st));
    CUDA_TRY(ip.upload(fc_indptr, (size_t)nf + 1, st));
    CUDA_TRY(o_t.ensure(nhf * sizeof(double)));
    CUDA_TRY(o_T.ensure((size_t)nf * sizeof(double)));
    CUDA_TRY(o_d.ensure(nhf * 9 * sizeof(double)));
    const int block = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nf + block - 1) / block, (int64_t)kSMs * 16));
    tpfa_diff_kernel<<<grid, block, 0, st>>>(nf, g->geo, dk.as<double>(), g->face_cells.as<int32_t>(), ip.as<int32_t>(),
                                             o_t.as<double>(), o_T.as<double>(), o_d.as<double>());
    pb_count_launch_();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(t_hf, o_t.p, nhf * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(T, o_T.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(dT_dk, o_d.p, nhf * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_upwind(pb_facegrid *g, const double *darcy_flux, const uint8_t *bc_bits, int32_t *upstream_cell,
                         double *neumann_diag, double *dirichlet_diag) {
    if (!g || !darcy_flux || !bc_bits || !upstream_cell || !neumann_diag || !dirichlet_diag)
        return pb_fail_(PB_EINVAL, "null pointer");
    cudaStream_t st = g->stream;
    const int64_t nf = g->nf;
    DevBuf q, bc, up, neu, dir;
    CUDA_TRY(q.upload(darcy_flux, (size_t)nf, st));
    CUDA_TRY(bc.upload(bc_bits, (size_t)nf, st));
    CUDA_TRY(up.ensure((size_t)nf * sizeof(int32_t)));
    CUDA_TRY(neu.ensure((size_t)nf * sizeof(double)));
    CUDA_TRY(dir.ensure((size_t)nf * sizeof(double)));
    const int block = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nf + block - 1) / block, (int64_t)kSMs * 16));
    upwind_kernel<<<grid, block, 0, st>>>(nf, q.as<double>(), bc.as<uint8_t>(), g->face_cells.as<int32_t>(),
                                          up.as<int32_t>(), neu.as<double>(), dir.as<double>());
    pb_count_launch_();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(upstream_cell, up.p, (size_t)nf * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(neumann_diag, neu.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(dirichlet_diag, dir.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PB_OK;
}

// ---- interface upwinding (UpwindCoupling.discretize, numerics/fv/upwind.py:427-528): per mortar cell the sign of
// the interface flux and the two upstream masks (flux > 0: the higher-dimensional side is upstream)
__global__ void upwind_coupling_kernel(int64_t n, const double *__restrict__ lam, double *__restrict__ sgn,
                                       double *__restrict__ from_primary, double *__restrict__ from_secondary) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = lam[i];
        const double s = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v));   // np.sign (nan stays nan)
        sgn[i] = s;
        from_primary[i] = s > 0.0 ? 1.0 : 0.0;
        from_secondary[i] = s > 0.0 ? 0.0 : 1.0;
    }
}

extern "C" int pb_upwind_coupling(int64_t n, const double *interface_flux, double *sign, double *from_primary,
                                  double *from_secondary) {
    if (n < 0 || (n > 0 && (!interface_flux || !sign || !from_primary || !from_secondary)))
        return pb_fail_(PB_EINVAL, "bad arguments");
    if (n == 0) return PB_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return pb_fail_(PB_ECUDA, "no CUDA device: libporeb200 has no CPU path");
    DevBuf in, o0, o1, o2;
    CUDA_TRY(in.upload(interface_flux, (size_t)n, 0));
    CUDA_TRY(o0.ensure((size_t)n * 8)); CUDA_TRY(o1.ensure((size_t)n * 8)); CUDA_TRY(o2.ensure((size_t)n * 8));
    const int block = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t)kSMs * 16));
    upwind_coupling_kernel<<<grid, block>>>(n, in.as<double>(), o0.as<double>(), o1.as<double>(), o2.as<double>());
    pb_count_launch_();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(sign, o0.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(from_primary, o1.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(from_secondary, o2.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return PB_OK;
}
