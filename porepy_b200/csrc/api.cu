// api.cu -- C ABI of libporeb200.so (see include/poreb200.h), device memory management and
// kernel launches.  sm_100a only; there is no CPU path in this library: every compute entry
// point needs a CUDA device and fails with PB_ECUDA otherwise.
#include <cuda_runtime.h>

#include <atomic>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/poreb200.h"
#include "mpsa_node.cuh"
#include "node_kernels.cuh"
#include "plan_host.hpp"

using namespace pb;

// ------------------------------------------------------------------------------------
// error state
// ------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int64_t g_err_node = -1;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int pb_fail_(int code, const std::string &msg) { return fail(code, msg); }  // for spmv.cu
void pb_count_launch_() { g_launches++; }
#define CUDA_TRY(x)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess)                                                             \
            return fail(PB_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));        \
    } while (0)

extern "C" const char *pb_last_error(void) { return g_err.c_str(); }
extern "C" int64_t pb_last_error_node(void) { return g_err_node; }
extern "C" int64_t pb_launch_count(void) { return g_launches.load(); }
extern "C" int pb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
    return n;
}
extern "C" int pb_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return PB_OK;
}

// ------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes && p) return cudaSuccess;
        release();
        if (n == 0) n = 8;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    cudaError_t upload(const T *h, size_t count, cudaStream_t st) {
        cudaError_t e = ensure(count * sizeof(T));
        if (e != cudaSuccess) return e;
        if (count == 0) return cudaSuccess;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    template <class T>
    cudaError_t upload(const std::vector<T> &v, cudaStream_t st) { return upload(v.data(), v.size(), st); }
    template <class T>
    T *as() const { return (T *)p; }
};

struct NodeClass {
    int team = 32;
    int n = 0;
    int64_t smem_doubles = 0;  // per team
    DevBuf nodes;
};

struct pb_plan {
    HostPlan H;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // plan arrays
    DevBuf fn_indptr, node_sc_ptr, sc_cell, node_sf_ptr, sf_face, sf_sides, sf_bloc, slot_sf, node_nb,
        sc_ncn, posfc_ptr, posfb_ptr, poscc_ptr, poscb_ptr, pos_fc, pos_fb, pos_cc, pos_cb, fc_indptr,
        fb_indptr, cc_indptr, cb_indptr;
    // geometry
    DevBuf nodes, fnorm, fcent, farea, ccent, cvol;
    bool have_geo = false;
    PlanView view{};
    GeoView geo{};
    DevBuf err;
    // mpfa
    std::vector<NodeClass> mpfa_cls;
    DevBuf perm, bc, robw;
    bool have_robw = false;
    double eta = 0.0;
    bool mpfa_ready = false;
    DevBuf o_flux, o_bflux, o_bpc, o_bpf, o_vs, o_bpvs;
    // mpsa
    std::vector<NodeClass> mpsa_cls;
    int mpsa_cls_nalpha = -1;
    DevBuf stiff, vbc, vrobw, alpha;
    bool have_vrobw = false;
    int n_alpha = 0;
    double veta = 0.0;
    bool mpsa_ready = false;
    DevBuf o_stress, o_bstress, o_bdc, o_bdf;
    DevBuf o_dd[PB_MAX_ALPHA], o_bdd[PB_MAX_ALPHA], o_sg[PB_MAX_ALPHA], o_cons[PB_MAX_ALPHA],
        o_bdp[PB_MAX_ALPHA];
};

// ------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------
template <int ND, int TEAM>
__global__ void __launch_bounds__(TEAM == 32 ? 128 : TEAM)
    mpfa_kernel(PlanView P, GeoView G, MpfaParams prm, MpfaOut o, const int32_t *__restrict__ nodes,
                int n_nodes, int smem_doubles, int *err) {
    extern __shared__ double smem[];
    GpuTeam<TEAM> t;
    const int teams_per_block = blockDim.x / TEAM;
    const int team_in_block = threadIdx.x / TEAM;
    double *smd = smem + (size_t)team_in_block * smem_doubles;
    for (int i = blockIdx.x * teams_per_block + team_in_block; i < n_nodes;
         i += gridDim.x * teams_per_block)
        mpfa_node<ND>(t, P, G, prm, o, (int64_t)nodes[i], smd, err);
}

template <int ND, int TEAM>
__global__ void __launch_bounds__(TEAM == 32 ? 128 : TEAM)
    mpsa_kernel(PlanView P, GeoView G, MpsaParams prm, MpsaOut o, const int32_t *__restrict__ nodes,
                int n_nodes, int smem_doubles, int *err) {
    extern __shared__ double smem[];
    GpuTeam<TEAM> t;
    const int teams_per_block = blockDim.x / TEAM;
    const int team_in_block = threadIdx.x / TEAM;
    double *smd = smem + (size_t)team_in_block * smem_doubles;
    for (int i = blockIdx.x * teams_per_block + team_in_block; i < n_nodes;
         i += gridDim.x * teams_per_block)
        mpsa_node<ND>(t, P, G, prm, o, (int64_t)nodes[i], smd, err);
}

static const size_t kMaxSmem = 227 * 1024;
static const int kSMs = 148;

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
static int team_for(int n_unknowns) {
    if (n_unknowns <= 16) return 32;
    if (n_unknowns <= 40) return 64;
    if (n_unknowns <= 80) return 128;
    return 256;
}

template <class F>
static int build_classes(pb_plan *p, std::vector<NodeClass> &out, int unknowns_per_sf, F smem_of) {
    const HostPlan &H = p->H;
    const int teams[4] = {32, 64, 128, 256};
    std::vector<int32_t> lists[4];
    int64_t smem[4] = {0, 0, 0, 0};
    for (int64_t s = 0; s < H.nn; ++s) {
        int nsc = H.node_sc_ptr[s + 1] - H.node_sc_ptr[s];
        int nsf = H.node_sf_ptr[s + 1] - H.node_sf_ptr[s];
        if (nsf == 0) continue;
        int t = team_for(nsf * unknowns_per_sf);
        int ci = t == 32 ? 0 : t == 64 ? 1 : t == 128 ? 2 : 3;
        int64_t need = smem_of(nsf, nsc, H.node_nb[s]);
        // small teams share a CTA 4-ways: move up a class when 4 teams would not fit
        while (ci < 3 && need * 8 * (teams[ci] == 32 ? 4 : 1) > (int64_t)kMaxSmem) ++ci;
        lists[ci].push_back((int32_t)s);
        smem[ci] = std::max(smem[ci], need);
    }
    out.clear();
    for (int ci = 0; ci < 4; ++ci) {
        if (lists[ci].empty()) continue;
        out.emplace_back();
        NodeClass &c = out.back();
        c.team = teams[ci];
        c.n = (int)lists[ci].size();
        c.smem_doubles = smem[ci];
        if (c.nodes.upload(lists[ci], p->stream) != cudaSuccess) return fail(PB_ECUDA, "upload of node list failed");
    }
    return PB_OK;
}

extern "C" int pb_plan_create(int nd, int64_t nc, int64_t nf, int64_t nn, const int32_t *cf_indptr,
                              const int32_t *cf_indices, const int8_t *cf_data,
                              const int32_t *fn_indptr, const int32_t *fn_indices, pb_plan **out) {
    if (!out || !cf_indptr || !cf_indices || !cf_data || !fn_indptr || !fn_indices)
        return fail(PB_EINVAL, "null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PB_ECUDA, "no CUDA device: libporeb200 has no CPU path");
    pb_plan *p = new pb_plan;
    std::string err;
    int rc = build_host_plan(nd, nc, nf, nn, cf_indptr, cf_indices, cf_data, fn_indptr, fn_indices,
                             p->H, err);
    if (rc) {
        delete p;
        return fail(rc, err);
    }
    auto bail = [&](const char *what, cudaError_t e) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(e);
        delete p;
        return fail(PB_ECUDA, m);
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaEventCreate(&p->e0)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreate(&p->e1)) != cudaSuccess) return bail("event", e);
    HostPlan &H = p->H;
    cudaStream_t st = p->stream;
#define UP(field, vec)                                                      \
    if ((e = p->field.upload(vec, st)) != cudaSuccess) return bail(#field, e);
    UP(fn_indptr, H.fn_indptr) UP(node_sc_ptr, H.node_sc_ptr) UP(sc_cell, H.sc_cell)
    UP(node_sf_ptr, H.node_sf_ptr) UP(sf_face, H.sf_face) UP(sf_sides, H.sf_sides)
    UP(sf_bloc, H.sf_bloc) UP(slot_sf, H.slot_sf) UP(node_nb, H.node_nb) UP(sc_ncn, H.sc_ncn)
    UP(posfc_ptr, H.posfc_ptr) UP(posfb_ptr, H.posfb_ptr) UP(poscc_ptr, H.poscc_ptr)
    UP(poscb_ptr, H.poscb_ptr) UP(pos_fc, H.pos_fc) UP(pos_fb, H.pos_fb) UP(pos_cc, H.pos_cc)
    UP(pos_cb, H.pos_cb) UP(fc_indptr, H.pat[0].indptr) UP(fb_indptr, H.pat[1].indptr)
    UP(cc_indptr, H.pat[2].indptr) UP(cb_indptr, H.pat[3].indptr)
#undef UP
    if ((e = p->err.ensure(sizeof(int))) != cudaSuccess) return bail("err", e);
    PlanView &v = p->view;
    v.nd = nd; v.nc = nc; v.nf = nf; v.nn = nn;
    v.fn_indptr = p->fn_indptr.as<int32_t>();
    v.node_sc_ptr = p->node_sc_ptr.as<int32_t>(); v.sc_cell = p->sc_cell.as<int32_t>();
    v.node_sf_ptr = p->node_sf_ptr.as<int32_t>(); v.sf_face = p->sf_face.as<int32_t>();
    v.sf_sides = p->sf_sides.as<uint32_t>(); v.sf_bloc = p->sf_bloc.as<uint16_t>();
    v.slot_sf = p->slot_sf.as<uint16_t>(); v.node_nb = p->node_nb.as<int32_t>();
    v.sc_ncn = p->sc_ncn.as<int32_t>();
    v.posfc_ptr = p->posfc_ptr.as<int64_t>(); v.posfb_ptr = p->posfb_ptr.as<int64_t>();
    v.poscc_ptr = p->poscc_ptr.as<int64_t>(); v.poscb_ptr = p->poscb_ptr.as<int64_t>();
    v.pos_fc = p->pos_fc.as<int32_t>(); v.pos_fb = p->pos_fb.as<int32_t>();
    v.pos_cc = p->pos_cc.as<int32_t>(); v.pos_cb = p->pos_cb.as<int32_t>();
    v.fc_indptr = p->fc_indptr.as<int32_t>(); v.fb_indptr = p->fb_indptr.as<int32_t>();
    v.cc_indptr = p->cc_indptr.as<int32_t>(); v.cb_indptr = p->cb_indptr.as<int32_t>();
    rc = build_classes(p, p->mpfa_cls, 1,
                       [&](int nsf, int nsc, int nb) { return mpfa_smem_doubles(nd, nsf, nsc, nb); });
    if (rc) { delete p; return rc; }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail("sync", e);
    *out = p;
    return PB_OK;
}

extern "C" void pb_plan_destroy(pb_plan *p) {
    if (!p) return;
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->e0) cudaEventDestroy(p->e0);
    if (p->e1) cudaEventDestroy(p->e1);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" int pb_plan_sizes(const pb_plan *p, int64_t *num_subcells, int64_t *num_subfaces,
                             int64_t *num_subhalffaces, int32_t *max_sf, int32_t *max_sc) {
    if (!p) return fail(PB_EINVAL, "null plan");
    if (num_subcells) *num_subcells = p->H.S;
    if (num_subfaces) *num_subfaces = p->H.U;
    if (num_subhalffaces) *num_subhalffaces = p->H.H;
    if (max_sf) *max_sf = p->H.max_nsf;
    if (max_sc) *max_sc = p->H.max_nsc;
    return PB_OK;
}

extern "C" int pb_plan_pattern_size(const pb_plan *p, int which, int64_t *nrows, int64_t *nnz) {
    if (!p || which < 0 || which > 3) return fail(PB_EINVAL, "bad pattern id");
    *nrows = p->H.pat[which].nrows;
    *nnz = p->H.pat[which].nnz();
    return PB_OK;
}

extern "C" int pb_plan_pattern_get(const pb_plan *p, int which, int32_t *indptr, int32_t *indices) {
    if (!p || which < 0 || which > 3) return fail(PB_EINVAL, "bad pattern id");
    const Csr &c = p->H.pat[which];
    std::copy(c.indptr.begin(), c.indptr.end(), indptr);
    std::copy(c.indices.begin(), c.indices.end(), indices);
    return PB_OK;
}

extern "C" int pb_plan_set_geometry(pb_plan *p, const double *nodes, const double *face_normals,
                                    const double *face_centers, const double *face_areas,
                                    const double *cell_centers, const double *cell_volumes) {
    if (!p || !nodes || !face_normals || !face_centers || !face_areas || !cell_centers || !cell_volumes)
        return fail(PB_EINVAL, "null pointer");
    const HostPlan &H = p->H;
    cudaStream_t st = p->stream;
    CUDA_TRY(p->nodes.upload(nodes, 3 * H.nn, st));
    CUDA_TRY(p->fnorm.upload(face_normals, 3 * H.nf, st));
    CUDA_TRY(p->fcent.upload(face_centers, 3 * H.nf, st));
    CUDA_TRY(p->farea.upload(face_areas, H.nf, st));
    CUDA_TRY(p->ccent.upload(cell_centers, 3 * H.nc, st));
    CUDA_TRY(p->cvol.upload(cell_volumes, H.nc, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->geo = GeoView{p->nodes.as<double>(), p->fnorm.as<double>(), p->fcent.as<double>(),
                     p->farea.as<double>(), p->ccent.as<double>(), p->cvol.as<double>()};
    p->have_geo = true;
    return PB_OK;
}

// ------------------------------------------------------------------------------------
// error flag helpers
// ------------------------------------------------------------------------------------
static int check_singular(pb_plan *p) {
    int h = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(&h, p->err.p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (h != INT_MAX) {
        g_err_node = h;
        return fail(PB_ESINGULAR, "singular local system at node " + std::to_string(h));
    }
    return PB_OK;
}

// ------------------------------------------------------------------------------------
// MPFA
// ------------------------------------------------------------------------------------
extern "C" int pb_mpfa_upload(pb_plan *p, const double *perm, const uint8_t *bc,
                              const double *robin_weight, double eta) {
    if (!p || !perm || !bc) return fail(PB_EINVAL, "null pointer");
    if (!p->have_geo) return fail(PB_EINVAL, "pb_plan_set_geometry has not been called");
    const HostPlan &H = p->H;
    cudaStream_t st = p->stream;
    CUDA_TRY(p->perm.upload(perm, 9 * H.nc, st));
    CUDA_TRY(p->bc.upload(bc, H.nf, st));
    p->have_robw = robin_weight != nullptr;
    if (robin_weight) CUDA_TRY(p->robw.upload(robin_weight, H.nf, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->eta = eta;
    p->mpfa_ready = true;
    return PB_OK;
}

extern "C" int pb_mpfa_assemble(pb_plan *p, int want_flux, int want_trace, int want_vs, float *ms) {
    if (!p) return fail(PB_EINVAL, "null plan");
    if (!p->mpfa_ready) return fail(PB_EINVAL, "pb_mpfa_upload has not been called");
    const HostPlan &H = p->H;
    const int nd = H.nd;
    cudaStream_t st = p->stream;
    const size_t nfc = H.pat[0].nnz(), nfb = H.pat[1].nnz();
    MpfaOut o{};
    struct Req { DevBuf *b; double **slot; size_t n; bool want; };
    Req reqs[6] = {{&p->o_flux, &o.flux, nfc, want_flux != 0},
                   {&p->o_bflux, &o.bflux, nfb, want_flux != 0},
                   {&p->o_bpc, &o.bpc, nfc, want_trace != 0},
                   {&p->o_bpf, &o.bpf, nfb, want_trace != 0},
                   {&p->o_vs, &o.vs, nfc * nd, want_vs != 0 && want_flux != 0},
                   {&p->o_bpvs, &o.bpvs, nfc * nd, want_vs != 0 && want_trace != 0}};
    for (auto &r : reqs)
        if (r.want) CUDA_TRY(r.b->ensure(r.n * sizeof(double)));
    CUDA_TRY(cudaEventRecord(p->e0, st));
    for (auto &r : reqs) {
        if (!r.want) { *r.slot = nullptr; continue; }
        CUDA_TRY(cudaMemsetAsync(r.b->p, 0, r.n * sizeof(double), st));
        *r.slot = r.b->as<double>();
    }
    int init = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(p->err.p, &init, sizeof(int), cudaMemcpyHostToDevice, st));
    MpfaParams prm{p->perm.as<double>(), p->bc.as<uint8_t>(),
                   p->have_robw ? p->robw.as<double>() : nullptr, p->eta};
    for (const NodeClass &c : p->mpfa_cls) {
        int rc = PB_OK;
        int *err = p->err.as<int>();
        auto go = [&](auto kernel) {
            const int team = c.team;
            const int blk = team == 32 ? 128 : team;
            const int tpb = blk / team;
            const size_t smem = (size_t)c.smem_doubles * sizeof(double) * tpb;
            if (smem > kMaxSmem)
                return fail(PB_ENOTIMPL, "interaction region needs " + std::to_string(smem) +
                                             " B of shared memory (> 227 KB)");
            CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, blk, smem));
            if (per_sm < 1) per_sm = 1;
            int64_t need = ((int64_t)c.n + tpb - 1) / tpb;
            int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)kSMs * per_sm));
            kernel<<<grid, blk, smem, st>>>(p->view, p->geo, prm, o, c.nodes.as<int32_t>(), c.n,
                                            (int)c.smem_doubles, err);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return (int)PB_OK;
        };
        if (nd == 3) {
            switch (c.team) {
                case 32: rc = go(mpfa_kernel<3, 32>); break;
                case 64: rc = go(mpfa_kernel<3, 64>); break;
                case 128: rc = go(mpfa_kernel<3, 128>); break;
                default: rc = go(mpfa_kernel<3, 256>); break;
            }
        } else {
            switch (c.team) {
                case 32: rc = go(mpfa_kernel<2, 32>); break;
                case 64: rc = go(mpfa_kernel<2, 64>); break;
                case 128: rc = go(mpfa_kernel<2, 128>); break;
                default: rc = go(mpfa_kernel<2, 256>); break;
            }
        }
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(p->e1, st));
    CUDA_TRY(cudaEventSynchronize(p->e1));
    if (ms) CUDA_TRY(cudaEventElapsedTime(ms, p->e0, p->e1));
    return check_singular(p);
}

static int dl(pb_plan *p, DevBuf &b, double *h, size_t n) {
    if (!h) return PB_OK;
    if (!b.p || b.bytes < n * sizeof(double)) return fail(PB_EINVAL, "output was not assembled");
    CUDA_TRY(cudaMemcpyAsync(h, b.p, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    return PB_OK;
}

extern "C" int pb_mpfa_download(pb_plan *p, double *flux, double *bound_flux, double *bpc,
                                double *bpf, double *vs, double *bpvs) {
    if (!p) return fail(PB_EINVAL, "null plan");
    const HostPlan &H = p->H;
    const size_t nfc = H.pat[0].nnz(), nfb = H.pat[1].nnz();
    int rc;
    if ((rc = dl(p, p->o_flux, flux, nfc))) return rc;
    if ((rc = dl(p, p->o_bflux, bound_flux, nfb))) return rc;
    if ((rc = dl(p, p->o_bpc, bpc, nfc))) return rc;
    if ((rc = dl(p, p->o_bpf, bpf, nfb))) return rc;
    if ((rc = dl(p, p->o_vs, vs, nfc * H.nd))) return rc;
    if ((rc = dl(p, p->o_bpvs, bpvs, nfc * H.nd))) return rc;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PB_OK;
}

// ------------------------------------------------------------------------------------
// MPSA / Biot
// ------------------------------------------------------------------------------------
extern "C" int pb_mpsa_upload(pb_plan *p, const double *stiffness, const uint8_t *bc,
                              const double *robin_weight, double eta, int n_alpha,
                              const double *alpha) {
    if (!p || !stiffness || !bc) return fail(PB_EINVAL, "null pointer");
    if (!p->have_geo) return fail(PB_EINVAL, "pb_plan_set_geometry has not been called");
    if (n_alpha < 0 || n_alpha > PB_MAX_ALPHA) return fail(PB_EINVAL, "0 <= n_alpha <= 4");
    if (n_alpha > 0 && !alpha) return fail(PB_EINVAL, "null alpha");
    const HostPlan &H = p->H;
    const int nd = H.nd;
    cudaStream_t st = p->stream;
    CUDA_TRY(p->stiff.upload(stiffness, 81 * H.nc, st));
    CUDA_TRY(p->vbc.upload(bc, (size_t)nd * H.nf, st));
    p->have_vrobw = robin_weight != nullptr;
    if (robin_weight) CUDA_TRY(p->vrobw.upload(robin_weight, (size_t)nd * nd * H.nf, st));
    if (n_alpha) CUDA_TRY(p->alpha.upload(alpha, (size_t)n_alpha * 9 * H.nc, st));
    p->n_alpha = n_alpha;
    p->veta = eta;
    if (p->mpsa_cls_nalpha != n_alpha) {
        int rc = build_classes(p, p->mpsa_cls, nd, [&](int nsf, int nsc, int nb) {
            return mpsa_smem_doubles(nd, nsf, nsc, nb, n_alpha);
        });
        if (rc) return rc;
        p->mpsa_cls_nalpha = n_alpha;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    p->mpsa_ready = true;
    return PB_OK;
}

extern "C" int pb_mpsa_assemble(pb_plan *p, float *ms) {
    if (!p) return fail(PB_EINVAL, "null plan");
    if (!p->mpsa_ready) return fail(PB_EINVAL, "pb_mpsa_upload has not been called");
    const HostPlan &H = p->H;
    const int nd = H.nd;
    cudaStream_t st = p->stream;
    const size_t nfc = H.pat[0].nnz(), nfb = H.pat[1].nnz(), ncc = H.pat[2].nnz(), ncb = H.pat[3].nnz();
    MpsaOut o{};
    struct Req { DevBuf *b; double **slot; size_t n; };
    std::vector<Req> reqs = {{&p->o_stress, &o.stress, nfc * nd * nd},
                             {&p->o_bstress, &o.bstress, nfb * nd * nd},
                             {&p->o_bdc, &o.bdc, nfc * nd * nd},
                             {&p->o_bdf, &o.bdf, nfb * nd * nd}};
    for (int a = 0; a < p->n_alpha; ++a) {
        reqs.push_back({&p->o_dd[a], &o.dd[a], ncc * nd});
        reqs.push_back({&p->o_bdd[a], &o.bdd[a], ncb * nd});
        reqs.push_back({&p->o_sg[a], &o.sg[a], nfc * nd});
        reqs.push_back({&p->o_cons[a], &o.cons[a], ncc});
        reqs.push_back({&p->o_bdp[a], &o.bdp[a], nfc * nd});
    }
    for (auto &r : reqs) CUDA_TRY(r.b->ensure(r.n * sizeof(double)));
    CUDA_TRY(cudaEventRecord(p->e0, st));
    for (auto &r : reqs) {
        CUDA_TRY(cudaMemsetAsync(r.b->p, 0, r.n * sizeof(double), st));
        *r.slot = r.b->as<double>();
    }
    int init = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(p->err.p, &init, sizeof(int), cudaMemcpyHostToDevice, st));
    MpsaParams prm{p->stiff.as<double>(), p->vbc.as<uint8_t>(),
                   p->have_vrobw ? p->vrobw.as<double>() : nullptr, p->veta, p->n_alpha,
                   p->n_alpha ? p->alpha.as<double>() : nullptr};
    int *err = p->err.as<int>();
    for (const NodeClass &c : p->mpsa_cls) {
        auto go = [&](auto kernel) {
            const int team = c.team;
            const int blk = team == 32 ? 128 : team;
            const int tpb = blk / team;
            const size_t smem = (size_t)c.smem_doubles * sizeof(double) * tpb;
            if (smem > kMaxSmem)
                return fail(PB_ENOTIMPL, "interaction region needs " + std::to_string(smem) +
                                             " B of shared memory (> 227 KB)");
            CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, blk, smem));
            if (per_sm < 1) per_sm = 1;
            int64_t need = ((int64_t)c.n + tpb - 1) / tpb;
            int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)kSMs * per_sm));
            kernel<<<grid, blk, smem, st>>>(p->view, p->geo, prm, o, c.nodes.as<int32_t>(), c.n,
                                            (int)c.smem_doubles, err);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return (int)PB_OK;
        };
        int rc;
        if (nd == 3) {
            switch (c.team) {
                case 32: rc = go(mpsa_kernel<3, 32>); break;
                case 64: rc = go(mpsa_kernel<3, 64>); break;
                case 128: rc = go(mpsa_kernel<3, 128>); break;
                default: rc = go(mpsa_kernel<3, 256>); break;
            }
        } else {
            switch (c.team) {
                case 32: rc = go(mpsa_kernel<2, 32>); break;
                case 64: rc = go(mpsa_kernel<2, 64>); break;
                case 128: rc = go(mpsa_kernel<2, 128>); break;
                default: rc = go(mpsa_kernel<2, 256>); break;
            }
        }
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(p->e1, st));
    CUDA_TRY(cudaEventSynchronize(p->e1));
    if (ms) CUDA_TRY(cudaEventElapsedTime(ms, p->e0, p->e1));
    return check_singular(p);
}

extern "C" int pb_mpsa_download(pb_plan *p, double *stress, double *bound_stress, double *bdc,
                                double *bdf) {
    if (!p) return fail(PB_EINVAL, "null plan");
    const HostPlan &H = p->H;
    const size_t nd2 = (size_t)H.nd * H.nd;
    int rc;
    if ((rc = dl(p, p->o_stress, stress, H.pat[0].nnz() * nd2))) return rc;
    if ((rc = dl(p, p->o_bstress, bound_stress, H.pat[1].nnz() * nd2))) return rc;
    if ((rc = dl(p, p->o_bdc, bdc, H.pat[0].nnz() * nd2))) return rc;
    if ((rc = dl(p, p->o_bdf, bdf, H.pat[1].nnz() * nd2))) return rc;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PB_OK;
}

extern "C" int pb_biot_download(pb_plan *p, int a, double *dd, double *bdd, double *sg, double *cons,
                                double *bdp) {
    if (!p) return fail(PB_EINVAL, "null plan");
    if (a < 0 || a >= p->n_alpha) return fail(PB_EINVAL, "coupling tensor index out of range");
    const HostPlan &H = p->H;
    const size_t nd = H.nd;
    int rc;
    if ((rc = dl(p, p->o_dd[a], dd, H.pat[2].nnz() * nd))) return rc;
    if ((rc = dl(p, p->o_bdd[a], bdd, H.pat[3].nnz() * nd))) return rc;
    if ((rc = dl(p, p->o_sg[a], sg, H.pat[0].nnz() * nd))) return rc;
    if ((rc = dl(p, p->o_cons[a], cons, H.pat[2].nnz()))) return rc;
    if ((rc = dl(p, p->o_bdp[a], bdp, H.pat[0].nnz() * nd))) return rc;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PB_OK;
}
