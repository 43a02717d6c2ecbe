// mpfa_launch.cu -- instantiations of mpfa_kernel<ND, Solver> (2-D and 3-D) and their launcher.
#include "assembly_kernels.cuh"

int pb_launch_mpfa_(pb_plan *p, const MpfaParams &prm, const MpfaOut &o) {
    const int nd = p->H.nd;
    for (const NodeClass &c : p->mpfa_cls) {
        int rc = PB_OK;
        if (nd == 3) { PB_LAUNCH_CFG(mpfa_kernel, 3, prm, o) } else { PB_LAUNCH_CFG(mpfa_kernel, 2, prm, o) }
        if (rc) return rc;
    }
    return PB_OK;
}
