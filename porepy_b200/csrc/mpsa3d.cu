// mpsa3d.cu -- instantiations of mpsa_kernel<3, Solver> (MPSA / Biot, 3-D grids) and their launcher.
#include "assembly_kernels.cuh"

int pb_launch_mpsa3_(pb_plan *p, const MpsaParams &prm, const MpsaOut &o) {
    for (const NodeClass &c : p->mpsa_cls) {
        int rc = PB_OK;
        PB_LAUNCH_CFG(mpsa_kernel, 3, prm, o)
        if (rc) return rc;
    }
    return PB_OK;
}
