"""Where the "topology_plan" stage of bench.py's e2e call goes on the host (one GPU): destruction of the previous plan,
fingerprint, pb_plan_create, geometry upload -- with the grid arrays page-locked as in the bench."""
import gc
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import fv  # noqa: E402

g = bench.make_grid("tet", (55, 55, 55))
bench.pin_grid(g) if hasattr(bench, "pin_grid") else None
k, bc, C, vbc = bench.make_params(g)
for i in range(6):
    t0 = time.perf_counter()
    if hasattr(g, "_b200_plan"):
        del g._b200_plan
    gc.collect()
    t1 = time.perf_counter()
    fp = fv.DevicePlan._fingerprint(g, g.cell_faces, g.face_nodes)
    t2 = time.perf_counter()
    plan = fv.DevicePlan(g)
    t3 = time.perf_counter()
    arrs, rot = fv.plan_geometry(g)
    t4 = time.perf_counter()
    plan.set_geometry(g)
    t5 = time.perf_counter()
    g._b200_plan = plan
    # a discretization in between, as in the bench (its outputs die with `data`)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    pb.Mpfa("flow").discretize(g, data)
    t6 = time.perf_counter()
    del data
    print(f"call {i}: destroy+gc {t1 - t0:.4f}  fingerprint {t2 - t1:.4f}  DevicePlan() {t3 - t2:.4f} (pb_plan_create "
          f"{plan.plan_seconds:.4f})  plan_geometry {t4 - t3:.4f}  set_geometry {t5 - t4:.4f}  mpfa discretize {t6 - t5:.4f}", flush=True)
