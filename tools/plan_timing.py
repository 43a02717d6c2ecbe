"""Wall time of the topology plan and its phases (needs a GPU):   POREB200_PLAN_TIMING=1 python tools/plan_timing.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402

g = bench.make_grid('tet', (55, 55, 55))
for i in range(4):
    if hasattr(g, '_b200_plan'):
        del g._b200_plan
    t = time.time()
    p = pb.DevicePlan.for_grid(g)
    print('DevicePlan.for_grid', round(time.time() - t, 4), 'plan_seconds', round(p.plan_seconds, 4), flush=True)
