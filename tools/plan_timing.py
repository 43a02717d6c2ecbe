import sys, time
sys.path.insert(0, '.')
import bench, porepy_b200 as pb
g = bench.make_grid('tet', (55,55,55))
for i in range(2):
    if hasattr(g, '_b200_plan'): del g._b200_plan
    t=time.time(); p = pb.DevicePlan.for_grid(g); print('DevicePlan.for_grid', round(time.time()-t,3))
