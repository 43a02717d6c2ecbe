"""Where the host time of one cold end-to-end call goes (GPU box): cProfile of ``pb.Mpfa.discretize`` +
``pb.Mpsa.discretize`` on a bench workload, after one warm-up call.   python tools/e2e_profile.py tet1m"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "tet100k"
kind, dims, _ = bench.WORKLOADS[w]
g = bench.make_grid(kind, dims)
k, bc, C, vbc = bench.make_params(g)


def call():
    d1 = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    m1 = pb.Mpfa("flow")
    m1.discretize(g, d1)
    d2 = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc})
    m2 = pb.Mpsa("mech")
    m2.discretize(g, d2)
    return m1.last_timing, m2.last_timing


call()
if hasattr(g, "_b200_plan"):
    del g._b200_plan
os.environ["POREB200_PLAN_TIMING"] = "1"
pr = cProfile.Profile()
pr.enable()
t = call()
pr.disable()
print(t)
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
