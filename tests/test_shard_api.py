"""``porepy_b200.shard.discretize_shard`` -- the one-call-per-rank API of the multi-GPU path -- driven
on CPU with the device plan replaced by the host build of the node routines (tests/emu): the sum of
the ranks' rows equals the unsplit discretization (applications/test_utils/common_xpfa_tests.py:
832-957), for MPFA and for Biot with a scalar and a tensor coupling, incl. a 2-rank gloo run."""
import os

import numpy as np
import pytest

import porepy_b200 as pb
from porepy_b200 import fv, shard as sh
from cases import flatten
from emu_binding import EmuBackedPlan
from golden_io import rel_err


def problem(kind):
    g = pb.structured_tet_grid([3, 2, 2]) if kind == "tet" else pb.cart_grid_3d([6, 3, 3], perturb=0.3)
    rng = np.random.default_rng(3)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    lab = np.where((x < 1e-10) | (x > 1 - 1e-10), "dir", "neu")
    lab[g.face_centers[2, bf] > 1 - 1e-10] = "rob"
    bc = pb.BoundaryCondition(g, bf, list(lab))
    bc.robin_weight = 0.5 + rng.random(g.num_faces)
    C = pb.FourthOrderTensor(np.exp(0.3 * rng.standard_normal(nc)), np.exp(0.3 * rng.standard_normal(nc)))
    vbc = pb.BoundaryConditionVectorial(g, bf[g.face_centers[2, bf] < 1e-10], "dir")
    flow = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    mech = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc,
                                           "scalar_vector_mappings": {"p": 0.8, "T": k}})
    return g, flow, mech


def unsplit(g, flow, mech):
    pb.Mpfa("flow").discretize(g, flow)
    pb.Biot("mech").discretize(g, mech)
    ref = dict(flow[pb.DISCRETIZATION_MATRICES]["flow"])
    ref.update(mech[pb.DISCRETIZATION_MATRICES]["mech"])
    return flatten(ref)


@pytest.mark.parametrize("kind", ["cart", "tet"])
@pytest.mark.parametrize("nparts", [2, 3])
def test_sum_of_rank_rows_equals_unsplit(kind, nparts, monkeypatch):
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    g, flow, mech = problem(kind)
    part = sh.partition_cells(g, nparts)
    per_rank = []
    for r in range(nparts):
        out = sh.discretize_shard(pb.Mpfa("flow"), g, flow, part, r)
        out.update(sh.discretize_shard(pb.Biot("mech"), g, mech, part, r))
        per_rank.append(out)
    got = flatten(sh.sum_shards(per_rank))
    ref = unsplit(g, flow, mech)
    assert set(got) == set(ref)
    for key in ref:
        assert got[key].shape == ref[key].shape
        assert rel_err(ref[key], got[key]) < 1e-12, key
    # a rank's result holds rows of its own faces only
    s0 = sh.extract_shard(g, part, 0)
    rows = np.unique(per_rank[0]["flux"].nonzero()[0])
    assert np.all(np.isin(rows, s0.faces[s0.own_face]))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fv.DevicePlan = EmuBackedPlan          # CPU stand-in for the kernels (this process only)
    g, flow, mech = problem("cart")
    part = sh.partition_cells(g, world)
    out = sh.discretize_shard(pb.Mpfa("flow"), g, flow, part, rank)
    out.update(sh.discretize_shard(pb.Biot("mech"), g, mech, part, rank))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    if rank == 0:
        got = flatten(sh.sum_shards(gathered))
        ref = unsplit(g, flow, mech)
        q.put(max(rel_err(ref[key], got[key]) for key in ref))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(60)
    assert err < 1e-12
