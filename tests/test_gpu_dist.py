"""Row-sharded multi-GPU solve (cuOptB200SolveDistributed) against the single-GPU solve of the same LP.
Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, size, tol, mode):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cuopt_b200 import capi, lpgen
        from cuopt_b200 import dist as cdist
        lp = lpgen.sparse_lp(size, size, 8, seed=21)
        comm = cdist.bootstrap(rank, world, device=torch.device("cuda", rank))
        p, (r0, r1) = cdist.local_problem(lp, rank, world)
        s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=mode)
        s.set("optimality_tolerance", tol)
        sol = capi.solve_distributed(p, s, comm)
        st = sol.stats()
        q.put(dict(rank=rank, rc=sol.return_code, err=sol.error_string, status=sol.termination_status,
                   its=st.number_of_steps_taken, obj=st.primal_objective, dobj=st.dual_objective, x=sol.primal(),
                   y=sol.dual(), rows=(r0, r1), rp=st.l2_primal_residual))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [1, 3])
def test_two_gpu_solve_matches_single_gpu(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from cuopt_b200 import capi, lpgen
    size, tol, world = 40_000, 1e-6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, size, tol, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda d: d["rank"])
    for p in procs:
        p.join(120)
    for r in res:
        assert r["rc"] == 0, r["err"]
    # single-GPU solve of the whole LP in this process
    lp = lpgen.sparse_lp(size, size, 8, seed=21)
    p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=mode)
    s.set("optimality_tolerance", tol)
    one = capi.solve(p, s)
    st1 = one.stats()
    assert one.termination_status == 1
    # every rank reports the same status / iteration count / objectives (identical decisions on all ranks)
    assert res[0]["status"] == res[1]["status"] == 1
    assert res[0]["its"] == res[1]["its"]
    assert res[0]["obj"] == res[1]["obj"] and res[0]["dobj"] == res[1]["dobj"]
    assert np.array_equal(res[0]["x"], res[1]["x"])
    # and agrees with the single-GPU run: objective to 1e-6 (planted optimum known), iterations within a major period
    assert res[0]["obj"] == pytest.approx(lp.optimal_objective, rel=1e-5)
    assert res[0]["obj"] == pytest.approx(st1.primal_objective, rel=1e-5)  # both stop at tolerance 1e-6
    assert abs(res[0]["its"] - st1.number_of_steps_taken) <= max(40, 0.1 * st1.number_of_steps_taken)
    # the dual blocks tile the dual vector
    y = np.concatenate([r["y"] for r in res])
    assert y.shape[0] == lp.m
    assert np.max(np.abs(y - one.dual())) <= 1e-4 * max(1.0, np.max(np.abs(one.dual())))
