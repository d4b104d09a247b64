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


def _worker(rank, world, port, q, size, extra_cols, tol, mode, transport, block_bytes):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      CUOPT_B200_DIST_MODE=transport)
    if block_bytes:
        os.environ["CUOPT_B200_GATHER_BLOCK_BYTES"] = str(block_bytes)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cuopt_b200 import capi, lpgen
        from cuopt_b200 import dist as cdist
        lp = lpgen.sparse_lp(size, size + extra_cols, 8, seed=21)
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


def _solve_on_gpus(world, size, tol, mode, transport, extra_cols=0, block_bytes=0):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, size, extra_cols, tol, mode, transport, block_bytes))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda d: d["rank"])
        for p in procs:
            p.join(60)
    finally:
        for p in procs:  # a rank that trapped or hung must not outlive the test
            if p.is_alive():
                p.kill()
    for r in res:
        assert r["rc"] == 0, r["err"]
    # every rank reports the same status / iteration count / objectives / primal vector (identical decisions everywhere)
    for r in res[1:]:
        assert r["status"] == res[0]["status"] and r["its"] == res[0]["its"]
        assert r["obj"] == res[0]["obj"] and r["dobj"] == res[0]["dobj"]
        assert np.array_equal(r["x"], res[0]["x"])
    return res


def _single_gpu(size, tol, mode):
    from cuopt_b200 import capi, lpgen
    lp = lpgen.sparse_lp(size, size, 8, seed=21)
    p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=mode)
    s.set("optimality_tolerance", tol)
    one = capi.solve(p, s)
    assert one.termination_status == 1
    return lp, one


# transport of the sharded attempt: p2p = NVLink peer stores fused into the kernels (default), nccl = all-gather +
# reduce-scatter, allreduce = replicated primal side (scheme (i))
@pytest.mark.parametrize("mode,transport", [(1, "p2p"), (1, "nccl"), (1, "allreduce"), (3, "p2p")])
def test_two_gpu_solve_matches_single_gpu(mode, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    size, tol, world = 40_000, 1e-6, 2
    res = _solve_on_gpus(world, size, tol, mode, transport)
    lp, one = _single_gpu(size, tol, mode)
    st1 = one.stats()
    assert res[0]["status"] == 1
    # agrees with the single-GPU run: objective to 1e-5 (planted optimum known; both stop at tolerance 1e-6),
    # the iteration count only loosely: partial sums are added in a different order, and at 1e-6 on this degenerate LP
    # the restart decisions amplify last-bit differences (measured: 18 360 .. 21 200 against 19 400 on one GPU,
    # 19 228 .. 29 640 against 27 512 in Fast1; scripts/dist_iteration_table.py)
    assert res[0]["obj"] == pytest.approx(lp.optimal_objective, rel=1e-5)
    assert res[0]["obj"] == pytest.approx(st1.primal_objective, rel=1e-5)
    assert res[0]["dobj"] == pytest.approx(st1.dual_objective, rel=1e-5)
    assert abs(res[0]["its"] - st1.number_of_steps_taken) <= max(40, 0.4 * st1.number_of_steps_taken)
    # the dual blocks tile the dual vector.  The vectors themselves are NOT compared with the single-GPU ones: the
    # planted LP is degenerate (half of x* sits on its bound), its optimal dual face is not a point, and two
    # tolerance-1e-6 runs land on it 17 % apart in norm (measured) while agreeing on both objectives to 1e-5.
    y = np.concatenate([r["y"] for r in res])
    assert y.shape[0] == lp.m
    assert res[0]["rp"] <= 1e-6 * (1.0 + np.linalg.norm(np.where(np.isfinite(lp.con_ub), lp.con_ub, lp.con_lb))) * 10


@pytest.mark.parametrize("world", [4, 8])
def test_four_and_eight_gpu_solve_matches_single_gpu(world):
    """The same comparison on 4 and 8 ranks (peer-store transport; slices of 1/4 and 1/8 of the columns, rank-ordered sums of
    4 / 8 partials): run with gpurun --gpus 4 / 8, skipped on smaller boxes."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    size, tol = 40_000, 1e-6
    res = _solve_on_gpus(world, size, tol, 1, "p2p")
    lp, one = _single_gpu(size, tol, 1)
    st1 = one.stats()
    assert res[0]["status"] == 1
    assert res[0]["obj"] == pytest.approx(lp.optimal_objective, rel=1e-5)
    assert res[0]["obj"] == pytest.approx(st1.primal_objective, rel=1e-5)
    assert res[0]["dobj"] == pytest.approx(st1.dual_objective, rel=1e-5)
    assert abs(res[0]["its"] - st1.number_of_steps_taken) <= max(40, 0.4 * st1.number_of_steps_taken)
    y = np.concatenate([r["y"] for r in res])
    assert y.shape[0] == lp.m and all(r["rows"][1] - r["rows"][0] == len(r["y"]) for r in res)


def test_peer_store_transport_is_deterministic_and_equals_nccl_transport():
    """With two ranks a + b == b + a, so the peer-store transport (rank-ordered sums) and the NCCL transport must
    produce bit-identical iterates; and the peer-store transport must reproduce itself run to run (no race between the
    NVLink stores, the flags and the consuming kernels)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    size, tol, world = 40_000, 1e-6, 2
    ragged = 37  # n = 40037: the last slice is shorter than the 32-aligned slice width
    a = _solve_on_gpus(world, size, tol, 1, "p2p", ragged)
    b = _solve_on_gpus(world, size, tol, 1, "p2p", ragged)
    c = _solve_on_gpus(world, size, tol, 1, "nccl", ragged)
    assert a[0]["status"] == 1
    for other in (b, c):
        assert other[0]["its"] == a[0]["its"]
        assert other[0]["obj"] == a[0]["obj"] and other[0]["dobj"] == a[0]["dobj"]
        assert np.array_equal(other[0]["x"], a[0]["x"])
        assert all(np.array_equal(other[r]["y"], a[r]["y"]) for r in range(world))


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_gpu_solve_with_gather_blocking(transport):
    """The large-LP kernels (column-blocked passes + element-wise epilogues / scatter) inside the sharded attempt:
    forced on a small LP (4 blocks for A_g, 2 for A_g^T), they must reach the same optimum as the fused kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    size, tol, world = 40_000, 1e-4, 2
    fused = _solve_on_gpus(world, size, tol, 1, transport)
    blocked = _solve_on_gpus(world, size, tol, 1, transport, block_bytes=100_000)
    assert fused[0]["status"] == blocked[0]["status"] == 1
    assert abs(blocked[0]["its"] - fused[0]["its"]) <= max(40, 0.4 * fused[0]["its"])
    assert blocked[0]["obj"] == pytest.approx(fused[0]["obj"], rel=1e-3)  # both are tolerance-1e-4 points
    lp, one = _single_gpu(size, tol, 1)
    assert blocked[0]["obj"] == pytest.approx(lp.optimal_objective, rel=1e-3)
