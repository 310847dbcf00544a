"""Data-parallel DCGAN plan: two ranks (gloo, both on cuda:0 - the collective runs on the host, so no
kernel waits on another process) fed IDENTICAL batches must reproduce the single-process run bit for bit: the summed
gradients times 1/world are the single-process gradients, the phase split / graph segments / all-reduce placement
(D's before the G phase, G's before its own update) are what is being tested."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, use_graph, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import dcgan as O
    import pcg_b200  # noqa: F401
    from pcg_b200.dcgan import DcganPlan
    torch.cuda.set_device(0)
    B = 8
    PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
    batches = [O.synth_batch(B, 400 + i) for i in range(2)]

    def run(plan):
        plan.G.load(PG)
        plan.D.load(PD)
        plan.refresh()
        for real, noise in batches:
            sc = plan.step(real.cuda(), noise.cuda()).clone()
        torch.cuda.synchronize()
        return plan.G.data.clone(), plan.D.data.clone(), sc

    single = run(DcganPlan(B, "cuda", use_graph=use_graph, tensor_cores=False))       # before the process group exists
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = DcganPlan(B, "cuda", use_graph=use_graph, tensor_cores=False)
    assert plan.world == world
    dp = run(plan)
    ok = all(torch.equal(a, b) for a, b in zip(single, dp))
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, ok, float((single[0] - dp[0]).abs().max()), float((single[1] - dp[1]).abs().max())))


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_ranks_with_identical_batches_match_single_process(use_graph):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, use_graph, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, dg, dd in res:
        assert ok, (rank, dg, dd)
