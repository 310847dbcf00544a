"""world_size-2 gloo test (CPU) of the data-parallel host logic of ``CounterGanTrainer``.

The native plan needs a GPU, so a stand-in plan records the phase order and fills the gradient arenas
with rank-dependent values.  What is checked is exactly what the host side is responsible for
(SURVEY.md §8e): D's gradients are all-reduced after phase 1 and before the D update (the classifier's deferred input
gradient phase runs in between), G's in two buckets - the arena tail [resblocks.k ..] after part 1 of the backward, the
head after part 2 - and both before the G update; both ranks end up with the SUM (the 1/N lives in Adam's grad_scale),
and no collective is issued when world_size == 1.
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Arena:
    def __init__(self, n, ntensors=0):
        self.data = torch.zeros(n)
        self.grad = torch.zeros(n)
        self.slots = [(i * (n // max(ntensors, 1)), n // max(ntensors, 1)) for i in range(ntensors)]


class _FakePlan:
    def __init__(self, tr, rank, log):
        self.tr, self.rank, self.log = tr, rank, log
        self.seen_d = self.seen_g = None

    def step_d_grads(self, *a):
        self.log.append("d_grads")
        self.tr.da.grad.fill_(1.0 + self.rank)

    def step_d_update(self):
        self.log.append("d_update")
        self.seen_d = self.tr.da.grad.clone()

    def step_g_grads(self, *a):
        self.log.append("g_grads")
        self.tr.ga.grad.fill_(10.0 * (1 + self.rank))

    def set_defer_c_bwd(self, on):
        self.deferred = on

    def step_c_bwd(self):
        self.log.append("c_bwd")

    def step_g_grads_part(self, x, y, t, m, part, split):
        self.log.append(f"g_grads_{part}")
        off = self.tr.ga.slots[3 + 8 * split][0]
        if part == 1:                     # the backward pass finishes the tail of the arena first
            self.tr.ga.grad[off:].fill_(10.0 * (1 + self.rank))
        else:
            self.tr.ga.grad[:off].fill_(10.0 * (1 + self.rank))

    def step_g_update(self):
        self.log.append("g_update")
        self.seen_g = self.tr.ga.grad.clone()

    def step(self, *a):
        self.log.append("fused_step")


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import trainer as T
    tr = T.CounterGanTrainer.__new__(T.CounterGanTrainer)
    tr.ga, tr.da = _Arena(3 + 8 * 6 + 4, 3 + 8 * 6 + 4), _Arena(8)      # generator: 55 tensors of one element each
    tr.n_res = 6
    tr.dist, tr.world = T._world()
    assert tr.world == world
    log = []
    plan = _FakePlan(tr, rank, log)
    # default: one reduction after each backward pass
    tr._run_phases(plan, None, None, None, None)
    assert plan.deferred is False and log == ["d_grads", "d_update", "g_grads", "g_update"], log
    assert torch.all(plan.seen_d == 3.0) and torch.all(plan.seen_g == 30.0)
    # overlap switches: deferred classifier input gradient, generator backward / reduction in two buckets
    os.environ.update(PCG_DP_DEFER_CBWD="1", PCG_DP_SPLIT="1")
    log.clear()
    tr._run_phases(plan, None, None, None, None)
    assert plan.deferred is True
    assert log == ["d_grads", "c_bwd", "d_update", "g_grads_1", "g_grads_2", "g_update"], log
    assert tr._dp_split() == (3, 27)
    # a generator too shallow to split falls back to one reduction after the whole backward
    tr.n_res = 1
    log.clear()
    tr._run_phases(plan, None, None, None, None)
    assert log == ["d_grads", "c_bwd", "d_update", "g_grads", "g_update"], log
    tr.n_res = 6
    # the updates saw the all-reduced SUM over ranks (1+2 and 10+20)
    assert torch.all(plan.seen_d == 3.0) and torch.all(plan.seen_g == 30.0)
    assert abs(tr._step_cfg_grad_scale() - 0.5) < 1e-12
    dist.destroy_process_group()
    ret[rank] = True


def test_two_rank_phase_order_and_allreduce():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + os.getpid() % 200
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


def test_single_rank_uses_fused_step():
    sys.path.insert(0, ROOT)
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import trainer as T
    tr = T.CounterGanTrainer.__new__(T.CounterGanTrainer)
    tr.ga, tr.da = _Arena(4), _Arena(4)
    tr.dist, tr.world = None, 1
    log = []
    tr._run_phases(_FakePlan(tr, 0, log), None, None, None, None)
    assert log == ["fused_step"]
    assert tr._step_cfg_grad_scale() == 1.0
