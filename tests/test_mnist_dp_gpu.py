"""Data-parallel MNIST CounteRGAN trainer, two gloo ranks sharing cuda:0 (the collective runs on the
host, so no kernel waits on another process):

  * exactness of the gradient averaging: a discriminator-only step on ONE process with batch 2B must equal TWO ranks with
    B samples each after the all-reduce and the 1/world scaling (fp32 plan, reduction-order tolerance).  The generator's
    BatchNorm uses per-replica statistics, so its influence is removed with an all-zero mask: x_cf = clamp(x + 0) = x;
  * the phase split (D grads | all-reduce | D update + G grads | all-reduce | G update) and its three graph segments:
    two ranks fed IDENTICAL batches reproduce the single-process trainer bit for bit.
"""
import os
import socket
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mods(seed=0):
    from pcg_b200.mnist.models.generator import ResidualGenerator
    from pcg_b200.mnist.models.discriminator import Discriminator
    from pcg_b200.mnist.models.classifier import CNNClassifier
    torch.manual_seed(seed)
    G = ResidualGenerator(base_ch=16, n_resblocks=2).cuda()       # two blocks: the generator backward splits at block 1
    D = Discriminator().cuda()
    C = CNNClassifier().cuda().eval()
    for p in C.parameters():
        p.requires_grad = False
    return G, D, C


CFG = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_classes=10, patch_size=7, num_modifiable_patches=10,
                            lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5, lambda_mask=2.0)


def _worker(rank, world, port, use_graph, overlap, out):
    import sys
    if overlap:        # the optional overlap schemes (trainer.py): deferred classifier input gradient, two G buckets
        os.environ.update(PCG_DP_DEFER_CBWD="1", PCG_DP_SPLIT="1")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import mnist_countergan as O
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import trainer as T
    torch.cuda.set_device(0)
    B = 4
    x, y, t, m = (v.cuda().contiguous() for v in O.synth_batch(2 * B, 31))
    zero = torch.zeros_like(m)

    # ---- single process, batch 2B, discriminator gradients only
    G, D, C = _mods()
    tr = T.CounterGanTrainer(G, D, C, CFG, "cuda", precision="fp32", use_graph=False)
    tr.plan(2 * B).step_d_grads(x, y, t, zero)
    torch.cuda.synchronize()
    g_single = tr.da.grad.clone()

    # ---- single process, whole iterations (identical-batch reference)
    def whole(use_dist_note):
        G, D, C = _mods()
        tr = T.CounterGanTrainer(G, D, C, CFG, "cuda", precision="fp32", use_graph=use_graph)
        if use_graph:
            T._warm_plan(tr, 2 * B)
        for _ in range(2):
            p = tr.step(x, y, t, m)
        torch.cuda.synchronize()
        return tr.ga.data.clone(), tr.da.data.clone(), p.scalars.clone(), tr.world

    single = whole(False)
    assert single[3] == 1
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    # ---- two ranks, B samples each
    G, D, C = _mods()
    tr = T.CounterGanTrainer(G, D, C, CFG, "cuda", precision="fp32", use_graph=False)
    assert tr.world == world
    sl = slice(rank * B, (rank + 1) * B)
    tr.plan(B).step_d_grads(x[sl].contiguous(), y[sl].contiguous(), t[sl].contiguous(), zero[sl].contiguous())
    tr._allreduce(tr.da.grad)
    torch.cuda.synchronize()
    g_dp = tr.da.grad / world
    err = ((g_dp - g_single).double().norm() / g_single.double().norm()).item()

    dp = whole(True)
    assert dp[3] == world
    same = all(torch.equal(a, b) for a, b in zip(single[:3], dp[:3]))
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, err, same, float((single[0] - dp[0]).abs().max()), float((single[1] - dp[1]).abs().max())))


@pytest.mark.parametrize("use_graph,overlap", [(False, False), (True, False), (True, True)])
def test_two_ranks_match_single_process(use_graph, overlap):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, use_graph, overlap, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, same, dg, dd in res:
        assert err < 1e-5, (rank, "D-only 2B vs 2 x B gradient mismatch", err)
        assert same, (rank, "identical-batch DDP run differs from the single-process run", dg, dd)
