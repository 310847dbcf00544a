"""CPU: hazard analysis and stream assignment of the data-flow capture (pcg_b200/dataflow.py).

Recording never launches anything, so it runs on CPU tensors: the operator wrappers of pcg_b200.ops only note their
read / write sets.  Checked: read-after-write, write-after-read, write-after-write on overlapping memory (also slices
of one arena and column windows of one matrix), transitive reduction, stream assignment, and - on random programs - that
EVERY execution order the emitted dependencies allow produces the sequential result."""
import random

import torch

import pcg_b200  # noqa: F401
from pcg_b200 import dataflow as DF
from pcg_b200 import ops as K


def _t(*shape):
    return torch.zeros(*shape)


def test_basic_hazards_and_reduction():
    a, b, c, d = (_t(8, 4) for _ in range(4))

    def body():
        K.unary(a, K.RELU, b)           # 0
        K.unary(a, K.RELU, c)           # 1: independent of 0 (both only read a)
        K.binary(b, c, K.ADD, d)        # 2: RAW on b (0) and c (1)
        K.unary(d, K.RELU, a)           # 3: RAW on d (2); WAR on a (0, 1) is implied by 2 -> reduced away
        K.copy_cols(b, 0, d, 0, 2)      # 4: writes columns 0-1 of d: WAR against 3
        K.copy_cols(c, 0, d, 2, 2)      # 5: writes columns 2-3 of d: WAR against 3, but NOT ordered against 4
        K.unary(d, K.RELU, b)           # 6: reads all of d: after 4 and 5; WAR on b against 4 (implied)
    p = DF.record(body)
    assert [o.deps for o in p.ops] == [[], [], [0, 1], [2], [3], [3], [4, 5]]
    assert p.critical_path() == 5
    assert p.ops[0].stream != p.ops[1].stream and p.ops[4].stream != p.ops[5].stream
    assert p.n_streams == 2


def test_arena_slices_are_tracked_by_byte_range():
    scal, x = _t(16), _t(32)

    def body():
        K.reduce_scalar(x, scal[0:1])                       # 0
        K.reduce_scalar(x, scal[1:2])                       # 1: another slice of the same arena: independent
        K.combine([(1.0, scal[0:1]), (2.0, scal[1:2])], scal[2:3])       # 2: reads both
        K.reduce_scalar(x, scal[0:1])                       # 3: WAR against 2 (and WAW against 0, implied)
    p = DF.record(body)
    assert [o.deps for o in p.ops] == [[], [], [0, 1], [2]]


def test_batchnorm_state_and_in_place_operators():
    y, z, dz, dy = (_t(8, 4) for _ in range(4))
    g, b, rm, rv, dg, db = (_t(4) for _ in range(6))
    nbt = torch.zeros((), dtype=torch.int64)
    st = K.BNState.__new__(K.BNState)
    for n in ("mean", "rstd", "scale", "shift"):
        setattr(st, n, _t(4))
    st.c12, st.scratch, st.scratch2 = _t(8), _t(64), _t(64)
    p_, m_, v_, step = _t(16), _t(16), _t(16), torch.zeros(1, dtype=torch.int32)
    grad = _t(16)

    def body():
        K.bn_train_fwd(y, 8, 4, g, b, rm, rv, nbt, st, z)            # 0: writes the saved statistics
        K.bn_train_bwd(dz, y, 8, 4, g, st, dy, dg, db)               # 1: reads them
        K.bn_train_fwd(y, 8, 4, g, b, rm, rv, nbt, st, z)            # 2: overwrites them: after 1
        K.adam(p_, grad, m_, v_, step, 1e-3)                         # 3: independent
        K.adam(p_, grad, m_, v_, step, 1e-3)                         # 4: in place on the same state: after 3
    p = DF.record(body)
    assert [o.deps for o in p.ops] == [[], [0], [1], [], [3]]


def test_every_allowed_order_gives_the_sequential_result():
    """Random programs of tiny CPU 'operators' recorded through the same machinery; any interleaving of the streams that
    respects the emitted cross-stream dependencies must reproduce the sequential result."""
    rng = random.Random(0)
    for trial in range(20):
        bufs = [torch.arange(6, dtype=torch.float64) + i for i in range(6)]
        arena = torch.zeros(12, dtype=torch.float64)
        views = bufs + [arena[0:4], arena[4:8], arena[2:6]]
        prog = []
        for _ in range(40):
            srcs = rng.sample(range(len(views)), 2)
            dst = rng.randrange(len(views))
            prog.append((srcs, dst, rng.choice([1.0, 0.5, -2.0])))

        def apply(state, srcs, dst, k):
            n = min(state[dst].numel(), state[srcs[0]].numel(), state[srcs[1]].numel())
            state[dst][:n] = k * state[srcs[0]][:n] + state[srcs[1]][:n].flip(0)

        rec = DF.Recorder()
        for (srcs, dst, k) in prog:
            rec.add(None, (srcs, dst, k), {}, [views[s] for s in srcs], [views[dst]])
        p = DF.Program(rec.ops, max_streams=4)

        def fresh():
            bs = [torch.arange(6, dtype=torch.float64) + i for i in range(6)]
            ar = torch.zeros(12, dtype=torch.float64)
            return bs + [ar[0:4], ar[4:8], ar[2:6]], ar
        ref, ref_arena = fresh()
        for (srcs, dst, k) in prog:
            apply(ref, srcs, dst, k)
        for order_seed in range(5):
            orng = random.Random(order_seed)
            state, ar = fresh()
            done, heads = set(), [0] * p.n_streams
            per_stream = [[i for i, o in enumerate(p.ops) if o.stream == s] for s in range(p.n_streams)]
            while len(done) < len(p.ops):
                ready = []
                for s in range(p.n_streams):
                    if heads[s] < len(per_stream[s]):
                        i = per_stream[s][heads[s]]
                        if all(d in done for d in p.ops[i].deps):
                            ready.append((s, i))
                assert ready, "deadlock: the emitted dependencies are cyclic"
                s, i = orng.choice(ready)
                apply(state, *p.ops[i].args)
                done.add(i)
                heads[s] += 1
            assert all(torch.equal(a, b) for a, b in zip(state, ref)) and torch.equal(ar, ref_arena), (trial, order_seed)
