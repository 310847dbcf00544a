"""Data-flow capture of an operator-composed step plan.

The tabular CounteRGAN plans are ~300-900 libpcg launches per iteration, each a few microseconds of work: as ONE chain
of graph nodes they are bound by the kernel -> kernel dependency latency of a CUDA graph (1.26 us per node on a B200,
``tools/bench_graph_floor.py``), while independent chains of nodes run side by side almost for free (16 chains: 0.29 us
per node).  Most of a training iteration IS independent: weight gradients never feed the data-gradient chain, the three
critic passes and the frozen classifier only meet at a few tensors, the seven categorical heads are parallel.

``record(body)`` runs the plan's body with a recorder installed in ``pcg_b200.ops``: no kernel is launched, every
operator call is noted with the memory it reads and writes (declared per operator in ops.py).  ``Program.emit()``,
called under stream capture, launches the operators in program order but on several streams, each waiting (CUDA events)
for exactly the operators it depends on - read-after-write, write-after-read and write-after-write on overlapping
memory - so the captured graph has the true dependency edges instead of a single chain.  Results are bit-identical to
the sequential order (same kernels, same inputs); ``tests/test_dataflow_*`` check that and the hazard analysis itself.
"""
import torch

from . import ops as K


def _access(v):
    """(allocation, first byte, end byte, row length, first column, end column) of a tensor or of a column window
    ``(tensor, c0, c1)`` of a contiguous 2-D tensor; row length 0 = the whole tensor."""
    t, c0, c1 = (v if isinstance(v, tuple) else (v, 0, 0))
    p = t.data_ptr()
    ld = t.shape[-1] if isinstance(v, tuple) else 0
    return t.untyped_storage().data_ptr(), p, p + t.numel() * t.element_size(), ld, c0, c1


def _conflict(a, b):
    """Two accesses touch common memory: overlapping byte ranges, unless both are column windows of the very same
    tensor (same range, same row length) that do not intersect."""
    if not (a[1] < b[2] and b[1] < a[2]):
        return False
    if a[3] and a[3] == b[3] and a[1] == b[1] and a[2] == b[2]:
        return a[4] < b[5] and b[4] < a[5]
    return True


def _numel(v):
    return (v[0] if isinstance(v, tuple) else v).numel()


class _Op:
    __slots__ = ("fn", "args", "kwargs", "reads", "writes", "deps", "stream", "event", "needs_event")

    def __init__(self, fn, args, kwargs, reads, writes):
        self.fn, self.args, self.kwargs = fn, args, kwargs
        self.reads = [_access(t) for t in reads if _numel(t) > 0]
        self.writes = [_access(t) for t in writes if _numel(t) > 0]
        self.deps, self.stream, self.event, self.needs_event = [], 0, None, False


class Recorder:
    def __init__(self):
        self.ops = []

    def add(self, fn, args, kwargs, reads, writes):
        self.ops.append(_Op(fn, args, kwargs, reads, writes))


def record(body):
    """Runs ``body()`` with operator launches replaced by recording; returns the recorded ``Program``."""
    rec = Recorder()
    prev, K._rec = K._rec, rec
    try:
        body()
    finally:
        K._rec = prev
    return Program(rec.ops)


class Program:
    def __init__(self, ops, max_streams=None):
        import os
        self.ops = ops
        self.max_streams = int(os.environ.get("PCG_DATAFLOW_STREAMS", "32")) if max_streams is None else max_streams
        self._dependencies()
        self._assign_streams()

    # ------------------------------------------------------------------ hazards
    def _dependencies(self):
        """deps[i] = the operators i must wait for, transitively reduced.  Per allocation a list of live accesses
        (access, op, is_write); a whole-tensor write supersedes every access it fully covers."""
        live = {}
        anc = []                                   # ancestor bit sets
        for i, op in enumerate(self.ops):
            deps = set()
            for acc in op.reads:
                for (other, j, w) in live.get(acc[0], ()):
                    if w and _conflict(acc, other):
                        deps.add(j)
            for acc in op.writes:
                for (other, j, w) in live.get(acc[0], ()):
                    if _conflict(acc, other):
                        deps.add(j)
            deps.discard(i)
            # transitive reduction: drop a dependency that is an ancestor of another dependency
            keep = []
            for d in deps:
                if not any(e != d and (anc[e] >> d) & 1 for e in deps):
                    keep.append(d)
            op.deps = sorted(keep)
            a = 0
            for d in deps:
                a |= anc[d] | (1 << d)
            anc.append(a)
            for acc in op.writes:
                ent = live.setdefault(acc[0], [])
                if acc[3] == 0:                    # a whole-tensor write supersedes every access it fully covers
                    ent[:] = [e for e in ent if not (acc[1] <= e[0][1] and e[0][2] <= acc[2] and e[1] != i)]
                ent.append((acc, i, True))
            for acc in op.reads:
                live.setdefault(acc[0], []).append((acc, i, False))
        self._anc = anc

    # ------------------------------------------------------------------ streams
    def _assign_streams(self):
        """Greedy: continue on the stream whose last operator is one of the dependencies (the latest one); otherwise on a
        stream whose last operator is an ancestor anyway (no false edge); otherwise open a new stream; when the pool is
        exhausted reuse the stream that has been idle longest (one false edge)."""
        tails = []                                 # last op index per stream
        for i, op in enumerate(self.ops):
            choice = None
            for d in sorted(op.deps, reverse=True):
                for s, t in enumerate(tails):
                    if t == d:
                        choice = s
                        break
                if choice is not None:
                    break
            if choice is None:
                for s, t in enumerate(tails):
                    if t >= 0 and (self._anc[i] >> t) & 1:
                        choice = s
                        break
            if choice is None:
                if len(tails) < self.max_streams:
                    tails.append(-1)
                    choice = len(tails) - 1
                else:
                    choice = min(range(len(tails)), key=lambda s: tails[s])
            op.stream = choice
            for d in op.deps:
                if self.ops[d].stream != choice:   # same stream: program order per stream already implies it
                    self.ops[d].needs_event = True
            tails[choice] = i
        self.n_streams = len(tails)

    # ------------------------------------------------------------------ statistics
    def critical_path(self):
        depth = []
        for op in self.ops:
            depth.append(1 + max((depth[d] for d in op.deps), default=0))
        return max(depth, default=0)

    def timed_critical_path(self, repeats=8):
        """Times every operator on the device (its own little CUDA graph replayed ``repeats`` times between two events, so
        that host launch latency is out of the number) and returns (longest dependency path in us, [(op index, name, us)]
        along it, total us of all operators): what bounds the emitted graph once launch latency is hidden by the parallel
        branches.  Executes every operator several times in program order: the caller restores any state."""
        from . import graphs
        n = len(self.ops)
        us = [0.0] * n
        prev, K._rec = K._rec, None
        try:
            for i, op in enumerate(self.ops):
                op.fn(*op.args, **op.kwargs)
                g = graphs.capture(lambda: [op.fn(*op.args, **op.kwargs) for _ in range(repeats)], sync=True)
                g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                us[i] = e0.elapsed_time(e1) * 1e3 / repeats
                del g
        finally:
            K._rec = prev
        best, back = [0.0] * n, [-1] * n
        for i, op in enumerate(self.ops):
            for d in op.deps:
                if best[d] > best[i]:
                    best[i], back[i] = best[d], d
            best[i] += us[i]
        end = max(range(n), key=lambda i: best[i]) if n else -1
        path, i = [], end
        while i >= 0:
            path.append((i, getattr(self.ops[i].fn, "__name__", "?"), us[i]))
            i = back[i]
        return (best[end] if n else 0.0), path[::-1], sum(us)

    # ------------------------------------------------------------------ emission
    def emit(self, streams=None):
        """Launches the recorded operators on ``n_streams`` streams (stream 0 = the current one) with event waits for the
        cross-stream dependencies; every side stream forks from and joins the current stream, so the whole program is
        ordered like one launch for the caller (and is one graph under capture)."""
        main = torch.cuda.current_stream()
        if streams is None:
            streams = getattr(self, "_streams", None)
            if streams is None or len(streams) < self.n_streams - 1:
                streams = self._streams = [torch.cuda.Stream() for _ in range(self.n_streams - 1)]
        pool = [main] + list(streams)
        started = [True] + [False] * (self.n_streams - 1)
        last_on = [-1] * self.n_streams
        fork = torch.cuda.Event()
        fork.record(main)
        prev, K._rec = K._rec, None
        try:
            for i, op in enumerate(self.ops):
                s = pool[op.stream]
                if not started[op.stream]:
                    s.wait_event(fork)
                    started[op.stream] = True
                for d in op.deps:
                    dop = self.ops[d]
                    if dop.stream == op.stream:
                        continue                   # stream order
                    s.wait_event(dop.event)
                with torch.cuda.stream(s):
                    op.fn(*op.args, **op.kwargs)
                if op.needs_event:
                    op.event = torch.cuda.Event()
                    op.event.record(s)
                last_on[op.stream] = i
            for k in range(1, self.n_streams):
                if started[k]:
                    main.wait_stream(pool[k])
        finally:
            K._rec = prev
