"""Launch floor of a CUDA graph made of tiny dependent kernels (what bounds the op-composed tabular plans), and how much
of it parallel branches recover.   python tools/bench_graph_floor.py [nodes]
A node is one libpcg element-wise launch on a 1 KB tensor (pcg_unary SCALE): its work is nil, its cost is the
kernel -> kernel dependency latency inside the graph.  chains = 1: all nodes in one dependency chain; chains = c: the
same number of nodes captured as c independent chains on c streams (forked from and joined to the capture stream)."""
import json
import sys
sys.path.insert(0, '.')
import torch
import pcg_b200  # noqa: F401
from pcg_b200 import graphs, ops as K

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 896
out = {}
for chains in (1, 2, 4, 8, 16):
    bufs = [torch.zeros(256, device="cuda") for _ in range(chains)]
    streams = [torch.cuda.Stream() for _ in range(chains)]
    per = nodes // chains

    def body():
        main = torch.cuda.current_stream()
        for c in range(chains):
            s = main if c == 0 else streams[c]
            if c:
                s.wait_stream(main)
            with torch.cuda.stream(s):
                for _ in range(per):
                    K.unary(bufs[c], K.SCALE, bufs[c], 1.0)
        for c in range(1, chains):
            main.wait_stream(streams[c])
    body()
    torch.cuda.synchronize()
    g = graphs.capture(body)
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[chains] = {"nodes": per * chains, "ms_per_replay": round(ms, 4), "us_per_node": round(ms * 1e3 / (per * chains), 3)}
    print(f"chains={chains:2d}  nodes={per * chains}  {ms:.3f} ms / replay  = {ms * 1e3 / (per * chains):.2f} us per node")
print(json.dumps(out))
