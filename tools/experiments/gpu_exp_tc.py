"""GPU bring-up experiments for the tcgen05 convolution kernels (run under gpurun, each mode in
its own process with a timeout so a deadlocked kernel cannot hang the box)."""
import ctypes
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import pcg_b200  # noqa: E402
from pcg_b200 import _lib  # noqa: E402

L = _lib.load()
P = _lib.ptr
dev = "cuda"


def sync():
    torch.cuda.synchronize()


def nhwc_bf16(x):  # NCHW fp32 -> NHWC bf16 contiguous
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def pack(w):
    Cout, Cin, k, _ = w.shape
    f = torch.empty(Cout, k * k, Cin, dtype=torch.bfloat16, device=dev)
    d = torch.empty(Cin, k * k, Cout, dtype=torch.bfloat16, device=dev)
    _lib.check(L.pcg_pack_conv_weights_tc(P(w), Cout, Cin, k, P(f), P(d), _lib.stream_ptr()))
    return f, d


def exp_im2col():
    torch.manual_seed(0)
    N, H, W, C = 3, 28, 28, 64
    x = torch.randn(N, C, H, W, device=dev)
    xn = nhwc_bf16(x)
    bad = 0
    for (first, r, s, stride) in [(0, 0, 0, 1), (0, 1, 1, 1), (128, 2, 0, 1), (784 - 20, 0, 2, 1), (2304, 2, 2, 1),
                                  (0, 0, 0, 2), (128, 1, 2, 2)]:
        out = torch.zeros(128, 64, dtype=torch.bfloat16, device=dev)
        _lib.check(L.pcg_debug_im2col_tile(P(xn), N, H, W, C, 3, stride, 1, first, r, s, 0, P(out), _lib.stream_ptr()))
        sync()
        Ho = (H + 2 - 3) // stride + 1
        xp = F.pad(xn.float(), (0, 0, 1, 1, 1, 1))  # pad W and H by 1 (NHWC)
        exp = torch.zeros(128, 64, device=dev)
        for i in range(128):
            p = first + i
            n, rem = divmod(p, Ho * Ho)
            if n >= N:
                continue
            ho, wo = divmod(rem, Ho)
            exp[i] = xp[n, ho * stride + r, wo * stride + s]
        err = (out.float() - exp).abs().max().item()
        print(f"im2col first={first} tap=({r},{s}) stride={stride}: max err {err}")
        bad += err > 0
    print("IM2COL", "OK" if bad == 0 else "FAIL")


def exp_fprop():
    torch.manual_seed(1)
    for (N, Cin, Cout, stride, HW) in [(8, 64, 64, 1, 28), (3, 64, 64, 1, 28), (8, 64, 128, 2, 14), (8, 128, 256, 2, 7),
                                       (16, 256, 256, 2, 4)]:
        x = torch.randn(N, Cin, HW, HW, device=dev)
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * (2.0 / (Cin * 9)) ** 0.5
        b = torch.randn(Cout, device=dev) * 0.1
        xn = nhwc_bf16(x)
        wf, wd = pack(w)
        Ho = (HW + 2 - 3) // stride + 1
        M = N * Ho * Ho
        out = torch.full((N, Ho, Ho, Cout), 7.0, dtype=torch.bfloat16, device=dev)
        grid = L.pcg_conv_tc_grid(ctypes.c_longlong(M), Cout)
        stats = torch.zeros(grid, 2 * Cout, device=dev) if Cout <= 256 and Cout in (64, 128, 256) else None
        add = torch.randn(N, Ho, Ho, Cout, device=dev).to(torch.bfloat16)
        ref = F.conv2d(xn.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, stride=stride, padding=1)
        for mode in ("plain", "lrelu_add"):
            t0 = time.time()
            if mode == "plain":
                _lib.check(L.pcg_conv_tc_fprop(P(xn), N, HW, HW, Cin, P(wf), Cout, 3, stride, 1, P(b), 0,
                                               ctypes.c_float(0.2), None, P(out), P(stats), _lib.stream_ptr()))
                exp = ref
            else:
                _lib.check(L.pcg_conv_tc_fprop(P(xn), N, HW, HW, Cin, P(wf), Cout, 3, stride, 1, P(b), 1,
                                               ctypes.c_float(0.2), P(add), P(out), None, _lib.stream_ptr()))
                exp = F.leaky_relu(ref, 0.2) + add.float().permute(0, 3, 1, 2)
            sync()
            got = out.float().permute(0, 3, 1, 2)
            err = (got - exp).abs().max().item()
            print(f"fprop N={N} {Cin}->{Cout} s{stride} {HW}x{HW} {mode}: max abs err {err:.4e} "
                  f"(ref max {exp.abs().max().item():.3f}) {time.time() - t0:.3f}s")
            if mode == "plain" and stats is not None:
                s = stats.sum(0)
                es = ref.sum(dim=(0, 2, 3))
                eq = (ref * ref).sum(dim=(0, 2, 3))
                print("   stats rel err sum %.3e sumsq %.3e" % (((s[:Cout] - es).abs().max() / es.abs().max()).item(),
                                                               ((s[Cout:] - eq).abs().max() / eq.abs().max()).item()))
        # dgrad via rotated weights (stride 1 only)
        if stride == 1 and Cin == Cout:
            dy = torch.randn(N, Cout, HW, HW, device=dev)
            dyn = nhwc_bf16(dy)
            dx = torch.empty(N, HW, HW, Cin, dtype=torch.bfloat16, device=dev)
            _lib.check(L.pcg_conv_tc_fprop(P(dyn), N, HW, HW, Cout, P(wd), Cin, 3, 1, 1, None, 0,
                                           ctypes.c_float(0.2), None, P(dx), None, _lib.stream_ptr()))
            sync()
            xr = xn.float().permute(0, 3, 1, 2).requires_grad_(True)
            yy = F.conv2d(xr, w.to(torch.bfloat16).float(), None, padding=1)
            (gx,) = torch.autograd.grad(yy, xr, dyn.float().permute(0, 3, 1, 2))
            err = (dx.float().permute(0, 3, 1, 2) - gx).abs().max().item()
            print(f"   dgrad max abs err {err:.4e} (ref max {gx.abs().max().item():.3f})")
    print("FPROP DONE")


def exp_wgrad():
    torch.manual_seed(2)
    for N in (8, 3, 64):
        HW = 28
        x = torch.randn(N, 64, HW, HW, device=dev)
        dy = torch.randn(N, 64, HW, HW, device=dev) * 0.1
        xn, dyn = nhwc_bf16(x), nhwc_bf16(dy)
        M = N * HW * HW
        g = L.pcg_conv_tc_wgrad_grid(ctypes.c_longlong(M))
        part = torch.zeros(g, 9 * 64 * 64, device=dev)
        dw = torch.zeros(64, 64, 3, 3, device=dev)
        _lib.check(L.pcg_conv_tc_wgrad64(P(xn), P(dyn), N, HW, HW, P(part), P(dw), _lib.stream_ptr()))
        sync()
        w = torch.zeros(64, 64, 3, 3, device=dev, requires_grad=True)
        yy = F.conv2d(xn.float().permute(0, 3, 1, 2), w, None, padding=1)
        (gw,) = torch.autograd.grad(yy, w, dyn.float().permute(0, 3, 1, 2))
        err = (dw - gw).abs().max().item()
        print(f"wgrad N={N}: max abs err {err:.4e} (ref max {gw.abs().max().item():.3f})")
    print("WGRAD DONE")


def exp_perf():
    torch.manual_seed(3)
    N, HW = 512, 28
    xn = torch.randn(N, HW, HW, 64, device=dev).to(torch.bfloat16)
    w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
    b = torch.zeros(64, device=dev)
    wf, wd = pack(w)
    out = torch.empty_like(xn)
    M = N * HW * HW
    grid = L.pcg_conv_tc_grid(ctypes.c_longlong(M), 64)
    stats = torch.zeros(grid, 128, device=dev)
    g = L.pcg_conv_tc_wgrad_grid(ctypes.c_longlong(M))
    part = torch.zeros(g, 9 * 64 * 64, device=dev)
    dw = torch.zeros(64, 64, 3, 3, device=dev)
    st = _lib.stream_ptr()

    def run_f():
        _lib.check(L.pcg_conv_tc_fprop(P(xn), N, HW, HW, 64, P(wf), 64, 3, 1, 1, P(b), 0, ctypes.c_float(0.2), None,
                                       P(out), P(stats), st))

    def run_w():
        _lib.check(L.pcg_conv_tc_wgrad64(P(xn), P(out), N, HW, HW, P(part), P(dw), st))

    for name, fn in (("fprop+stats", run_f), ("wgrad", run_w)):
        for _ in range(3):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / 20
        fl = 2.0 * M * 64 * 576
        print(f"{name}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  ({(2 * M * 64 * 2) / ms / 1e6:.0f} GB/s act traffic)")
    # cuDNN reference point (bf16 channels_last)
    xc = xn.permute(0, 3, 1, 2)
    wc = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        F.conv2d(xc, wc, None, padding=1)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        F.conv2d(xc, wc, None, padding=1)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / 20
    print(f"cudnn bf16 NHWC fprop: {ms * 1e3:.1f} us  {2.0 * M * 64 * 576 / ms / 1e9:.1f} TFLOP/s")
    print("PERF DONE")


def exp_v2():
    for variant in (0, 1):
        L.pcg_conv_tc64_set_variant(variant)
        torch.manual_seed(4)
        print("=== variant", variant)
        for (N, HW) in [(8, 28), (3, 28), (5, 14)]:
            x = torch.randn(N, 64, HW, HW, device=dev)
            w = torch.randn(64, 64, 3, 3, device=dev) * (2.0 / 576) ** 0.5
            b = torch.randn(64, device=dev) * 0.1
            xn = nhwc_bf16(x)
            wf, wd = pack(w)
            grid = L.pcg_conv_tc64_grid(N, HW, HW)
            out = torch.full((N, HW, HW, 64), 7.0, dtype=torch.bfloat16, device=dev)
            stats = torch.zeros(grid, 128, device=dev)
            _lib.check(L.pcg_conv_tc64_fprop(P(xn), N, HW, HW, P(wf), P(b), 0, ctypes.c_float(0.2), None, None, 0, P(out),
                                             P(stats), _lib.stream_ptr()))
            sync()
            ref = F.conv2d(xn.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=1)
            got = out.float().permute(0, 3, 1, 2)
            s = stats.sum(0)
            print(f"v2 fprop N={N} {HW}x{HW}: max abs err {(got - ref).abs().max().item():.4e} (ref max {ref.abs().max().item():.2f})"
                  f"  stats rel err {((s[:64] - ref.sum(dim=(0, 2, 3))).abs().max() / ref.sum(dim=(0, 2, 3)).abs().max()).item():.2e}"
                  f" {((s[64:] - (ref * ref).sum(dim=(0, 2, 3))).abs().max() / (ref * ref).sum(dim=(0, 2, 3)).abs().max()).item():.2e}")
            dy = torch.randn(N, 64, HW, HW, device=dev) * 0.1
            dyn = nhwc_bf16(dy)
            part = torch.zeros(grid, 9 * 64 * 64, device=dev)
            dw = torch.zeros(64, 64, 3, 3, device=dev)
            _lib.check(L.pcg_conv_tc64_wgrad(P(xn), P(dyn), N, HW, HW, P(part), P(dw), _lib.stream_ptr()))
            sync()
            wz = torch.zeros(64, 64, 3, 3, device=dev, requires_grad=True)
            yy = F.conv2d(xn.float().permute(0, 3, 1, 2), wz, None, padding=1)
            (gw,) = torch.autograd.grad(yy, wz, dyn.float().permute(0, 3, 1, 2))
            print(f"v2 wgrad N={N}: max abs err {(dw - gw).abs().max().item():.4e} (ref max {gw.abs().max().item():.2f})")
    print("V2 DONE")


def exp_v2perf():
    L.pcg_conv_tc64_set_variant(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    torch.manual_seed(3)
    N, HW = 512, 28
    bufs = [torch.randn(N, HW, HW, 64, device=dev).to(torch.bfloat16) for _ in range(4)]
    outs = [torch.empty_like(bufs[0]) for _ in range(4)]
    w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
    b = torch.zeros(64, device=dev)
    wf, wd = pack(w)
    grid = L.pcg_conv_tc64_grid(N, HW, HW)
    stats = torch.zeros(grid, 128, device=dev)
    part = torch.zeros(grid, 9 * 64 * 64, device=dev)
    dw = torch.zeros(64, 64, 3, 3, device=dev)
    st = _lib.stream_ptr()
    M = N * HW * HW
    it = {"i": 0}

    def run_f():
        i = it["i"] = (it["i"] + 1) % 4
        _lib.check(L.pcg_conv_tc64_fprop(P(bufs[i]), N, HW, HW, P(wf), P(b), 0, ctypes.c_float(0.2), None, None, 0,
                                         P(outs[i]), P(stats), st))

    def run_d():
        i = it["i"] = (it["i"] + 1) % 4
        _lib.check(L.pcg_conv_tc64_fprop(P(bufs[i]), N, HW, HW, P(wd), None, 0, ctypes.c_float(0.2), P(outs[(i + 1) % 4]),
                                         None, 0, P(outs[i]), None, st))

    def run_w():
        i = it["i"] = (it["i"] + 1) % 4
        _lib.check(L.pcg_conv_tc64_wgrad(P(bufs[i]), P(outs[i]), N, HW, HW, P(part), P(dw), st))

    for name, fn in (("v2 fprop+stats", run_f), ("v2 dgrad+add", run_d), ("v2 wgrad+reduce", run_w)):
        for _ in range(3):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name}: {ms * 1e3:.1f} us  {2.0 * M * 64 * 576 / ms / 1e9:.1f} TFLOP/s")
    print("V2PERF DONE")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    {"im2col": exp_im2col, "fprop": exp_fprop, "wgrad": exp_wgrad, "perf": exp_perf, "v2": exp_v2, "v2perf": exp_v2perf}[sys.argv[1]]()
