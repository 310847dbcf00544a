"""Executable specification of the operand algebra of the 64->64 halo-tile kernels (csrc/conv_tc64.cu), on CPU.

The CUDA kernels express a 3x3 / pad-1 convolution through *shifted views* of zero-padded pixel tiles held in shared
memory; the bookkeeping (which view feeds which accumulator block with which tap matrix) is the part that cannot be
read off a formula.  These tests replay exactly that bookkeeping with dense torch matmuls on small shapes and compare
with ``F.conv2d`` / autograd, so the tables in the kernel comments stay checked where no GPU is available:

  * conv_tc64s_fprop_kernel: row classes h % 4; an input block of class c is the tap-row r = c - co + 1 operand of the
    output classes co = c-1, c, c+1; wrap-arounds use the class-0 block one row down / the class-3 block one row up;
  * conv_tc64_wgrad_kernel: dW[r][s] = sum_p X[p + r*WP] dY[p - s] with the column shift on the dY side (accumulator 0:
    r = 0, 1 x s = 2, 1, 0) and split between the operands for r = 2 (accumulator 1: s = a + b, a in {0, 1}, b in {2, 0}).
"""
import torch
import torch.nn.functional as F


def _region(x, n, cls, row0, nrows, WP):
    """What the 5-D TMA box delivers: rows row0 .. row0+nrows-1 of class `cls` of image n, columns -1 .. W, zero filled
    outside the image; returned as [nrows * WP + slack, C] linear positions."""
    C, H, W = x.shape[1:]
    out = torch.zeros(nrows * WP + 2 * WP + 4, C, dtype=x.dtype)
    for i in range(nrows):
        idx = row0 + i
        h = 4 * idx + cls
        if idx < 0 or h >= H:
            continue
        out[i * WP + 1:i * WP + 1 + W] = x[n, :, h, :].t()
    return out


def test_row_class_stacked_forward_bookkeeping():
    torch.manual_seed(0)
    N, C, H, W, R = 2, 5, 12, 6, 2          # H % 4 == 0; R rows of each class per super-tile (H / 4 = 3 -> ragged)
    WP, M = W + 2, R * (W + 2)
    x = torch.randn(N, C, H, W, dtype=torch.float64)
    w = torch.randn(C, C, 3, 3, dtype=torch.float64)          # [co_ch, ci, r, s]
    ref = F.conv2d(x, w, padding=1)
    got = torch.zeros_like(ref)
    nidx = H // 4
    # (input class, region row offset rho, first output class, tap rows of the stacked matrices)
    table = [(1, 0, 0, (2, 1, 0)),                      # class 1 -> co 0, 1, 2
             (0, 1, 3, (2,)), (0, 0, 0, (1, 0)),        # class 0: one row down -> co 3; same row -> co 0, 1
             (2, 0, 1, (2, 1, 0)),                      # class 2 -> co 1, 2, 3
             (3, 0, 0, (0,)), (3, 1, 2, (2, 1))]        # class 3 (region starts one row up): up -> co 0; same -> co 2, 3
    for n in range(N):
        for i0 in range(0, nidx, R):
            D = torch.zeros(4, M, C, dtype=torch.float64)                       # accumulator: four output-class blocks
            for cls, rho, co0, taps in table:
                reg = _region(x, n, cls, i0 - 1 if cls == 3 else i0, R + 1, WP)
                for s in range(3):
                    A = reg[rho * WP + s:rho * WP + s + M]                      # shifted view, M positions x C
                    for j, r in enumerate(taps):                                # tap matrices stacked along N
                        D[co0 + j] += A @ w[:, :, r, s].t()
            for co in range(4):
                for pos in range(M):
                    hh, ww = divmod(pos, WP)
                    if ww < W and i0 + hh < nidx:
                        got[n, :, 4 * (i0 + hh) + co, ww] = D[co, pos]
    assert torch.allclose(got, ref, atol=1e-10)


def test_weight_gradient_shifted_views_on_both_operands():
    torch.manual_seed(1)
    N, C, H, W, R = 2, 4, 6, 5, 3
    WP, K = W + 2, R * (W + 2) + 2          # K covers p = pos + s up to R*WP - 1 + 2
    x = torch.randn(N, C, H, W, dtype=torch.float64)
    dy = torch.randn(N, C, H, W, dtype=torch.float64)
    wz = torch.zeros(C, C, 3, 3, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, wz, padding=1), wz, dy)            # [co, ci, r, s]
    dw = torch.zeros(3, 3, C, C, dtype=torch.float64)                            # [r][s][ci][co]
    PAD = 8                                                                      # zero rows in front of the dY tile
    for n in range(N):
        for h0 in range(0, H, R):
            xt = torch.zeros((R + 2) * WP + K + 4, C, dtype=torch.float64)       # halo tile: rows h0-1 .., columns -1 ..
            for i in range(R + 2):
                h = h0 - 1 + i
                if 0 <= h < H:
                    xt[i * WP + 1:i * WP + 1 + W] = x[n, :, h, :].t()
            dt = torch.zeros(PAD + R * WP + K + 4, C, dtype=torch.float64)       # dY tile behind PAD zero rows
            for i in range(R):
                if h0 + i < H:
                    dt[PAD + i * WP:PAD + i * WP + W] = dy[n, :, h0 + i, :].t()
            # accumulator 0: A = X views r = 0, 1; B = dY started 2, 1, 0 rows early (block jb <-> s = 2 - jb)
            for r in (0, 1):
                A = xt[r * WP:r * WP + K]
                for jb in range(3):
                    s = 2 - jb
                    B = dt[PAD - s:PAD - s + K]
                    dw[r, s] += A.t() @ B
            # accumulator 1: r = 2, column shift split: A shifted by a = 0, 1; B started b = 2, 0 rows early; s = a + b
            for a in (0, 1):
                A = xt[2 * WP + a:2 * WP + a + K]
                for b in (2, 0):
                    if a + b < 3:
                        dw[2, a + b] += A.t() @ dt[PAD - b:PAD - b + K]
    assert torch.allclose(dw.permute(3, 2, 0, 1), ref, atol=1e-10)


def _dgrad_s2_by_parity_classes(dy, w, H, W, ksize):
    """The geometry conv_tc_dgrad_s2 (csrc/conv_tc.cu) hands to the implicit-GEMM kernel, replayed densely: the data
    gradient of a stride-2 / pad-1 convolution as four stride-1 problems, one per (hi % 2, wi % 2) class of input
    pixels.  Class (ph, pw) has Ah x Aw pixels; pixel (a, b) reads a th x tw window of dY starting at (a + lo_h, b + lo_w)
    (zero outside), and window element (t, u) carries the filter tap (r, s) that maps it onto input row 2a + ph:
    hi + 1 - r = 2 * ho  ->  r = hi + 1 - 2 * (a + lo_h + t)."""
    N, Cout, Ho, Wo = dy.shape
    Cin = w.shape[1]
    dx = torch.zeros(N, Cin, H, W, dtype=dy.dtype)
    for cls in range(4):
        ph, pw = cls >> 1, cls & 1
        Ah, Aw = (H + 1 - ph) // 2, (W + 1 - pw) // 2
        th = 2 if ksize == 4 else (2 if ph else 1)
        tw = 2 if ksize == 4 else (2 if pw else 1)
        lo_h = -1 if (ksize == 4 and ph == 0) else 0
        lo_w = -1 if (ksize == 4 and pw == 0) else 0
        for a in range(Ah):
            for b in range(Aw):
                acc = torch.zeros(N, Cin, dtype=dy.dtype)
                for t in range(th):
                    for u in range(tw):
                        ho, wo = a + lo_h + t, b + lo_w + u
                        r, s = 2 * a + ph + 1 - 2 * ho, 2 * b + pw + 1 - 2 * wo
                        assert 0 <= r < ksize and 0 <= s < ksize          # every window element is a real filter tap
                        if 0 <= ho < Ho and 0 <= wo < Wo:
                            acc += dy[:, :, ho, wo] @ w[:, :, r, s]
                dx[:, :, 2 * a + ph, 2 * b + pw] = acc
    return dx


def test_stride2_data_gradient_parity_classes():
    torch.manual_seed(2)
    for ksize, H in ((3, 7), (3, 14), (4, 8), (4, 16)):       # 3x3: MNIST discriminator / classifier; 4x4: DCGAN
        N, Cin, Cout, W = 2, 3, 4, H
        Ho = (H + 2 - ksize) // 2 + 1
        x = torch.zeros(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
        w = torch.randn(Cout, Cin, ksize, ksize, dtype=torch.float64)
        dy = torch.randn(N, Cout, Ho, Ho, dtype=torch.float64)
        (ref,) = torch.autograd.grad(F.conv2d(x, w, stride=2, padding=1), x, dy)
        assert torch.allclose(_dgrad_s2_by_parity_classes(dy, w, H, W, ksize), ref, atol=1e-10), (ksize, H)
