"""CPU (no GPU): the algebra of the fp32 mode's split operands (csrc/precise.cu, gtos_b200/ops32.py), restated with torch
CPU bf16 rounding - the layout gtos_split3 writes, the role pairing of the three GEMM forms the mode uses (forward,
input gradient with a transposed weight, weight gradient on the row-stacked view) and the error bound the mode rests on.
The GPU tests (tests/test_gpu_fp32_mode.py) hold the kernels to the same statements."""
import torch


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def up8(n):
    return (n + 7) // 8 * 8


def split3(x, role):
    """what gtos_split3 writes: [rows, 3 kp] thirds (hi, lo, hi) for role 0, (hi, hi, lo) for role 1, zero pads"""
    rows, cols = x.shape
    kp = up8(cols)
    hi = bf(x)
    lo = bf(x - hi)
    z = torch.zeros(rows, kp - cols)
    a, b, c = (hi, lo, hi) if role == 0 else (hi, hi, lo)
    return torch.cat([a, z, b, z, c, z], 1)


def stack(x3):
    rows, k3 = x3.shape
    return x3.reshape(rows * 3, k3 // 3)


def rel_err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def test_split_halves_carry_sixteen_significand_bits():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 300, generator=g) * torch.logspace(-6, 6, 300)
    hi, lo = bf(x), bf(x - bf(x))
    assert ((x - hi - lo).abs() <= x.abs() * 2.0 ** -16).all()
    assert torch.equal((x - hi), (x - hi).to(torch.float32))            # the residual is exact in fp32 before rounding
    s = split3(x, 0)
    assert s.shape == (64, 3 * 304) and torch.equal(s[:, 300:304], torch.zeros(64, 4))


def test_the_three_gemm_forms_pair_roles_correctly_and_reach_fp32_accuracy():
    g = torch.Generator().manual_seed(1)
    for M, K, N in [(50, 100, 36), (200, 512, 96), (33, 300, 40)]:
        x = torch.randn(M, K, generator=g) * 3
        W = torch.randn(N, K, generator=g)
        dy = torch.randn(M, N, generator=g)
        ref_y = x.double() @ W.double().t()
        ref_dx = dy.double() @ W.double()
        ref_dW = dy.double().t() @ x.double()
        xs, Ws = split3(x, 0), split3(W, 1)                              # activations role 0, weights role 1
        y = xs @ Ws.t()                                                  # ONE GEMM over the tripled K
        hi_x, lo_x, hi_w, lo_w = bf(x), bf(x - bf(x)), bf(W), bf(W - bf(W))
        three = hi_x @ hi_w.t() + lo_x @ hi_w.t() + hi_x @ lo_w.t()
        assert rel_err(y, three) < 1e-6                                  # exactly the three wanted terms, no lo * lo
        assert rel_err(y, ref_y) < 2e-5
        assert rel_err(bf(x) @ bf(W).t(), ref_y) > 50 * rel_err(y, ref_y)     # plain bf16 operands: ~2^-9
        dys, Wts = split3(dy, 1), split3(W.t().contiguous(), 0)          # gradients role 1, transposed weights role 0
        assert rel_err(dys @ Wts.t(), ref_dx) < 2e-5
        dW = stack(dys)[:, :N].t() @ stack(xs)[:, :K]                    # weight gradient: rows stacked, roles 1 x 0
        assert rel_err(dW, ref_dW) < 2e-5
        # pairing two operands of the SAME role would lose a cross term (hi*hi counted twice, one hi*lo missing)
        bad = split3(x, 0) @ split3(W, 0).t()
        assert rel_err(bad, ref_y) > 20 * rel_err(y, ref_y)
