import ctypes as C, sys, torch
sys.path.insert(0, '.')
from gtos_b200 import _lib, ops
lib = _lib.load(); dev = torch.device('cuda:0'); torch.manual_seed(1)
def rel_err(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()
st = torch.cuda.current_stream().cuda_stream

print("== FFN pieces ==")
M, D, Fd = 300, 128, 256
x = torch.randn(M, D, device=dev, requires_grad=True)
W1 = (torch.randn(Fd, D, device=dev) * 0.1).requires_grad_(); b1 = (torch.randn(Fd, device=dev) * 0.1).requires_grad_()
W2 = (torch.randn(D, Fd, device=dev) * 0.1).requires_grad_(); b2 = (torch.randn(D, device=dev) * 0.1).requires_grad_()
w = torch.randn(M, D, device=dev)
h = ops.ffn(x, None, W1, b1, W2, b2, 0.0)
# reference with bf16-rounded operands to take mask flips out of the picture
xb = x.detach().bfloat16().float().requires_grad_(); W1r = W1.detach().bfloat16().float().requires_grad_(); W2r = W2.detach().bfloat16().float().requires_grad_()
b1r = b1.detach().clone().requires_grad_(); b2r = b2.detach().clone().requires_grad_()
hid = torch.relu(xb @ W1r.t() + b1r).bfloat16().float()
hid_full = torch.relu(xb @ W1r.t() + b1r)
href = hid_full @ W2r.t() + b2r
print("fwd", rel_err(h, href))
gs = torch.autograd.grad((h * w).sum(), [x, W1, b1, W2, b2]); rs = torch.autograd.grad((href * w).sum(), [xb, W1r, b1r, W2r, b2r])
for n, a, r in zip(["dx", "dW1", "db1", "dW2", "db2"], gs, rs): print(n, rel_err(a, r))

print("== attention bwd pieces ==")
for (T, S, B, H, hd) in [(33, 70, 5, 1, 512), (33, 70, 5, 2, 128), (6, 8, 3, 4, 8), (40, 40, 4, 2, 64)]:
    D = H * hd
    q = torch.randn(T, B, D, device=dev, requires_grad=True); k = torch.randn(S, B, D, device=dev, requires_grad=True); v = torch.randn(S, B, D, device=dev, requires_grad=True)
    scale = hd ** -0.5
    qh, kh, vh = q.view(T, B, H, hd), k.view(S, B, H, hd), v.view(S, B, H, hd)
    s = torch.einsum("tbhd,sbhd->bhts", qh, kh) * scale
    p = torch.softmax(s, -1); ref = torch.einsum("bhts,sbhd->tbhd", p, vh).reshape(T, B, D)
    dout = torch.randn_like(ref); (ref * dout).sum().backward()
    probs = torch.empty(B, H, T, S, device=dev); out = torch.empty(T * B, D, device=dev)
    d = ops._attn_desc(T, S, B, H, hd)
    d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D
    d.scale, d.p_drop = scale, 0.0
    d.probs, d.out, d.ldo = probs.data_ptr(), out.data_ptr(), D
    _lib.check(lib.gtos_attn_fwd(C.byref(d), st))
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v); ds = torch.empty(B, H, T, S, device=dev)
    d.dout, d.lddo = dout.data_ptr(), D; d.dscores_ts = ds.data_ptr()
    d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = dq.data_ptr(), D, dk.data_ptr(), D, dv.data_ptr(), D
    _lib.check(lib.gtos_attn_bwd(C.byref(d), st)); torch.cuda.synchronize()
    dsref = torch.autograd.grad((p * 1).sum() * 0 + (ref * dout).sum(), s, retain_graph=False, allow_unused=True) if False else None
    print((T, S, B, H, hd), "probs", rel_err(probs, p), "out", rel_err(out.view(T, B, D), ref), "dq", rel_err(dq, q.grad), "dk", rel_err(dk, k.grad), "dv", rel_err(dv, v.grad))
    if hd > 64:
        for c in range(0, D, 64):
            print("   chunk", c, "dq", rel_err(dq[..., c:c+64], q.grad[..., c:c+64]), "dk", rel_err(dk[..., c:c+64], k.grad[..., c:c+64]), "dv", rel_err(dv[..., c:c+64], v.grad[..., c:c+64]))
