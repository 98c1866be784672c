"""GPU probe: attention fwd/bwd and the row-wise kernels vs plain torch fp32. Development tool."""
import ctypes as C, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from vilbert_b200 import _lib as L

lib = L.lib(); dev = torch.device("cuda:0"); torch.manual_seed(0)
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
BF = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20)).item()


def report(name, errs, tol):
    worst = max(errs.values())
    print(f"{'PASS' if worst < tol else 'FAIL'}  {name:50s} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)


def attn_case(B, H, Nq, Nk, D, cross, peaked=1.0, iters=0):
    Hd = H * D
    if cross:
        qsrc = (torch.randn(B * Nq, 3 * Hd, device=dev) * peaked).to(BF)
        ksrc = (torch.randn(B * Nk, 3 * Hd, device=dev) * peaked).to(BF)
    else:
        qsrc = ksrc = (torch.randn(B * Nq, 3 * Hd, device=dev) * peaked).to(BF)
    q, k, v = qsrc[:, :Hd], ksrc[:, Hd:2 * Hd], ksrc[:, 2 * Hd:]
    lens = torch.randint(1, Nk + 1, (B,), device=dev); lens[0] = Nk
    mask = ((torch.arange(Nk, device=dev)[None] >= lens[:, None]).float() * -10000.0).contiguous()
    O = torch.zeros(B * Nq, Hd, device=dev, dtype=BF); lse = torch.zeros(B, H, Nq, device=dev)
    dO = (torch.randn(B * Nq, Hd, device=dev)).to(BF)
    dqb = torch.zeros(B * Nq, 3 * Hd, device=dev, dtype=BF); dkb = torch.zeros(B * Nk, 3 * Hd, device=dev, dtype=BF)
    delta = torch.zeros(B, H, Nq, device=dev)
    a = L.AttnArgs()
    a.B, a.H, a.Nq, a.Nk, a.D = B, H, Nq, Nk, D
    a.Q, a.ldq, a.K, a.ldk, a.V, a.ldv = q.data_ptr(), 3 * Hd, k.data_ptr(), 3 * Hd, v.data_ptr(), 3 * Hd
    a.mask, a.scale = mask.data_ptr(), 1.0 / math.sqrt(D)
    a.O, a.ldo, a.lse = O.data_ptr(), Hd, lse.data_ptr()
    a.dO, a.lddo = dO.data_ptr(), Hd
    a.dQ, a.lddq = dqb[:, :Hd].data_ptr(), 3 * Hd
    a.dK, a.lddk = dkb[:, Hd:2 * Hd].data_ptr(), 3 * Hd
    a.dV, a.lddv = dkb[:, 2 * Hd:].data_ptr(), 3 * Hd
    a.delta = delta.data_ptr()
    L.check(lib.vb_attention_fwd(C.byref(a), ST())); L.check(lib.vb_attention_bwd(C.byref(a), ST()))
    torch.cuda.synchronize()
    # reference
    qf = q.float().view(B, Nq, H, D).permute(0, 2, 1, 3).detach().requires_grad_(True)
    kf = k.float().view(B, Nk, H, D).permute(0, 2, 1, 3).detach().requires_grad_(True)
    vf = v.float().view(B, Nk, H, D).permute(0, 2, 1, 3).detach().requires_grad_(True)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D) + mask[:, None, None, :]
    p = torch.softmax(s, -1)
    o = (p @ vf).permute(0, 2, 1, 3).reshape(B * Nq, Hd)
    o.backward(dO.float())
    errs = dict(O=rel(O, o), dQ=rel(dqb[:, :Hd].view(B, Nq, H, D).permute(0, 2, 1, 3), qf.grad),
                dK=rel(dkb[:, Hd:2 * Hd].view(B, Nk, H, D).permute(0, 2, 1, 3), kf.grad),
                dV=rel(dkb[:, 2 * Hd:].view(B, Nk, H, D).permute(0, 2, 1, 3), vf.grad),
                lse=rel(lse * math.log(2.0), torch.logsumexp(s, -1)))
    msg = f"attn B{B} H{H} Nq{Nq} Nk{Nk} D{D} cross{int(cross)} peak{peaked}"
    if iters:
        for fn, nm in ((lib.vb_attention_fwd, "fwd"), (lib.vb_attention_bwd, "bwd")):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            for _ in range(3): fn(C.byref(a), ST())
            e0.record()
            for _ in range(iters): fn(C.byref(a), ST())
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 4.0 * B * H * Nq * Nk * D * (1 if nm == "fwd" else 2.5)
            msg += f" {nm} {ms*1e3:.1f}us {fl/ms/1e9:.1f}TF"
    report(msg, errs, 2e-2)


print("=== attention", torch.cuda.get_device_name(0), flush=True)
for args in [(2, 4, 9, 9, 16, False), (3, 3, 11, 11, 32, False), (2, 2, 12, 7, 32, True), (2, 2, 7, 12, 32, True),
             (4, 12, 36, 36, 64, False), (4, 8, 100, 100, 128, False), (4, 8, 36, 100, 128, True), (4, 8, 100, 36, 128, True),
             (2, 8, 306, 306, 128, False), (2, 8, 257, 306, 128, True), (2, 12, 257, 257, 64, False), (3, 8, 65, 129, 128, True)]:
    attn_case(*args)
attn_case(4, 8, 100, 100, 128, False, peaked=4.0)
attn_case(4, 12, 36, 36, 64, False, peaked=6.0)
attn_case(64, 12, 36, 36, 64, False, iters=20)
attn_case(64, 8, 100, 100, 128, False, iters=20)
attn_case(64, 8, 36, 100, 128, True, iters=20)
attn_case(64, 8, 100, 36, 128, True, iters=20)

print("=== layernorm")
for (M, H) in [(37, 64), (50, 96), (2304, 768), (6400, 1024), (33, 2048), (128, 128)]:
    x = torch.randn(M, H, device=dev) * 2 + 0.5; g = torch.randn(H, device=dev); b = torch.randn(H, device=dev)
    y32 = torch.empty(M, H, device=dev); y16 = torch.empty(M, H, device=dev, dtype=BF); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    L.check(lib.vb_layernorm_fwd(x.data_ptr(), H, g.data_ptr(), b.data_ptr(), 1e-12, y32.data_ptr(), y16.data_ptr(), H, mean.data_ptr(), rstd.data_ptr(), M, H, None, 0, None, None, ST()))
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (H,), gr, br, 1e-12)
    dy = torch.randn(M, H, device=dev); yr.backward(dy)
    dx32 = torch.empty(M, H, device=dev); dx16 = torch.empty(M, H, device=dev, dtype=BF); dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev)
    L.check(lib.vb_layernorm_bwd(dy.data_ptr(), H, x.data_ptr(), H, g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx32.data_ptr(), dx16.data_ptr(), H, None, 0,
                                 dg.data_ptr(), db.data_ptr(), None, M, H, None, None, ST()))
    torch.cuda.synchronize()
    report(f"layernorm M{M} H{H}", dict(y=rel(y32, yr), y16=max(rel(y16, yr) - 4e-3, 0), dx=rel(dx32, xr.grad), dg=rel(dg, gr.grad), db=rel(db, br.grad)), 1e-4)
# LN bwd with gelu' fusion
M, H = 64, 2048
x = torch.randn(M, H, device=dev); g = torch.randn(H, device=dev); b = torch.randn(H, device=dev); pre = torch.randn(M, H, device=dev).to(BF)
mean = x.mean(-1); rstd = 1 / torch.sqrt(x.var(-1, unbiased=False) + 1e-12); dy = torch.randn(M, H, device=dev)
xr = x.clone().requires_grad_(True); F.layer_norm(xr, (H,), g, b, 1e-12).backward(dy)
gp = pre.float()
dx16 = torch.empty(M, H, device=dev, dtype=BF); dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev)
L.check(lib.vb_layernorm_bwd(dy.data_ptr(), H, x.data_ptr(), H, g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), None, dx16.data_ptr(), H, pre.data_ptr(), H, dg.data_ptr(), db.data_ptr(), None, M, H, None, None, ST()))
torch.cuda.synchronize()
report("layernorm bwd + gelu'", dict(dx16=max(rel(dx16, xr.grad * gp) - 4e-3, 0)), 1e-3)

print("=== misc rowops")
# casts
x = torch.randn(1000003, device=dev); y = torch.empty(1000003, device=dev, dtype=BF)
L.check(lib.vb_cast_f32_to_bf16(x.data_ptr(), y.data_ptr(), x.numel(), 0, None, None, ST())); torch.cuda.synchronize()
report("cast flat", dict(e=(y.float() - x.to(BF).float()).abs().max().item()), 1e-9)
x = torch.randn(77, 3129, device=dev); y = torch.zeros(77, 3136, device=dev, dtype=BF)
L.check(lib.vb_cast2d_f32_to_bf16(x.data_ptr(), 3129, y.data_ptr(), 3136, 77, 3129, 0.5, ST())); torch.cuda.synchronize()
report("cast2d", dict(e=(y[:, :3129].float() - (x * 0.5).to(BF).float()).abs().max().item(), pad=y[:, 3129:].abs().max().item()), 1e-9)
# embeddings
B, Nt, H, V = 5, 9, 64, 50
ids = torch.randint(0, V, (B, Nt), device=dev); ids[0, 3] = 0; tt = torch.randint(0, 2, (B, Nt), device=dev); task = torch.randint(0, 20, (B,), device=dev)
word = torch.randn(V, H, device=dev); pos = torch.randn(40, H, device=dev); typ = torch.randn(2, H, device=dev); tk = torch.randn(20, H, device=dev)
for has_task in (False, True):
    No = Nt + int(has_task)
    out = torch.empty(B, No, H, device=dev)
    L.check(lib.vb_embed_text_fwd(ids.data_ptr(), tt.data_ptr(), task.data_ptr() if has_task else None, word.data_ptr(), pos.data_ptr(), typ.data_ptr(),
                                  tk.data_ptr() if has_task else None, out.data_ptr(), B, Nt, H, ST()))
    wr, pr, tr, kr = (t.clone().requires_grad_(True) for t in (word, pos, typ, tk))
    e = F.embedding(ids, wr, padding_idx=0) + F.embedding(torch.arange(Nt, device=dev)[None].expand(B, Nt), pr) + F.embedding(tt, tr)
    if has_task:
        e = torch.cat([e[:, :1], F.embedding(task[:, None], kr), e[:, 1:]], 1)
    d = torch.randn_like(e); e.backward(d)
    dw, dp, dt, dk = (torch.zeros_like(t) for t in (word, pos, typ, tk))
    L.check(lib.vb_embed_text_bwd(d.contiguous().data_ptr(), ids.data_ptr(), tt.data_ptr(), task.data_ptr() if has_task else None, dw.data_ptr(), dp.data_ptr(),
                                  dt.data_ptr(), dk.data_ptr() if has_task else None, B, Nt, H, ST()))
    torch.cuda.synchronize()
    errs = dict(out=rel(out, e), dw=rel(dw, wr.grad), dp=rel(dp, pr.grad), dt=rel(dt, tr.grad))
    if has_task: errs["dk"] = rel(dk, kr.grad)
    report(f"embed_text task{int(has_task)}", errs, 1e-5)
# loc proj
M, H = 333, 96
loc = torch.rand(M, 5, device=dev); W = torch.randn(H, 5, device=dev); b = torch.randn(H, device=dev); out = torch.empty(M, H, device=dev)
L.check(lib.vb_loc_proj_fwd(loc.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), M, H, ST()))
dy = torch.randn(M, H, device=dev); dW = torch.zeros(H, 5, device=dev); db = torch.zeros(H, device=dev)
L.check(lib.vb_loc_proj_bwd(dy.data_ptr(), loc.data_ptr(), dW.data_ptr(), db.data_ptr(), M, H, ST())); torch.cuda.synchronize()
report("loc_proj", dict(out=rel(out, loc @ W.t() + b), dW=rel(dW, dy.t() @ loc), db=rel(db, dy.sum(0))), 1e-5)
# colsum
for dt_ in (torch.float32, BF):
    X = torch.randn(2304, 776, device=dev).to(dt_); out = torch.zeros(770, device=dev)
    L.check(lib.vb_colsum(X.data_ptr(), int(dt_ == BF), 776, out.data_ptr(), 2304, 770, ST())); torch.cuda.synchronize()
    report(f"colsum {dt_}", dict(e=rel(out, X.float().sum(0)[:770])), 1e-5)
# small linear
for (M, K, N) in [(64, 1024, 1), (64, 1024, 3), (32, 2048, 2), (6400, 1024, 1)]:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev); add = torch.randn(M, device=dev)
    y = torch.empty(M, N, device=dev)
    L.check(lib.vb_small_linear_fwd(x.data_ptr(), K, W.data_ptr(), b.data_ptr(), add.data_ptr(), y.data_ptr(), M, K, N, None, ST()))
    dy = torch.randn(M, N, device=dev); dx = torch.ones(M, K, device=dev); dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    L.check(lib.vb_small_linear_bwd(dy.data_ptr(), x.data_ptr(), K, W.data_ptr(), dx.data_ptr(), K, 1, dW.data_ptr(), db.data_ptr(), M, K, N, None, ST())); torch.cuda.synchronize()
    report(f"small_linear M{M} K{K} N{N}", dict(y=rel(y, x @ W.t() + b + add[:, None]), dx=rel(dx, 1 + dy @ W), dW=rel(dW, dy.t() @ x), db=rel(db, dy.sum(0))), 1e-5)
# pooled fuse / relu / axpy / bce / mask
a = torch.randn(64, 1024, device=dev); b = torch.randn(64, 1024, device=dev); o32 = torch.empty_like(a); o16 = torch.empty(64, 1024, device=dev, dtype=BF)
L.check(lib.vb_fuse_pooled_fwd(a.data_ptr(), b.data_ptr(), o32.data_ptr(), o16.data_ptr(), a.numel(), 1, None, 0, None, None, ST()))
d = torch.randn_like(a); da = torch.ones_like(a); db = torch.ones_like(a)
L.check(lib.vb_fuse_pooled_bwd(d.data_ptr(), a.data_ptr(), b.data_ptr(), da.data_ptr(), db.data_ptr(), a.numel(), 1, None, ST())); torch.cuda.synchronize()
report("fuse_pooled", dict(o=rel(o32, a * b), da=rel(da, 1 + d * b), db=rel(db, 1 + d * a)), 1e-6)
y = torch.randn(64, 1024, device=dev); dy = torch.randn_like(y); dx16 = torch.empty(64, 1024, device=dev, dtype=BF); dx32 = torch.empty_like(y)
L.check(lib.vb_relu_bwd(dy.data_ptr(), y.data_ptr(), dx16.data_ptr(), dx32.data_ptr(), y.numel(), ST())); torch.cuda.synchronize()
report("relu_bwd", dict(e=rel(dx32, dy * (y > 0))), 1e-6)
z = torch.randn(64, 3129, device=dev) * 3; t = (torch.rand(64, 3129, device=dev) < 0.001).float() * 0.6
loss = torch.zeros(1, device=dev); dz = torch.empty_like(z); dz16 = torch.zeros(64, 3136, device=dev, dtype=BF)
L.check(lib.vb_bce_logits_loss(z.data_ptr(), t.data_ptr(), loss.data_ptr(), dz.data_ptr(), dz16.data_ptr(), 3136, 64, 3129, 1.0, ST())); torch.cuda.synchronize()
zr = z.clone().requires_grad_(True); lr = F.binary_cross_entropy_with_logits(zr, t, reduction="mean") * 3129; lr.backward()
report("bce_logits", dict(loss=abs(loss.item() - lr.item()) / lr.item(), dz=rel(dz, zr.grad), dz16=max(rel(dz16[:, :3129], zr.grad) - 4e-3, 0)), 1e-5)
m = (torch.rand(7, 13, device=dev) < 0.6).long(); o0 = torch.empty(7, 13, device=dev); o1 = torch.empty(7, 14, device=dev)
L.check(lib.vb_mask_to_additive(m.data_ptr(), o0.data_ptr(), 7, 13, 0, ST())); L.check(lib.vb_mask_to_additive(m.data_ptr(), o1.data_ptr(), 7, 13, 1, ST())); torch.cuda.synchronize()
ref0 = (1.0 - m.float()) * -10000.0
report("mask_to_additive", dict(a=(o0 - ref0).abs().max().item(), b=(o1[:, 1:] - ref0).abs().max().item(), c=o1[:, 0].abs().max().item()), 1e-9)
print("=== done")
