"""Where does the fp16-mode backward error come from? Emulates the kernel's arithmetic choices in torch."""
import os, sys, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vilbert_b200 import _lib as L
BF = torch.bfloat16
dev = "cuda"; lib = L.lib()
B, H, Nq, Nk, D = 2, 8, 100, 100, 128
g_ = torch.Generator(device=dev).manual_seed(0)
Hd = H * D
src32 = torch.randn(B * Nq, 3 * Hd, device=dev, generator=g_)
src = src32.to(torch.float16)
q, k, v = src[:, :Hd], src[:, Hd:2 * Hd], src[:, 2 * Hd:]
mask = torch.zeros(B, Nk, device=dev)
Ot = torch.zeros(B * Nq, Hd, device=dev, dtype=torch.float16); lse = torch.zeros(B, H, Nq, device=dev)
dO = torch.randn(B * Nq, Hd, device=dev, generator=g_).to(BF)
dqkv = torch.zeros(B * Nq, 3 * Hd, device=dev, dtype=BF); delta = torch.zeros(B, H, Nq, device=dev)
a = L.AttnArgs()
a.B, a.H, a.Nq, a.Nk, a.D = B, H, Nq, Nk, D
a.Q, a.ldq, a.K, a.ldk, a.V, a.ldv = q.data_ptr(), 3 * Hd, k.data_ptr(), 3 * Hd, v.data_ptr(), 3 * Hd
a.mask, a.scale = mask.data_ptr(), 1.0 / math.sqrt(D)
a.O, a.ldo, a.lse = Ot.data_ptr(), Hd, lse.data_ptr()
a.dO, a.lddo = dO.data_ptr(), Hd
a.dQ, a.lddq, a.dK, a.lddk, a.dV, a.lddv = dqkv[:, :Hd].data_ptr(), 3 * Hd, dqkv[:, Hd:2 * Hd].data_ptr(), 3 * Hd, dqkv[:, 2 * Hd:].data_ptr(), 3 * Hd
a.delta = delta.data_ptr(); a.qkv_fp16 = 1
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L.check(lib.vb_attention_fwd(C.byref(a), st)); L.check(lib.vb_attention_bwd(C.byref(a), st)); torch.cuda.synchronize()
def heads(x, N): return x.float().view(B, N, H, D).permute(0, 2, 1, 3)
def rel(x, y): return ((x - y).abs().max() / y.abs().max()).item()
mine = dict(dQ=heads(dqkv[:, :Hd], Nq), dK=heads(dqkv[:, Hd:2 * Hd], Nk), dV=heads(dqkv[:, 2 * Hd:], Nk))
dOh = heads(dO, Nq)
def exact(qq, kk, vv):
    qf, kf, vf = (heads(t, n).detach().requires_grad_(True) for t, n in ((qq, Nq), (kk, Nk), (vv, Nk)))
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    o = torch.softmax(s, -1) @ vf
    o.backward(dOh)
    return dict(dQ=qf.grad, dK=kf.grad, dV=vf.grad)
e16 = exact(q, k, v); eb = exact(q.to(BF), k.to(BF), v.to(BF))
print("kernel vs exact(fp16 inputs):", {n: f"{rel(mine[n], e16[n]):.1e}" for n in mine})
print("kernel vs exact(bf16-rounded inputs):", {n: f"{rel(mine[n], eb[n]):.1e}" for n in mine})
print("exact(bf16-rounded) vs exact(fp16):", {n: f"{rel(eb[n], e16[n]):.1e}" for n in mine})
# emulation of the kernel: S from bf16-rounded q,k; P = exp(S - lse_fwd); delta from fp16 O; P, dS rounded to bf16
qb, kb, vb = heads(q.to(BF), Nq), heads(k.to(BF), Nk), heads(v.to(BF), Nk)
S = qb @ kb.transpose(-1, -2) / math.sqrt(D)
for variant in ("lse_fwd+delta_fwd", "lse_self+delta_fwd", "lse_fwd+delta_self", "lse_self+delta_self"):
    ls = lse * math.log(2.0) if "lse_fwd" in variant else torch.logsumexp(S, -1)
    P = torch.exp(S - ls[..., None])
    dP = dOh @ vb.transpose(-1, -2)
    dl = (dOh * heads(Ot, Nq)).sum(-1) if "delta_fwd" in variant else (P * dP).sum(-1)
    dS = P * (dP - dl[..., None])
    Pr, dSr = P.to(BF).float(), dS.to(BF).float()
    em = dict(dQ=dSr @ kb / math.sqrt(D), dK=dSr.transpose(-1, -2) @ qb / math.sqrt(D), dV=Pr.transpose(-1, -2) @ dOh)
    print(f"emulation[{variant}] vs exact(bf16-rounded):", {n: f"{rel(em[n], eb[n]):.1e}" for n in mine}, "| kernel vs emulation:", {n: f"{rel(mine[n], em[n]):.1e}" for n in mine})
