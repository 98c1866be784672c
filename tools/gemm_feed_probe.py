"""Operand-feed probe for the tcgen05 GEMM main loop (development tool).

One output tile per CTA and a long K (128 k-blocks) so the main loop dominates; the per-CTA clock64 timeline gives
cycles per k-block as a function of (tile width, number of CTAs running, how the CTAs share operands). Tells whether
the 128-wide main loop is bound by L2 bytes (scales with CTAs), by per-TMA-request cost (constant), or by the MMA.
Also times torch.matmul (cuBLAS) on the model's GEMM shapes as a yardstick for what the sizes allow.
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vilbert_b200 import _lib as L
lib = L.lib(); dev = "cuda"; BF = torch.bfloat16


def feed(mt, nt, K, bn, a_mn=False, b_mn=False, cl=1):
    M, N = 128 * mt, bn * nt
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).to(BF)
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(BF)
    out = torch.empty(M, N, device=dev, dtype=BF)
    dbg = torch.zeros(148 * 10, dtype=torch.int64, device=dev)
    g = L.GemmArgs(); g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = A.data_ptr(), (M if a_mn else K), int(a_mn)
    g.B, g.ldb, g.b_mn_major = B.data_ptr(), (N if b_mn else K), int(b_mn)
    g.alpha, g.out_bf16, g.ld_out_bf16, g.split_k, g.block_n = 1.0, out.data_ptr(), N, 1, bn
    g.cluster_m = cl
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2): L.check(lib.vb_gemm_bf16(C.byref(g), st))
    torch.cuda.synchronize()
    g.dbg_timeline = dbg.data_ptr()
    L.check(lib.vb_gemm_bf16(C.byref(g), st)); torch.cuda.synchronize()
    full = dbg.view(148, 10).cpu(); full = full[full[:, 0] != 0]
    lead = full[full[:, 4] != 0]                       # in pair mode only the leader CTA issues MMAs
    loop = (lead[:, 4] - lead[:, 3]).float()          # first full_bar -> last MMA committed
    kb = K // 64
    print(f"  tiles {mt:3d}x{nt}  bn{bn} a_mn{int(a_mn)} b_mn{int(b_mn)} cluster{cl}: {len(full):3d} CTAs, cycles per k-block median {loop.median().item()/kb:6.0f} "
          f"max {loop.max().item()/kb:6.0f}  ({(16384 + bn * 128 // cl) / (loop.median().item()/kb):5.1f} B/clk/SM, "
          f"{(16384 + bn * 128 // cl) * len(full) / (loop.median().item()/kb):7.0f} B/clk chip)", flush=True)


print("=== main-loop feed: one tile per CTA, K = 8192")
for bn in (128, 256):
    for (mt, nt) in [(1, 1), (8, 1), (37, 1), (74, 1), (148, 1), (37, 4), (74, 2), (18, 8), (4, 37)]:
        if mt * nt <= 148:
            feed(mt, nt, 8192, bn)
print("=== CTA pairs, tcgen05 cta_group::2 (cluster_m=2): tiles are 256 x bn per pair")
for bn in (128, 256):
    for (mt, nt) in [(2, 1), (8, 1), (74, 1), (148, 1), (36, 4), (74, 2), (18, 8), (4, 37)]:
        feed(mt, nt, 8192, bn, cl=2)
    feed(36, 4, 8192, bn, a_mn=True, b_mn=True, cl=2)
    feed(36, 4, 8192, bn, b_mn=True, cl=2)
print("=== MN-major operands (wgrad form)")
for bn in (128, 256):
    feed(37, 4, 8192, bn, a_mn=True, b_mn=True)
    feed(37, 4, 8192, bn, b_mn=True)

print("=== cuBLAS yardstick (torch.matmul bf16, launches queued, CUDA events)")
def cublas(M, N, K, name, tA=False, tB=True):
    a = torch.randn(K, M, device=dev).to(BF).t() if tA else torch.randn(M, K, device=dev).to(BF)
    b = torch.randn(N, K, device=dev).to(BF).t() if tB else torch.randn(K, N, device=dev).to(BF)
    o = torch.empty(M, N, device=dev, dtype=BF)
    for _ in range(5): torch.matmul(a, b, out=o)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30): torch.matmul(a, b, out=o)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(f"  {name:28s} M{M} N{N} K{K}: {us:7.1f} us {2.0*M*N*K/us/1e6:7.1f} TF", flush=True)
for (M, N, K, name, tA, tB) in [
    (2304, 2304, 768, "text QKV", False, True), (2304, 768, 768, "text out-proj", False, True),
    (2304, 3072, 768, "text FFN1", False, True), (2304, 768, 3072, "text FFN2", False, True),
    (6400, 3072, 1024, "image QKV", False, True), (6400, 1024, 1024, "image out-proj", False, True),
    (6400, 1024, 3072, "image dgrad QKV", False, False), (1024, 1024, 6400, "image wgrad", True, False),
    (3072, 1024, 6400, "image wgrad qkv", True, False), (768, 768, 2304, "text wgrad", True, False),
    (8192, 8192, 8192, "square 8192", False, True)]:
    cublas(M, N, K, name, tA, tB)
print("=== done")
