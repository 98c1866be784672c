"""GPU probe for vb_gemm_bf16: correctness spot checks + kernel-time throughput on the model's GEMM shapes. Development tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from vilbert_b200 import _lib as L
from _gpu_util import gemm_case

print("=== gemm probe", torch.cuda.get_device_name(0), time.ctime(), flush=True)
for kw in [dict(bias=True, res=True), dict(bias=True, act=L.VB_ACT_GELU, out_bf16=True), dict(act=L.VB_ACT_DGELU, out_bf16=True), dict(atomic=True, split_k=0)]:
    for shp in [(2304, 768, 768), (300, 200, 136), (333, 1601, 1024)]:
        err, _ = gemm_case(*shp, **kw)
        print(f"{'PASS' if err < 2e-3 else 'FAIL'} {kw} {shp} err={err:.2e}", flush=True)
print("=== kernel-time throughput (launches queued ahead)")
for (M, N, K, kw, name) in [
    (2304, 2304, 768, dict(bias=True, out_bf16=True), "text QKV"),
    (2304, 768, 768, dict(bias=True, res=True), "text out-proj"),
    (2304, 3072, 768, dict(bias=True, act=L.VB_ACT_GELU, out_bf16=True), "text FFN1"),
    (2304, 768, 3072, dict(bias=True, res=True), "text FFN2"),
    (6400, 3072, 1024, dict(bias=True, out_bf16=True), "image QKV"),
    (6400, 1024, 1024, dict(bias=True, res=True), "image out-proj"),
    (6400, 1024, 1024, dict(bias=True, act=L.VB_ACT_GELU, out_bf16=True), "image FFN1"),
    (6400, 1024, 2048, dict(bias=True, res=True), "image embed"),
    (2304, 30522, 768, dict(bias=True), "LM decoder"),
    (8192, 8192, 8192, dict(out_bf16=True), "square 8192"),
    (2304, 768, 768, dict(b_mn=True, res=True), "text dgrad out-proj"),
    (2304, 768, 3072, dict(b_mn=True, act=L.VB_ACT_DGELU, out_bf16=True), "text dgrad FFN2 (K=768->N=3072)"),
    (6400, 1024, 3072, dict(b_mn=True, res=True), "image dgrad QKV"),
    (768, 768, 2304, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "text wgrad out-proj"),
    (3072, 768, 2304, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "text wgrad FFN1"),
    (1024, 1024, 6400, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "image wgrad"),
    (3072, 1024, 6400, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "image wgrad qkv"),
]:
    line = f"{name:34s} M{M} N{N} K{K}:"
    for bn in (128, 256):
        _, ms = gemm_case(M, N, K, block_n=bn, check=False, iters=20, **kw)
        line += f"  bn{bn} {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF"
    print(line, flush=True)
print("=== done")
