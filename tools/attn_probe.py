"""Attention forward / backward kernel times on the model's four shapes (development tool): single-pass backward vs the
two-kernel backward (VB_ATTN_BWD_TWO_KERNELS=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _gpu_util import attn_case
print("two-kernel backward" if os.environ.get("VB_ATTN_BWD_TWO_KERNELS") else "single-pass backward")
for args, nm in (((64, 12, 36, 36, 64, False), "text self"), ((64, 8, 100, 100, 128, False), "image self"),
                 ((64, 8, 36, 100, 128, True), "text q x image kv"), ((64, 8, 100, 36, 128, True), "image q x text kv")):
    errs, timing = attn_case(*args, iters=30)
    print(f"  {nm:20s} B{args[0]} H{args[1]} Nq{args[2]} Nk{args[3]} D{args[4]}: {timing}   max err {max(errs.values()):.2e}", flush=True)
