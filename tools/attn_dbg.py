import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _gpu_util import attn_case
print("TWO_KERNELS", os.environ.get("VB_ATTN_BWD_TWO_KERNELS"))
for args in [(2, 8, 257, 306, 128, True), (2, 8, 100, 100, 128, False), (2, 8, 129, 140, 128, True), (2, 4, 257, 306, 64, True)]:
    for fp16 in (False, True):
        errs, _ = attn_case(*args, fp16=fp16)
        print(args, "fp16", fp16, {k: f"{v:.1e}" for k, v in errs.items()})
