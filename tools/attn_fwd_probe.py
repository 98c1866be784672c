"""Forward attention on the model's four (Nq, Nk, D) shapes: tcgen05 kernel (vb_attn_tc.cu) vs the mma.sync kernel
(VB_ATTN_TC=0). Development tool; run each setting in its own process (the switch is read once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _gpu_util import attn_case
print("VB_ATTN_TC =", os.environ.get("VB_ATTN_TC", "(default)"))
for name, args in [("text self 36x36x64 B64 H12", (64, 12, 36, 36, 64, False)), ("image self 100x100x128 B64 H8", (64, 8, 100, 100, 128, False)),
                   ("cross text-q 36x100x128 B64 H8", (64, 8, 36, 100, 128, True)), ("cross image-q 100x36x128 B64 H8", (64, 8, 100, 36, 128, True))]:
    for fp16 in (False, True):
        errs, timing = attn_case(*args, iters=20, fp16=fp16)
        print(f"{name:34s} fp16={int(fp16)} O err {errs['O']:.1e} lse {errs['lse']:.1e} |{timing}")
