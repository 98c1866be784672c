"""Launches ONE instance of each representative kernel (for `ncu --set full`; keep the report small). Development tool."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from vilbert_b200 import _lib as L
from _gpu_util import gemm_case, attn_case
gemm_case(6400, 1024, 1024, bias=True, res=True, check=False)                           # F32 + residual epilogue (out-proj / FFN2)
gemm_case(2304, 3072, 768, b_mn=True, act=L.VB_ACT_DGELU, out_bf16=True, check=False)   # DGELU dgrad
gemm_case(6400, 3072, 1024, bias=True, out_bf16=True, check=False)                      # QKV bf16
gemm_case(1024, 1024, 6400, a_mn=True, b_mn=True, atomic=True, split_k=0, check=False)  # wgrad split-K
gemm_case(2304, 768, 768, bias=True, res=True, check=False)                             # small text GEMM
lib = L.lib(); dev = "cuda"; M, H = 6400, 1024
x = torch.randn(M, H, device=dev); g = torch.randn(H, device=dev); b = torch.randn(H, device=dev)
y32 = torch.empty(M, H, device=dev); y16 = torch.empty(M, H, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dy = torch.randn(M, H, device=dev); dx32 = torch.empty(M, H, device=dev); dx16 = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev); dbias = torch.zeros(H, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.vb_layernorm_fwd(x.data_ptr(), H, g.data_ptr(), b.data_ptr(), 1e-12, y32.data_ptr(), y16.data_ptr(), H, mean.data_ptr(), rstd.data_ptr(), M, H, None, 0, None, None, st)
lib.vb_layernorm_bwd(dy.data_ptr(), H, x.data_ptr(), H, g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx32.data_ptr(), dx16.data_ptr(), H, None, 0,
                     dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), M, H, None, None, 0, None, st)
torch.cuda.synchronize()
# attention last (the check inside attn_case also launches torch kernels, which the -k filter ignores)
attn_case(64, 8, 100, 100, 128, False)
torch.cuda.synchronize()
