"""GPU probe for vb_gemm_bf16: correctness vs torch on every operand-major combination, epilogue
variant and edge shape, a descriptor sweep for MN-major operands if the default encoding is wrong,
and a first throughput reading. Development tool (run under gpurun); the pytest parity tests are
tests/test_gemm_gpu.py."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vilbert_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
LOG = open(os.path.join("gpurun_out", "gemm_probe.log"), "a") if os.path.isdir("gpurun_out") else sys.stdout


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    if LOG is not sys.stdout:
        LOG.write(s + "\n"); LOG.flush()


def run(M, N, K, a_mn=False, b_mn=False, bias=False, res=False, act=0, out_bf16=False, atomic=False,
        split_k=1, block_n=0, dbg=(0, 0, 0, 0), alpha=1.0, check=True, iters=0):
    # logical A[M,K], B[N,K]
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    pad8 = lambda x: (x + 7) // 8 * 8
    if a_mn:
        A_st = torch.zeros(K, pad8(M), device=dev, dtype=torch.bfloat16); A_st[:, :M] = A.t(); lda = pad8(M)
    else:
        A_st = torch.zeros(M, pad8(K), device=dev, dtype=torch.bfloat16); A_st[:, :K] = A; lda = pad8(K)
    if b_mn:
        B_st = torch.zeros(K, pad8(N), device=dev, dtype=torch.bfloat16); B_st[:, :N] = B.t(); ldb = pad8(N)
    else:
        B_st = torch.zeros(N, pad8(K), device=dev, dtype=torch.bfloat16); B_st[:, :K] = B; ldb = pad8(K)
    bias_t = torch.randn(N, device=dev) if bias else None
    res_t = torch.randn(M, N, device=dev) if res else None
    aux_t = torch.randn(M, N, device=dev).bfloat16() if act == L.VB_ACT_DGELU else None
    out32 = torch.full((M, N), 0.0 if atomic else float("nan"), device=dev)
    out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if out_bf16 else None
    pre16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if (act == L.VB_ACT_GELU and out_bf16 and N % 8 == 0) else None
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = A_st.data_ptr(), lda, int(a_mn)
    g.B, g.ldb, g.b_mn_major = B_st.data_ptr(), ldb, int(b_mn)
    g.alpha = alpha
    g.bias = bias_t.data_ptr() if bias else None
    g.residual, g.ld_res = (res_t.data_ptr(), N) if res else (None, 0)
    g.aux, g.ld_aux = (aux_t.data_ptr(), N) if aux_t is not None else (None, 0)
    g.act = act
    g.out_f32, g.ld_out_f32 = out32.data_ptr(), N
    g.out_bf16, g.ld_out_bf16 = (out16.data_ptr(), N) if out_bf16 and not atomic else (None, 0)
    g.out_pre, g.ld_out_pre = (pre16.data_ptr(), N) if pre16 is not None else (None, 0)
    g.atomic_out, g.split_k, g.block_n, g.max_ctas = int(atomic), split_k, block_n, 0
    g.dbg_lbo_a, g.dbg_sbo_a, g.dbg_lbo_b, g.dbg_sbo_b = dbg
    st = lib.vb_gemm_bf16(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(st, "vb_gemm_bf16")
    torch.cuda.synchronize()
    err = None
    if check:
        ref = alpha * (A.float() @ B.float().t())
        if bias: ref = ref + bias_t
        if act == L.VB_ACT_GELU:
            pre_ref = ref.clone(); ref = ref * 0.5 * (1 + torch.erf(ref / 2 ** 0.5))
        elif act == L.VB_ACT_RELU: ref = ref.clamp_min(0)
        elif act == L.VB_ACT_DGELU:
            x = aux_t.float()
            ref = ref * (0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5)
        if res: ref = ref + res_t
        scale = ref.abs().max().item() + 1e-9
        err = ((out32 - ref).abs().max() / scale).item()
        if out16 is not None and not atomic:
            err = max(err, ((out16.float() - ref).abs().max() / scale).item() - 4e-3)
        if pre16 is not None:
            err = max(err, ((pre16.float() - pre_ref).abs().max() / (pre_ref.abs().max() + 1e-9)).item() - 4e-3)
        if err != err: err = float("inf")
    ms = None
    if iters:
        if atomic: out32.zero_()
        for _ in range(3): lib.vb_gemm_bf16(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters): lib.vb_gemm_bf16(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    return err, ms


def report(name, err, ms=None, flops=None, tol=2e-3):
    ok = err is None or err < tol
    extra = ""
    if ms: extra = f"  {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s"
    log(f"{'PASS' if ok else 'FAIL'}  {name:58s} err={err if err is None else f'{err:.2e}'}{extra}")
    return ok


log("=== gemm probe", torch.cuda.get_device_name(0), time.ctime())
ok_tn = True
for (M, N, K, bn) in [(128, 128, 64, 128), (128, 256, 64, 256), (256, 256, 128, 0), (384, 512, 256, 0),
                      (2304, 768, 768, 0), (2304, 2304, 768, 0), (6400, 1024, 1024, 128), (6400, 3072, 1024, 256),
                      (100, 72, 40, 0), (333, 1601, 1024, 0), (130, 30522, 768, 0), (64, 1024, 768, 0)]:
    err, _ = run(M, N, K, block_n=bn)
    ok_tn &= report(f"TN plain M{M} N{N} K{K} bn{bn}", err)
for kw in [dict(bias=True), dict(bias=True, act=L.VB_ACT_GELU, out_bf16=True), dict(bias=True, res=True),
           dict(act=L.VB_ACT_DGELU, out_bf16=True), dict(bias=True, act=L.VB_ACT_RELU, out_bf16=True),
           dict(atomic=True, split_k=3), dict(atomic=True, split_k=0), dict(alpha=0.125, res=True)]:
    for (M, N, K) in [(2304, 768, 768), (300, 200, 136)]:
        err, _ = run(M, N, K, **kw)
        report(f"TN {kw} M{M} N{N} K{K}", err)

combos = [(False, True, "dgrad  A k-major, B mn-major"), (True, True, "wgrad  A mn-major, B mn-major"), (True, False, "A mn-major, B k-major")]
for a_mn, b_mn, name in combos:
    good = True
    for (M, N, K, bn) in [(128, 128, 64, 128), (256, 256, 128, 256), (768, 768, 2304, 0), (1000, 520, 200, 0), (3072, 768, 6400, 0)]:
        err, _ = run(M, N, K, a_mn=a_mn, b_mn=b_mn, block_n=bn)
        good &= report(f"{name} M{M} N{N} K{K} bn{bn}", err)
    if not good:
        log("   default MN-major descriptor failed; sweeping (lbo, sbo) candidates")
        cands = [(8192, 1024), (1024, 8192), (1024, 2048), (2048, 1024), (8192, 2048), (128, 1024), (1024, 128)]
        for ca in (cands if a_mn else [(0, 0)]):
            for cb in (cands if b_mn else [(0, 0)]):
                try:
                    err, _ = run(256, 256, 128, a_mn=a_mn, b_mn=b_mn, block_n=256, dbg=(ca[0], ca[1], cb[0], cb[1]))
                except Exception as e:  # noqa
                    err = float("inf"); log("   exception", e)
                if err < 2e-3:
                    log(f"   CANDIDATE OK a(lbo,sbo)={ca} b(lbo,sbo)={cb} err={err:.2e}")

log("=== throughput (single GEMM, back-to-back launches, includes host tensor-map encode)")
for (M, N, K, kw, name) in [
    (2304, 2304, 768, dict(bias=True, out_bf16=True), "text QKV"),
    (2304, 768, 768, dict(bias=True, res=True), "text out-proj"),
    (2304, 3072, 768, dict(bias=True, act=L.VB_ACT_GELU, out_bf16=True), "text FFN1"),
    (2304, 768, 3072, dict(bias=True, res=True), "text FFN2"),
    (6400, 3072, 1024, dict(bias=True, out_bf16=True), "image QKV"),
    (6400, 1024, 1024, dict(bias=True, res=True), "image out-proj"),
    (6400, 1024, 2048, dict(bias=True), "image embed"),
    (2304, 30522, 768, dict(bias=True), "LM decoder"),
    (8192, 8192, 8192, dict(out_bf16=True), "square 8192"),
    (2304, 768, 768, dict(b_mn=True, res=True), "text dgrad out-proj"),
    (768, 768, 2304, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "text wgrad out-proj"),
    (1024, 1024, 6400, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "image wgrad"),
    (3072, 1024, 6400, dict(a_mn=True, b_mn=True, atomic=True, split_k=0), "image wgrad qkv"),
]:
    for bn in (128, 256):
        try:
            err, ms = run(M, N, K, block_n=bn, check=False, iters=20, **kw)
            report(f"{name} M{M} N{N} K{K} bn{bn}", None, ms, 2.0 * M * N * K)
        except Exception as e:
            log("   exception", name, e)
log("=== done")
