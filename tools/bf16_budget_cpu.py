"""CPU experiment: how far are outputs AND gradients from fp32 when every matmul operand (forward and
backward) is rounded to bf16 with fp32 accumulation — i.e. the inherent error of the engine's "bf16 mode"
independent of any kernel bug. Patches F.linear / torch.matmul inside the oracle with an autograd Function
whose backward also rounds its operands (dy, x, W)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import vilbert_oracle as O

r = lambda t: t.to(torch.bfloat16).to(torch.float32)


class LinBF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w); ctx.hb = b is not None
        y = r(x) @ r(w).t()
        return y + b if b is not None else y
    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyr = r(dy)
        dx = dyr @ r(w)
        dw = dyr.reshape(-1, dyr.shape[-1]).t() @ r(x).reshape(-1, x.shape[-1])
        return dx, dw, (dy.reshape(-1, dy.shape[-1]).sum(0) if ctx.hb else None)


class MMBF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return r(a) @ r(b)
    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        return r(dy) @ r(b).transpose(-1, -2), r(a).transpose(-1, -2) @ r(dy)


def run(cfgj, B, Nv, Nt, seed=0, qk_scale=1.0):
    cfg = O.make_config(cfgj)
    P = O.synth_params(cfg, seed=seed, qk_scale=qk_scale)
    inp = O.synth_inputs(cfg, B, Nv, Nt, seed=1234 + seed)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"], inp["task_ids"])
    tgt = O.synth_vqa_target(B, 3129)
    res = []
    for mode in ("fp32", "bf16"):
        Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
        Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
        of, om = F.linear, torch.matmul
        if mode == "bf16":
            F.linear = lambda x, w, b=None: LinBF.apply(x, w, b)
            torch.matmul = lambda a, b: MMBF.apply(a, b)
        try:
            bert_o, heads_o = O.vilbert_for_vl_tasks(Pg, cfg, *args)
            l = O.vqa_loss(heads_o[0], tgt)
            l.backward()
        finally:
            F.linear, torch.matmul = of, om
        res.append((bert_o, heads_o, {k: v.grad for k, v in Pg.items() if v.grad is not None}))
    (b0, h0, g0), (b1, h1, g1) = res
    rel = lambda a, b: ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
    rl2 = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
    print("outputs max-rel:", {n: f"{rel(a, b):.1e}" for n, a, b in list(zip(O.BERT_OUT_NAMES, b1, b0)) + list(zip(O.HEAD_NAMES, h1, h0))})
    gmax = max(v.abs().max().item() for v in g0.values())
    e = sorted(((rel(g1[k], g0[k]), rl2(g1[k], g0[k]), k) for k in g0 if g0[k].abs().max().item() > 1e-4 * gmax), reverse=True)
    print(f"grads (VQA loss only): n={len(e)} worst max-rel {e[0][0]:.2e} ({e[0][2]}), median max-rel {e[len(e)//2][0]:.2e}, median rel-L2 {sorted(x[1] for x in e)[len(e)//2]:.2e}, worst rel-L2 {max(x[1] for x in e):.2e}")


gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
torch.set_num_threads(8)
print("tiny B4"); run(json.load(open(os.path.join(gold, "tiny_b4.json")))["config"], 4, 11, 9)
print("base22 B2"); run(json.load(open(os.path.join(gold, "base_2layer_2conect_cfg1.json")))["config"], 2, 36, 20)
print("base66 B8"); run(json.load(open(os.path.join(gold, "base_6layer_6conect_b4.json")))["config"], 8, 100, 36)
