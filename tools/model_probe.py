"""GPU probe: full model (engine plan) vs the oracle, outputs and gradients, per tensor. Development tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vilbert_b200.config import BertConfig
from vilbert_b200.engine import Engine
from oracle import vilbert_oracle as O

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20)).item()


def run(name, cfgj, B, Nv, Nt, seed=0, qk_scale=1.0, grads=True, verbose=False, which=None):
    cfg = O.make_config(cfgj)
    P = O.synth_params(cfg, seed=seed, device=dev, qk_scale=qk_scale)
    inp = O.synth_inputs(cfg, B, Nv, Nt, seed=1234 + seed, device=dev)
    eng = Engine(BertConfig.from_dict(cfgj), dev)
    for k in eng.ps.entries:
        eng.ps.p(k).copy_(P[k])
    eng.refresh_weights()
    names = O.HEAD_NAMES if which is None else which
    plan = eng.plan(B, Nt, Nv, grad_outputs=names if grads else ())
    plan.load_inputs(inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["task_ids"])
    plan.run_forward()
    torch.cuda.synchronize()
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"], inp["task_ids"])
    bert_o, heads_o = O.vilbert_for_vl_tasks(Pg, cfg, *args)
    errs = {}
    for n, r in list(zip(O.BERT_OUT_NAMES, bert_o)) + list(zip(O.HEAD_NAMES, heads_o)):
        errs[n] = rel(plan.outputs[n].reshape(r.shape), r)
    print(f"--- {name}: outputs  worst={max(errs.values()):.2e}")
    for k, v in errs.items():
        print(f"     {k:28s} {v:.2e}{'   <-- FAIL' if not v < 1e-2 else ''}")
    if not grads:
        return
    tgt = O.synth_vqa_target(B, 3129, device=dev)

    def total_loss(heads):
        l = 0
        for n, h in zip(O.HEAD_NAMES, heads):
            if n not in names: continue
            l = l + (O.vqa_loss(h, tgt) if n == "vil_prediction" else 0.1 * h.float().clamp(-50, 50).pow(2).mean())
        return l
    lo = total_loss(heads_o); lo.backward()
    mine = [plan.outputs[n].detach().clone().requires_grad_(True) for n in O.HEAD_NAMES]
    lm = total_loss(mine); lm.backward()
    eng.zero_grad()
    for n, t in zip(O.HEAD_NAMES, mine):
        if n in names:
            plan.gout[n].copy_(t.grad.reshape(plan.gout[n].shape))
    plan.run_backward()
    torch.cuda.synchronize()
    gerr = {}
    for k in eng.ps.entries:
        rg = Pg[k].grad
        mg = eng.ps.g(k)
        if rg is None:
            if mg.abs().max().item() != 0: gerr[k] = float("inf")
            continue
        gerr[k] = rel(mg, rg)
    bad = {k: v for k, v in gerr.items() if not v < 1e-2}
    print(f"--- {name}: loss oracle {lo.item():.5f} mine {lm.item():.5f}; grads worst={max(gerr.values()):.2e} median={sorted(gerr.values())[len(gerr)//2]:.2e} n_bad={len(bad)}/{len(gerr)}")
    items = sorted(gerr.items(), key=lambda kv: -kv[1])
    for k, v in (items if verbose else items[:25]):
        print(f"     {k:70s} {v:.2e}{'   <-- FAIL' if not v < 1e-2 else ''}")


if __name__ == "__main__":
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    tiny = json.load(open(os.path.join(gold, "tiny_b4.json")))["config"]
    base22 = json.load(open(os.path.join(gold, "base_2layer_2conect_cfg1.json")))["config"]
    base66 = json.load(open(os.path.join(gold, "base_6layer_6conect_b4.json")))["config"]
    t0 = time.time()
    run("tiny B4", tiny, 4, 11, 9, verbose=True)
    run("tiny tasktok odd B3", dict(tiny, task_specific_tokens=True), 3, 7, 12, seed=1)
    run("tiny peaked", tiny, 2, 37, 21, seed=2, qk_scale=8.0)
    run("tiny vqa-only", tiny, 4, 11, 9, which=("vil_prediction",))
    run("base22 cfg1 B2", base22, 2, 36, 20)
    run("base66 B8", base66, 8, 100, 36)
    run("base66 B32 peaked tasktok", dict(base66, task_specific_tokens=True), 32, 101, 23, seed=3, qk_scale=4.0)
    print("elapsed", time.time() - t0)
