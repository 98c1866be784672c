"""ORACLE — TEST INFRASTRUCTURE. Pins oracle/vilbert_oracle.py against the real reference and writes
the golden fixtures of tests/golden/. Run in the build container (needs /root/reference):

    python oracle/make_golden.py

For every case it (1) builds the reference VILBertForVLTasks (or BertForMultiModalPreTraining) from a
config, loads oracle.synth_params into it with load_state_dict, (2) runs reference and oracle on the
same synthetic inputs in fp32 eval mode, forward and backward, (3) asserts they agree to 1e-5 relative
(max-abs / max-abs), and (4) stores summaries (and full tensors for the tiny case) that
tests/test_oracle_golden.py re-checks without the reference.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader, vilbert_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
TOL = 1e-5

TINY = dict(vocab_size=120, hidden_size=64, num_hidden_layers=3, num_attention_heads=4, intermediate_size=128,
            max_position_embeddings=40, type_vocab_size=2, v_feature_size=48, v_target_size=21, v_hidden_size=96,
            v_num_hidden_layers=2, v_num_attention_heads=3, v_intermediate_size=80, bi_hidden_size=64,
            bi_num_attention_heads=2, v_biattention_id=[0, 1], t_biattention_id=[1, 2],
            hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, v_hidden_dropout_prob=0.1,
            v_attention_probs_dropout_prob=0.1)


def ref_config_json(name):
    with open(os.path.join(ref_loader.REFERENCE_ROOT, "config", name)) as f:
        return json.load(f)


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def summary(t):
    t = t.detach().double().flatten()
    n = t.numel(); k = min(16, n)
    idx = (torch.arange(k, dtype=torch.int64) * (n - 1)) // max(k - 1, 1)
    return dict(shape=None, mean=t.mean().item(), absmax=t.abs().max().item(), l2=t.norm().item(),
                samples=t[idx].tolist(), sample_idx=idx.tolist())


def run_case(ref, name, cfg_json, B, Nv, Nt, task_tokens=False, qk_scale=1.0, full=False, grads=True, seed=0):
    cfgj = dict(cfg_json)
    if task_tokens:
        cfgj["task_specific_tokens"] = True
    cfg = O.make_config(cfgj)
    rcfg = ref.BertConfig.from_dict(cfgj)
    torch.manual_seed(0)
    model = ref.VILBertForVLTasks(rcfg, num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=seed, qk_scale=qk_scale)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    assert all("decoder.weight" in m for m in missing), missing
    model.tie_weights()
    model.eval()
    inp = O.synth_inputs(cfg, B, Nv, Nt, seed=1234 + seed)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"],
            inp["image_attention_mask"], inp["co_attention_mask"], inp["task_ids"])
    # ---- reference
    ref_heads = model(*args)[:9]
    ref_bert = model.bert(*args)[:4]
    # ---- oracle
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    bert_o, heads_o = O.vilbert_for_vl_tasks(Pg, cfg, *args)
    errs = {}
    for n, a, b in zip(O.BERT_OUT_NAMES, bert_o, ref_bert):
        errs[n] = rel(a, b)
    for n, a, b in zip(O.HEAD_NAMES, heads_o, ref_heads):
        errs[n] = rel(a, b)
    grad_errs = {}
    gold = dict(name=name, config=cfgj, B=B, Nv=Nv, Nt=Nt, task_tokens=task_tokens, qk_scale=qk_scale, seed=seed,
                outputs={}, grads={})
    for n, a in list(zip(O.BERT_OUT_NAMES, bert_o)) + list(zip(O.HEAD_NAMES, heads_o)):
        s = summary(a); s["shape"] = list(a.shape); gold["outputs"][n] = s
    if grads:
        tgt = O.synth_vqa_target(B, 3129)
        # a loss touching every head so that every parameter with a grad path is exercised
        def total_loss(heads, bert):
            l = O.vqa_loss(heads[0], tgt)
            for h in heads[1:]:
                l = l + 0.1 * h.float().clamp(-50, 50).pow(2).mean()
            return l
        lo = total_loss(heads_o, bert_o)
        lo.backward()
        model.zero_grad()
        lr = total_loss(ref_heads, ref_bert)
        lr.backward()
        errs["loss"] = abs(lo.item() - lr.item()) / abs(lr.item())
        gold["loss"] = lr.item()
        ref_named = dict(model.named_parameters())
        for k, v in Pg.items():
            if k == "cls.predictions.decoder.weight":
                continue
            rg = ref_named[k].grad
            if rg is None:
                assert v.grad is None or v.grad.abs().max() == 0, k
                continue
            grad_errs[k] = rel(v.grad, rg)
            s = summary(v.grad); s["shape"] = list(v.grad.shape); gold["grads"][k] = s
    worst = max(list(errs.values()) + list(grad_errs.values()))
    print(f"{name:28s} outputs worst {max(errs.values()):.2e}  grads worst {max(grad_errs.values()) if grad_errs else 0:.2e}")
    assert worst < TOL, (name, {k: v for k, v in {**errs, **grad_errs}.items() if v >= TOL})
    gold["pin"] = dict(worst_output_rel=max(errs.values()), worst_grad_rel=max(grad_errs.values()) if grad_errs else 0.0,
                       tolerance=TOL, reference="facebookresearch/vilbert-multi-task@f22b84a vilbert/vilbert.py")
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(gold, f)
    if full:
        torch.save(dict(inputs={k: v for k, v in inp.items() if v is not None},
                        bert={n: a.detach() for n, a in zip(O.BERT_OUT_NAMES, ref_bert)},
                        heads={n: a.detach() for n, a in zip(O.HEAD_NAMES, ref_heads)},
                        grads={k: ref_named[k].grad.clone() for k in grad_errs}),
                   os.path.join(GOLD, name + ".pt"))


def run_pretraining_case(ref, name, cfg_json, B, Nv, Nt):
    cfg = O.make_config(cfg_json)
    rcfg = ref.BertConfig.from_dict(dict(cfg_json))
    model = ref.BertForMultiModalPreTraining(rcfg)
    P = O.synth_params(cfg, seed=3, with_task_heads=False)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, B, Nv, Nt, seed=77)
    g = torch.Generator().manual_seed(5)
    lm = torch.full((B, Nt), -1, dtype=torch.long)
    sel = torch.rand(B, Nt, generator=g) < 0.15; sel[:, 1] = True
    lm[sel] = torch.randint(0, cfg["vocab_size"], (int(sel.sum()),), generator=g)
    il = torch.full((B, Nv - 1), -1, dtype=torch.long); il[torch.rand(B, Nv - 1, generator=g) < 0.15] = 1; il[:, 0] = 1
    it = torch.softmax(torch.randn(B, Nv - 1, cfg["v_target_size"], generator=g), -1)
    ns = torch.randint(0, 2, (B,), generator=g)
    a = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"],
         inp["image_attention_mask"], lm, il, it, ns)
    lr = model(*a)
    lo = O.pretraining_losses(P, cfg, *a)
    errs = [abs(x.item() - y.item()) / abs(y.item()) for x, y in zip(lo, lr)]
    print(f"{name:28s} losses rel err {max(errs):.2e}  ({[round(float(x), 5) for x in lr]})")
    assert max(errs) < TOL
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(dict(name=name, config=cfg_json, B=B, Nv=Nv, Nt=Nt, losses=[float(x) for x in lr],
                       pin=dict(worst=max(errs), tolerance=TOL)), f)


def check_visual_targets(ref, cfg_json):
    """config.visual_target 1 (feature regression) and 2 (noise-contrastive estimation with sampled negatives, vilbert.py:1507-1575)
    of BertForMultiModalPreTraining: the three losses and every parameter gradient. For 2 the reference samples its negatives from
    torch's global generator; the oracle draws in the same order, so under the same manual_seed the sample is identical."""
    for vt in (1, 2):
        cfgj = dict(cfg_json, visual_target=vt, v_target_size=48, num_negative=20)
        cfg = O.make_config(cfgj)
        model = ref.BertForMultiModalPreTraining(ref.BertConfig.from_dict(dict(cfgj)))
        P = O.synth_params(cfg, seed=3, with_task_heads=False)
        model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
        B, Nv, Nt = 4, 9, 8
        inp = O.synth_inputs(cfg, B, Nv, Nt, seed=77)
        g = torch.Generator().manual_seed(5)
        lm = torch.full((B, Nt), -1, dtype=torch.long); lm[:, 1] = torch.randint(0, cfg["vocab_size"], (B,), generator=g)
        il = torch.full((B, Nv - 1), -1, dtype=torch.long); il[:, 0] = 1; il[:, 3] = 1; il[2, 7] = 1
        it = torch.randn(B, Nv - 1, 48, generator=g)
        ns = torch.randint(0, 2, (B,), generator=g)
        a = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], lm, il, it, ns)
        torch.manual_seed(4242)
        lr = model(*a)
        sum(lr).sum().backward()
        Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
        Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
        torch.manual_seed(4242)
        neg = O.nce_negative_indices(B, Nv - 1, 20) if vt == 2 else None
        lo = O.pretraining_losses(Pg, cfg, *a, neg_index=neg)
        sum(lo).backward()
        named = dict(model.named_parameters())
        worst = max(abs(x.item() - y.item()) / abs(y.item()) for x, y in zip(lo, lr))
        for k, v in Pg.items():
            if k != "cls.predictions.decoder.weight" and named[k].grad is not None:
                worst = max(worst, rel(v.grad, named[k].grad))
        print(f"{'visual_target ' + str(vt):28s} worst {worst:.2e}  ({[round(float(x.detach()), 5) for x in lr]})")
        assert worst < TOL
        with open(os.path.join(GOLD, f"tiny_visual_target_{vt}.json"), "w") as f:
            json.dump(dict(name=f"tiny_visual_target_{vt}", config=cfgj, B=B, Nv=Nv, Nt=Nt, losses=[float(x.detach()) for x in lr],
                           neg_index=neg.tolist() if neg is not None else None, seed=4242, pin=dict(worst=worst, tolerance=TOL)), f)


def check_all_encoded_layers(ref, cfg_json):
    """output_all_encoded_layers=True: per-connection-layer states + poolers on the last connection layer's output."""
    cfg = O.make_config(cfg_json)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfg_json)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 4, 11, 9, seed=1234)
    r = model.bert(inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"],
                   output_all_encoded_layers=True)
    o = O.bert_model(P, cfg, inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"],
                     inp["image_attention_mask"], output_all_encoded_layers=True)
    errs = [rel(a, b) for a, b in zip(list(o[0]) + list(o[1]) + [o[2], o[3]], list(r[0]) + list(r[1]) + [r[2], r[3]])]
    print(f"{'all_encoded_layers':28s} worst {max(errs):.2e}")
    assert len(r[0]) == len(o[0]) and max(errs) < TOL


def check_fast_mode(ref, cfg_json):
    """fast_mode=True (eval_retrieval.py): text batch 1 broadcast to the image batch at the first connection layer."""
    cfgj = dict(cfg_json, fast_mode=True)
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 5, 11, 9, seed=4321)
    txt = (inp["input_txt"][:1], inp["token_type_ids"][:1], inp["attention_mask"][:1])
    with torch.no_grad():
        r = model(txt[0], inp["input_imgs"], inp["image_loc"], txt[1], txt[2], inp["image_attention_mask"], inp["co_attention_mask"][:1])[:9]
        _, o = O.vilbert_for_vl_tasks(P, cfg, txt[0], inp["input_imgs"], inp["image_loc"], txt[1], txt[2], inp["image_attention_mask"])
    errs = [rel(a, b) for a, b in zip(o, r)]
    print(f"{'fast_mode':28s} worst {max(errs):.2e}")
    assert max(errs) < TOL and tuple(o[0].shape) == (5, 3129)
    gold = dict(name="tiny_fast_mode", config=cfgj, B=5, Nv=11, Nt=9, seed=0, input_seed=4321,
                outputs={n: dict(summary(a), shape=list(a.shape)) for n, a in zip(O.HEAD_NAMES, r)},
                pin=dict(worst_output_rel=max(errs), tolerance=TOL))
    with open(os.path.join(GOLD, "tiny_fast_mode.json"), "w") as f:
        json.dump(gold, f)


def check_in_batch_pairs(ref, cfg_json):
    """in_batch_pairs=True (vilbert.py:1008-1040) through BertModel (the reference's VILBertForVLTasks adds the unexpanded image mask
    to vision_logit and cannot run with it): the four BertModel outputs at batch b^2 and the gradients of a loss on them."""
    cfgj = dict(cfg_json, in_batch_pairs=True)
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 3, 11, 9, seed=555)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    r = model.bert(*args)[:4]
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    o = O.bert_model(Pg, cfg, *args)
    w = [torch.linspace(0.5, 1.5, x.numel()).view_as(x) for x in r]
    sum((a * b).sum() for a, b in zip(r, w)).backward()
    sum((a * b).sum() for a, b in zip(o, w)).backward()
    named = dict(model.named_parameters())
    worst = max(rel(a, b) for a, b in zip(o, r))
    for k, v in Pg.items():
        rg = named[k].grad if k != "cls.predictions.decoder.weight" else None
        if rg is not None and rg.abs().max() > 0:
            worst = max(worst, rel(v.grad, rg))
    print(f"{'in_batch_pairs':28s} worst {worst:.2e}; output batch {o[0].shape[0]}")
    assert worst < TOL and o[0].shape[0] == 9
    with open(os.path.join(GOLD, "tiny_in_batch_pairs.json"), "w") as f:
        json.dump(dict(name="tiny_in_batch_pairs", config=cfgj, B=3, Nv=11, Nt=9, seed=0, input_seed=555,
                       outputs={n: dict(summary(a), shape=list(a.shape)) for n, a in zip(O.BERT_OUT_NAMES, r)},
                       pin=dict(worst=worst, tolerance=TOL)), f)


def check_visualization(ref, cfg_json):
    """visualization=True + output_all_attention_masks=True: the attn_data dicts of every layer (vilbert.py:451-458, 610-617,
    813-821) against the oracle's attention hook (probabilities, queries, keys)."""
    cfgj = dict(cfg_json, visualization=True)
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 3, 11, 9, seed=777)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    with torch.no_grad():
        out = model(*args, None, False, True)
    at, av, ac = out[9]
    got = {}
    O.ATTN_HOOK = lambda name, p, q, k: got.__setitem__(name, (p, q, k))
    try:
        with torch.no_grad():
            O.vilbert_for_vl_tasks(P, cfg, *args)
    finally:
        O.ATTN_HOOK = None
    worst = 0.0
    for i, d in enumerate(at):
        p, q, k = got[f"bert.encoder.layer.{i}.attention.self.dropout"]
        worst = max(worst, rel(p, d["attn"]), rel(q, d["queries"]), rel(k, d["keys"]))
    for i, d in enumerate(av):
        p, q, k = got[f"bert.encoder.v_layer.{i}.attention.self.dropout"]
        worst = max(worst, rel(p, d["attn"]), rel(q, d["queries"]), rel(k, d["keys"]))
    for i, d in enumerate(ac):
        p1, q1, k1 = got[f"bert.encoder.c_layer.{i}.biattention.dropout1"]
        p2, q2, k2 = got[f"bert.encoder.c_layer.{i}.biattention.dropout2"]
        worst = max(worst, rel(p1, d["attn1"]), rel(q1, d["queries1"]), rel(k1, d["keys1"]), rel(p2, d["attn2"]), rel(q2, d["querues2"]), rel(k2, d["keys2"]))
    print(f"{'visualization':28s} worst {worst:.2e}; {len(at)} text, {len(av)} image, {len(ac)} connection layers")
    assert worst < TOL and len(at) == cfg["num_hidden_layers"] and len(ac) == len(cfg["v_biattention_id"])
    with open(os.path.join(GOLD, "tiny_visualization.json"), "w") as f:
        json.dump(dict(name="tiny_visualization", config=cfgj, B=3, Nv=11, Nt=9, seed=0, input_seed=777,
                       attn_text_last=dict(summary(at[-1]["attn"]), shape=list(at[-1]["attn"].shape)),
                       attn1_last=dict(summary(ac[-1]["attn1"]), shape=list(ac[-1]["attn1"].shape)),
                       attn2_last=dict(summary(ac[-1]["attn2"]), shape=list(ac[-1]["attn2"].shape)),
                       pin=dict(worst=worst, tolerance=TOL)), f)


def check_fixed_layers(ref, cfg_json):
    """fixed_t_layer / fixed_v_layer: the first layers of each stream run under no_grad (vilbert.py:968-1003): same outputs,
    no gradient into those layers or anything before them."""
    cfgj = dict(cfg_json, fixed_t_layer=1, fixed_v_layer=0, v_biattention_id=[0, 1], t_biattention_id=[1, 2])
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 4, 11, 9, seed=1234)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    tgt = O.synth_vqa_target(4, 3129)
    r = model(*args)[:9]
    O.vqa_loss(r[0], tgt).backward()
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    _, o = O.vilbert_for_vl_tasks(Pg, cfg, *args)
    O.vqa_loss(o[0], tgt).backward()
    named = dict(model.named_parameters())
    worst = max(rel(a, b) for a, b in zip(o, r))
    frozen = []
    for k, v in Pg.items():
        if k == "cls.predictions.decoder.weight":
            continue
        rg = named[k].grad
        if rg is None or rg.abs().max() == 0:
            assert v.grad is None or v.grad.abs().max() == 0, k
            frozen.append(k)
        else:
            worst = max(worst, rel(v.grad, rg))
    print(f"{'fixed_layers':28s} worst {worst:.2e}; {len(frozen)} tensors without gradient")
    assert worst < TOL and "bert.encoder.layer.0.output.dense.weight" in frozen and "bert.embeddings.word_embeddings.weight" in frozen
    assert "bert.encoder.layer.1.output.dense.weight" not in frozen
    with open(os.path.join(GOLD, "tiny_fixed_layers.json"), "w") as f:
        json.dump(dict(name="tiny_fixed_layers", config=cfgj, B=4, Nv=11, Nt=9, frozen=sorted(frozen), loss=float(O.vqa_loss(r[0], tgt)),
                       pin=dict(worst=worst, tolerance=TOL)), f)


def check_roberta(ref, cfg_json):
    """model="roberta" (vilbert.py:370-393, 1295-1296): RobertaEmbeddings builds position ids starting at padding_idx + 1 and hands
    them to BertEmbeddings.forward, which overwrites them with arange(seq_length) (vilbert.py:347-351) — so the embeddings are the
    BERT ones, and the task-token variant cannot run (RobertaEmbeddings.forward has no task_ids argument: BertModel passes them as
    position_ids and BertEmbeddings then indexes task_embeddings with None)."""
    cfgj = dict(cfg_json, model="roberta")
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    assert type(model.bert.embeddings).__name__ == "RobertaEmbeddings"
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 4, 11, 9, seed=1234)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    with torch.no_grad():
        r = model(*args)[:9]
        _, o = O.vilbert_for_vl_tasks(P, cfg, *args[:6])
    worst = max(rel(a, b) for a, b in zip(o, r))
    cfgt = dict(cfgj, task_specific_tokens=True)
    mt = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgt)), num_labels=1, default_gpu=False).eval()
    try:
        with torch.no_grad():
            mt(*args, task_ids=torch.zeros(4, 1, dtype=torch.long))
        task_tokens_run = True
    except Exception as e:  # noqa: BLE001 - any failure documents that the combination does not run in the reference
        task_tokens_run = False
        print(f"{'roberta + task tokens':28s} reference raises {type(e).__name__}")
    print(f"{'roberta':28s} worst {worst:.2e}")
    assert worst < TOL and not task_tokens_run
    with open(os.path.join(GOLD, "tiny_roberta.json"), "w") as f:
        json.dump(dict(name="tiny_roberta", config=cfgj, B=4, Nv=11, Nt=9, seed=0, input_seed=1234,
                       outputs={n: dict(summary(a), shape=list(a.shape)) for n, a in zip(O.HEAD_NAMES, r)},
                       pin=dict(worst_output_rel=worst, tolerance=TOL, task_tokens_run_in_reference=task_tokens_run)), f)


def check_dynamic_attention(ref, cfg_json):
    """config.dynamic_attention (BertImageSelfAttention, vilbert.py:557-586; --dynamic_attention of train_tasks.py:357-358): image
    queries / keys gated by the pooled text states. All head outputs and every parameter gradient of the VQA loss (incl. the
    dyLinear_q / dyLinear_k gates and the text stream, which now also receives gradient through the pooling)."""
    cfgj = dict(cfg_json, dynamic_attention=True)
    cfg = O.make_config(cfgj)
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    missing = [k for k in model.state_dict() if k not in P]
    assert not missing, missing
    model.load_state_dict(P, strict=False); model.tie_weights(); model.eval()
    inp = O.synth_inputs(cfg, 4, 11, 9, seed=1234)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    tgt = O.synth_vqa_target(4, 3129)
    r = model(*args)[:9]
    O.vqa_loss(r[0], tgt).backward()
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    _, o = O.vilbert_for_vl_tasks(Pg, cfg, *args[:6])
    O.vqa_loss(o[0], tgt).backward()
    named = dict(model.named_parameters())
    worst = max(rel(a, b) for a, b in zip(o, r))
    gates = {}
    for k, v in Pg.items():
        if k == "cls.predictions.decoder.weight" or named[k].grad is None:
            continue
        worst = max(worst, rel(v.grad, named[k].grad))
        if "dyLinear" in k:
            gates[k] = summary(named[k].grad)
    print(f"{'dynamic_attention':28s} worst {worst:.2e}; {len(gates)} gate tensors with gradient")
    assert worst < TOL and len(gates) == 4 * cfg["v_num_hidden_layers"]
    with open(os.path.join(GOLD, "tiny_dynamic_attention.json"), "w") as f:
        json.dump(dict(name="tiny_dynamic_attention", config=cfgj, B=4, Nv=11, Nt=9, seed=0, input_seed=1234,
                       outputs={n: dict(summary(a), shape=list(a.shape)) for n, a in zip(O.HEAD_NAMES, r)}, gate_grads=gates,
                       loss=float(O.vqa_loss(r[0], tgt).detach()), pin=dict(worst=worst, tolerance=TOL)), f)


class _MaskDropout(torch.nn.Module):
    """Stand-in for one nn.Dropout of the reference in train mode: multiplies by the engine's stateless mask for the given site
    name(s) (one name per call of the module within a forward, in call order) with the module's OWN probability."""

    def __init__(self, names, p, masks):
        super().__init__()
        self.names, self.p, self.masks, self.calls = list(names), p, masks, 0

    def forward(self, x):
        if not self.training:
            return x
        name = self.names[self.calls % len(self.names)]
        self.calls += 1
        m = self.masks.mask(name, self.p, tuple(x.shape), x.device)
        return x if m is None else x * m


def check_train_mode_dropout_placement(ref, cfg_json):
    """Train mode: WHERE every nn.Dropout of the reference acts and with WHICH probability. Every nn.Dropout module of the
    unmodified reference model is replaced, by its module path, with a mask multiplication that draws the engine's stateless mask
    for the site of that name and the module's own p (VILBertForVLTasks.dropout is called three times per forward: pooled fusion,
    sequence_output_v, sequence_output_t, vilbert.py:1677-1695). The oracle with oracle.DropMasks must then reproduce outputs and
    gradients exactly — which pins the placement, the tensor layout the mask indexes and the five distinct probabilities."""
    cfgj = dict(cfg_json, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.15, v_hidden_dropout_prob=0.2,
                v_attention_probs_dropout_prob=0.25)
    cfg = O.make_config(cfgj)
    step, head_p = 12345, 0.3
    model = ref.VILBertForVLTasks(ref.BertConfig.from_dict(dict(cfgj)), num_labels=1, dropout_prob=head_p, default_gpu=False)
    P = O.synth_params(cfg, seed=0)
    model.load_state_dict(P, strict=False); model.tie_weights(); model.train()
    masks = O.DropMasks(step, head_p=head_p)
    sites = {}
    for name, mod in list(model.named_modules()):
        if isinstance(mod, torch.nn.Dropout):
            parent = model
            parts = name.split(".")
            for q in parts[:-1]:
                parent = getattr(parent, q)
            names = ["dropout.pooled", "dropout.seq_v", "dropout.seq_t"] if name == "dropout" else [name]
            setattr(parent, parts[-1], _MaskDropout(names, mod.p, masks))
            sites[name] = mod.p
    inp = O.synth_inputs(cfg, 4, 11, 9, seed=1234)
    args = (inp["input_txt"], inp["input_imgs"], inp["image_loc"], inp["token_type_ids"], inp["attention_mask"], inp["image_attention_mask"], inp["co_attention_mask"])
    tgt = O.synth_vqa_target(4, 3129)
    r = model(*args)[:9]
    loss_r = O.vqa_loss(r[0], tgt) + r[2].sum() * 0.1 + r[6].mul(inp["image_attention_mask"].unsqueeze(2).float()).sum() * 0.01 + r[8].sum() * 0.01
    loss_r.backward()
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k != "cls.predictions.decoder.weight"}
    Pg["cls.predictions.decoder.weight"] = Pg["bert.embeddings.word_embeddings.weight"]
    _, o = O.vilbert_for_vl_tasks(Pg, cfg, *args[:6], drop=masks)
    loss_o = O.vqa_loss(o[0], tgt) + o[2].sum() * 0.1 + o[6].mul(inp["image_attention_mask"].unsqueeze(2).float()).sum() * 0.01 + o[8].sum() * 0.01
    loss_o.backward()
    named = dict(model.named_parameters())
    worst = max(rel(a, b) for a, b in zip(o, r))
    for k, v in Pg.items():
        if k != "cls.predictions.decoder.weight" and named[k].grad is not None:
            worst = max(worst, rel(v.grad, named[k].grad))
    with torch.no_grad():
        model.eval()
        r_eval = model(*args)[:9]
    differs = rel(r_eval[0], r[0])
    print(f"{'train-mode dropout sites':28s} worst {worst:.2e}; {len(sites)} nn.Dropout modules, probabilities {sorted(set(sites.values()))}; "
          f"train vs eval output differs by {differs:.2e}")
    assert worst < TOL and differs > 1e-2 and len(set(sites.values())) == 5
    with open(os.path.join(GOLD, "tiny_train_mode_dropout.json"), "w") as f:
        json.dump(dict(name="tiny_train_mode_dropout", config=cfgj, B=4, Nv=11, Nt=9, seed=0, input_seed=1234, step=step, head_p=head_p,
                       sites=sites, loss=float(loss_r.detach()),
                       outputs={n: dict(summary(a), shape=list(a.shape)) for n, a in zip(O.HEAD_NAMES, r)},
                       pin=dict(worst=worst, tolerance=TOL)), f)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_loader.load()
    torch.set_num_threads(8)
    run_case(ref, "tiny_b4", TINY, B=4, Nv=11, Nt=9, full=True)
    run_case(ref, "tiny_tasktok_odd_b3", TINY, B=3, Nv=7, Nt=12, task_tokens=True, full=True, seed=1)
    run_case(ref, "tiny_peaked_b2", TINY, B=2, Nv=37, Nt=21, qk_scale=8.0, full=True, seed=2)
    base22 = ref_config_json("bert_base_2layer_2conect.json")
    run_case(ref, "base_2layer_2conect_cfg1", base22, B=2, Nv=36, Nt=20, seed=0)           # BASELINE.json configs[0]
    base66 = ref_config_json("bert_base_6layer_6conect.json")
    run_case(ref, "base_6layer_6conect_b4", base66, B=4, Nv=100, Nt=36, seed=0)           # configs[1] shape, small B
    run_case(ref, "base_6layer_6conect_tasktok", base66, B=2, Nv=101, Nt=23, task_tokens=True, grads=False, seed=1)
    run_pretraining_case(ref, "tiny_pretraining_losses", TINY, B=4, Nv=9, Nt=8)
    check_all_encoded_layers(ref, TINY)
    check_fast_mode(ref, TINY)
    check_fixed_layers(ref, TINY)
    check_in_batch_pairs(ref, TINY)
    check_visualization(ref, TINY)
    check_roberta(ref, TINY)
    check_dynamic_attention(ref, TINY)
    check_train_mode_dropout_placement(ref, TINY)
    check_visual_targets(ref, TINY)
    print("oracle pinned against the reference on all cases; fixtures written to", GOLD)


if __name__ == "__main__":
    main()
