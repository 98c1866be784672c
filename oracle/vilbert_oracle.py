"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

fp32 restatement (plain torch, device-agnostic, autograd-capable) of the ViLBERT two-stream hot path
of facebookresearch/vilbert-multi-task @ f22b84a, function by function, keyed on the reference's own
state_dict names. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product (vilbert-multi-task_b200/) never does.

Pinning: the reference ships no golden vectors or model tests for this path, so
this restatement is pinned against the reference ITSELF: oracle/make_golden.py imports
/root/reference/vilbert/vilbert.py in the build container, checks that every output and gradient of
this file equals the reference's (fp32, max-abs-diff <= 1e-5 relative) on every config of
tests/golden/, and writes the fixtures that tests/test_oracle_golden.py re-checks everywhere
(including the GPU box, where /root/reference does not exist).

Each function cites the reference lines (vilbert/vilbert.py unless noted) it follows.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- config defaults
# BertConfig.__init__ defaults (vilbert.py:145-185); from_dict starts from these (:263-268).
CONFIG_DEFAULTS = dict(
    vocab_size=-1, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
    hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
    type_vocab_size=2, initializer_range=0.02, v_feature_size=2048, v_target_size=1601, v_hidden_size=768,
    v_num_hidden_layers=3, v_num_attention_heads=12, v_intermediate_size=3072, bi_hidden_size=1024,
    bi_num_attention_heads=16, v_attention_probs_dropout_prob=0.1, v_hidden_act="gelu", v_hidden_dropout_prob=0.1,
    v_initializer_range=0.2, v_biattention_id=[0, 1], t_biattention_id=[10, 11], visual_target=0, fast_mode=False,
    fixed_v_layer=0, fixed_t_layer=0, in_batch_pairs=False, fusion_method="mul", dynamic_attention=False,
    with_coattention=True, objective=0, num_negative=128, model="bert", task_specific_tokens=False,
    visualization=False,
)


def make_config(json_dict):
    cfg = dict(CONFIG_DEFAULTS)
    cfg.update(json_dict)
    return cfg


# --------------------------------------------------------------------------- dropout masks (train mode)
class DropMasks:
    """Bit-exact Python restatement of the engine's stateless dropout RNG (include/vilbert_b200.h: vb_dropout):
    keep(i) = hash32(i ^ hash32(site + step * 0x9E3779B9)) >= p * 2^32, hash32 = lowbias32, site = crc32(layer name),
    i = row-major element index of the tensor nn.Dropout acts on. Passing `drop=DropMasks(step)` to the oracle functions
    turns every nn.Dropout of the reference (train mode) into a multiplication with the engine's mask, so train-mode
    parity can be checked exactly like eval-mode parity. drop=None is eval mode."""

    M = 0xFFFFFFFF

    def __init__(self, step, head_p=0.1):
        self.step = int(step) & self.M
        self.head_p = head_p

    @classmethod
    def _hash32(cls, x):
        x = x & cls.M
        x = x ^ (x >> 16); x = (x * 0x7feb352d) & cls.M
        x = x ^ (x >> 15); x = (x * 0x846ca68b) & cls.M
        x = x ^ (x >> 16)
        return x

    def mask(self, name, p, shape, device):
        import zlib
        import numpy as np
        p32 = float(np.float32(p))
        if p32 <= 0.0:
            return None
        site = zlib.crc32(name.encode()) & self.M
        seed = int(self._hash32(torch.tensor([(site + self.step * 0x9E3779B9) & self.M], dtype=torch.int64))[0])
        n = 1
        for d in shape:
            n *= d
        idx = torch.arange(n, dtype=torch.int64, device=device) & self.M
        keep = self._hash32(idx ^ seed) >= int(p32 * 4294967296.0)
        scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        return keep.to(torch.float32).view(shape) * scale


def _drop(x, drop, name, p):
    if drop is None:
        return x
    m = drop.mask(name, p, tuple(x.shape), x.device)
    return x if m is None else x * m


# --------------------------------------------------------------------------- primitives
def gelu(x):
    """vilbert.py:111-117 — exact erf GELU."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-12):
    """vilbert.py:304-317 — biased variance, eps inside the sqrt, affine after."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def linear(P, name, x):
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


ATTN_HOOK = None   # optional callable(layer dropout name, probs [B,H,Nq,Nk], queries [B,H,Nq,D], keys [B,H,Nk,D]) used by the visualization tests


def _heads(x, n_heads):
    """transpose_for_scores, vilbert.py:416-422 / :732-739."""
    B, N, H = x.shape
    return x.view(B, N, n_heads, H // n_heads).permute(0, 2, 1, 3)


def attention(q, k, v, add_mask, n_heads, drop=None, drop_name=None, drop_p=0.0):
    """QK^T / sqrt(d) + mask -> softmax -> PV -> merge heads (vilbert.py:434-449, :593-608, :771-809).
    add_mask is the additive [B,1,1,Nk] mask (0 / -10000). Dropout on the probabilities is identity
    here (eval mode / p=0: the parity protocol of SURVEY.md §8c)."""
    q, k, v = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    s = s + add_mask
    p = torch.softmax(s, dim=-1)
    if ATTN_HOOK is not None:       # config.visualization: attn_data of the reference (probabilities before dropout is identity in eval)
        ATTN_HOOK(drop_name, p, q, k)
    p = _drop(p, drop, drop_name, drop_p)
    ctx = torch.matmul(p, v).permute(0, 2, 1, 3).contiguous()
    return ctx.view(ctx.shape[0], ctx.shape[1], -1)


# --------------------------------------------------------------------------- encoder blocks
def text_layer(P, pre, cfg, h, mask, drop=None):
    """BertLayer.forward, vilbert.py:527-533 (= BertSelfAttention :424-460, BertSelfOutput :470-474,
    BertIntermediate :500-503, BertOutput :513-517)."""
    a = pre + ".attention"
    ctx = attention(linear(P, a + ".self.query", h), linear(P, a + ".self.key", h), linear(P, a + ".self.value", h),
                    mask, cfg["num_attention_heads"], drop, a + ".self.dropout", cfg["attention_probs_dropout_prob"])
    hp = cfg["hidden_dropout_prob"]
    h1 = layer_norm(_drop(linear(P, a + ".output.dense", ctx), drop, a + ".output.dropout", hp) + h,
                    P[a + ".output.LayerNorm.weight"], P[a + ".output.LayerNorm.bias"])
    f = gelu(linear(P, pre + ".intermediate.dense", h1))
    return layer_norm(_drop(linear(P, pre + ".output.dense", f), drop, pre + ".output.dropout", hp) + h1,
                      P[pre + ".output.LayerNorm.weight"], P[pre + ".output.LayerNorm.bias"])


def image_layer(P, pre, cfg, h, mask, drop=None, txt=None, txt_mask2=None):
    """BertImageLayer.forward, vilbert.py:688-694 (self-attn :571-619, :629-633, :661-664, :674-678). Same block on the visual
    stream. config.dynamic_attention (:577-586): queries and keys are scaled per (sample, channel) by 1 + sigmoid(dyLinear(masked
    mean of the current text states)); txt_mask2 is attention_mask.unsqueeze(2) (:1344)."""
    a = pre + ".attention"
    q, k = linear(P, a + ".self.query", h), linear(P, a + ".self.key", h)
    if cfg.get("dynamic_attention"):
        pool = (txt * txt_mask2).sum(1) / txt_mask2.sum(1)
        q = q * (1 + torch.sigmoid(linear(P, a + ".self.dyLinear_q", pool))).unsqueeze(1)
        k = k * (1 + torch.sigmoid(linear(P, a + ".self.dyLinear_k", pool))).unsqueeze(1)
    ctx = attention(q, k, linear(P, a + ".self.value", h),
                    mask, cfg["v_num_attention_heads"], drop, a + ".self.dropout", cfg["v_attention_probs_dropout_prob"])
    hp = cfg["v_hidden_dropout_prob"]
    h1 = layer_norm(_drop(linear(P, a + ".output.dense", ctx), drop, a + ".output.dropout", hp) + h,
                    P[a + ".output.LayerNorm.weight"], P[a + ".output.LayerNorm.bias"])
    f = gelu(linear(P, pre + ".intermediate.dense", h1))
    return layer_norm(_drop(linear(P, pre + ".output.dense", f), drop, pre + ".output.dropout", hp) + h1,
                      P[pre + ".output.LayerNorm.weight"], P[pre + ".output.LayerNorm.bias"])


def connection_layer(P, pre, cfg, v, mask_v, t, mask_t, drop=None):
    """BertConnectionLayer.forward, vilbert.py:871-900. Subscript 1 = vision, 2 = text.
    BertBiAttention :738-823: ctx1 = text queries over vision keys/values, ctx2 = vision queries over
    text keys/values; BertBiOutput :844-855 with the argument swap of :890-892 (ctx2 -> vision stream
    via dense1/LayerNorm1, ctx1 -> text stream via dense2/LayerNorm2); q_dense1/2 are never called."""
    b = pre + ".biattention"
    nh = cfg["bi_num_attention_heads"]
    q1, k1, v1 = linear(P, b + ".query1", v), linear(P, b + ".key1", v), linear(P, b + ".value1", v)
    q2, k2, v2 = linear(P, b + ".query2", t), linear(P, b + ".key2", t), linear(P, b + ".value2", t)
    # dropout1 (v_attention_probs_dropout_prob) acts on probs1, dropout2 (attention_probs_dropout_prob) on probs2 (:730,:738,:778,:800)
    ctx1 = attention(q2, k1, v1, mask_v, nh, drop, b + ".dropout1", cfg["v_attention_probs_dropout_prob"])   # [B,Nt,Hb]
    ctx2 = attention(q1, k2, v2, mask_t, nh, drop, b + ".dropout2", cfg["attention_probs_dropout_prob"])     # [B,Nv,Hb]
    o = pre + ".biOutput"
    vp, tp = cfg["v_hidden_dropout_prob"], cfg["hidden_dropout_prob"]
    v1_ = layer_norm(_drop(linear(P, o + ".dense1", ctx2), drop, o + ".dropout1", vp) + v, P[o + ".LayerNorm1.weight"], P[o + ".LayerNorm1.bias"])
    t1_ = layer_norm(_drop(linear(P, o + ".dense2", ctx1), drop, o + ".dropout2", tp) + t, P[o + ".LayerNorm2.weight"], P[o + ".LayerNorm2.bias"])
    fv = gelu(linear(P, pre + ".v_intermediate.dense", v1_))
    v2_ = layer_norm(_drop(linear(P, pre + ".v_output.dense", fv), drop, pre + ".v_output.dropout", vp) + v1_,
                     P[pre + ".v_output.LayerNorm.weight"], P[pre + ".v_output.LayerNorm.bias"])
    ft = gelu(linear(P, pre + ".t_intermediate.dense", t1_))
    t2_ = layer_norm(_drop(linear(P, pre + ".t_output.dense", ft), drop, pre + ".t_output.dropout", tp) + t1_,
                     P[pre + ".t_output.LayerNorm.weight"], P[pre + ".t_output.LayerNorm.bias"])
    return v2_, t2_


def encoder(P, pre, cfg, t, v, mask_t, mask_v, drop=None, mask2=None):
    """BertEncoder.forward interleaving schedule, vilbert.py:934-1107 (fixed layers, in_batch_pairs,
    FAST_MODE off; with_coattention honoured). Returns the per-connection-layer outputs too
    (output_all_encoded_layers, :1075-1077)."""
    t_start = v_start = 0
    all_t, all_v = [], []
    n_t, n_v = cfg["num_hidden_layers"], cfg["v_num_hidden_layers"]
    for count, (v_end, t_end) in enumerate(zip(cfg["v_biattention_id"], cfg["t_biattention_id"])):
        # fixed_t_layer / fixed_v_layer (vilbert.py:968-1003): the first layers of each stream run under torch.no_grad()
        for i in range(t_start, t_end):
            if i < cfg.get("fixed_t_layer", 0):
                with torch.no_grad():
                    t = text_layer(P, f"{pre}.layer.{i}", cfg, t, mask_t, drop)
            else:
                t = text_layer(P, f"{pre}.layer.{i}", cfg, t, mask_t, drop)
        for i in range(v_start, v_end):
            if i < cfg.get("fixed_v_layer", 0):
                with torch.no_grad():
                    v = image_layer(P, f"{pre}.v_layer.{i}", cfg, v, mask_v, drop, t, mask2)
            else:
                v = image_layer(P, f"{pre}.v_layer.{i}", cfg, v, mask_v, drop, t, mask2)
        if count == 0 and cfg.get("in_batch_pairs"):
            # vilbert.py:1008-1040: every (text i, image j) combination of the batch becomes sample i * b + j
            b = t.shape[0]
            v = v.unsqueeze(0).expand(b, *v.shape).contiguous().view(b * b, v.shape[1], v.shape[2])
            mask_v = mask_v.unsqueeze(0).expand(b, *mask_v.shape).contiguous().view(b * b, 1, 1, mask_v.shape[-1])
            t = t.unsqueeze(1).expand(b, b, t.shape[1], t.shape[2]).contiguous().view(b * b, t.shape[1], t.shape[2])
            mask_t = mask_t.unsqueeze(1).expand(b, b, 1, 1, mask_t.shape[-1]).contiguous().view(b * b, 1, 1, mask_t.shape[-1])
        if count == 0 and cfg.get("fast_mode"):
            # FAST_MODE (vilbert.py:1042-1053): one caption against a batch of images — the text stream, computed once at batch 1
            # up to the first connection layer, is broadcast to the image batch from there on
            t = t.expand(v.shape[0], t.shape[1], t.shape[2])
            mask_t = mask_t.expand(v.shape[0], mask_t.shape[1], mask_t.shape[2], mask_t.shape[3])
        if cfg["with_coattention"]:
            v, t = connection_layer(P, f"{pre}.c_layer.{count}", cfg, v, mask_v, t, mask_t, drop)
        v_start, t_start = v_end, t_end
        all_t.append(t)
        all_v.append(v)
    for i in range(v_start, n_v):
        v = image_layer(P, f"{pre}.v_layer.{i}", cfg, v, mask_v, drop, t, mask2)
    for i in range(t_start, n_t):
        t = text_layer(P, f"{pre}.layer.{i}", cfg, t, mask_t, drop)
    return t, v, all_t, all_v


# --------------------------------------------------------------------------- embeddings / model
def text_embeddings(P, pre, cfg, input_ids, token_type_ids, task_ids, drop=None):
    """BertEmbeddings.forward, vilbert.py:346-367. Positions are always arange(seq) (:349-352); the task
    embedding row is inserted at index 1 after the sum (:358-362); LN after the concat. padding_idx=0 on
    the word embeddings (:328-330) only zeroes that row's gradient."""
    B, N = input_ids.shape
    pos = torch.arange(N, device=input_ids.device).unsqueeze(0).expand(B, N)
    e = (F.embedding(input_ids, P[pre + ".word_embeddings.weight"], padding_idx=0) + F.embedding(pos, P[pre + ".position_embeddings.weight"])
         + F.embedding(token_type_ids, P[pre + ".token_type_embeddings.weight"]))
    if cfg["task_specific_tokens"]:
        te = F.embedding(task_ids, P[pre + ".task_embeddings.weight"])
        e = torch.cat([e[:, 0:1], te, e[:, 1:]], dim=1)
    return _drop(layer_norm(e, P[pre + ".LayerNorm.weight"], P[pre + ".LayerNorm.bias"]), drop, pre + ".dropout", cfg["hidden_dropout_prob"])


def image_embeddings(P, pre, feat, loc, drop=None, p=0.0):
    """BertImageEmbeddings.forward, vilbert.py:1421-1432 (its dropout uses hidden_dropout_prob, :1419)."""
    return _drop(layer_norm(linear(P, pre + ".image_embeddings", feat) + linear(P, pre + ".image_location_embeddings", loc),
                            P[pre + ".LayerNorm.weight"], P[pre + ".LayerNorm.bias"]), drop, pre + ".dropout", p)


def bert_model(P, cfg, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None,
               image_attention_mask=None, co_attention_mask=None, task_ids=None, prefix="bert",
               output_all_encoded_layers=False, drop=None):
    """BertModel.forward, vilbert.py:1309-1406: default masks :1322-1329, task-token mask extension
    :1331-1334, additive masks (1-m)*-10000 :1341-1362, embeddings, encoder, poolers (:1116-1122,
    :1131-1137: Linear+ReLU on token 0). co_attention_mask is accepted and unused (:774-775, :796-797)."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_txt)
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_txt)
    if image_attention_mask is None:
        image_attention_mask = torch.ones(input_imgs.shape[0], input_imgs.shape[1], device=input_txt.device).type_as(input_txt)
    if cfg["task_specific_tokens"]:
        attention_mask = torch.cat([torch.ones_like(attention_mask[:, :1]), attention_mask], dim=1)
    dt = P[prefix + ".embeddings.word_embeddings.weight"].dtype
    mask_t = (1.0 - attention_mask[:, None, None, :].to(dt)) * -10000.0
    mask_v = (1.0 - image_attention_mask[:, None, None, :].to(dt)) * -10000.0
    t = text_embeddings(P, prefix + ".embeddings", cfg, input_txt, token_type_ids, task_ids, drop)
    v = image_embeddings(P, prefix + ".v_embeddings", input_imgs, image_loc, drop, cfg["hidden_dropout_prob"])
    mask2 = attention_mask.unsqueeze(2).to(dt)      # extended_attention_mask2 (:1344): the dynamic_attention text pooling weights
    t, v, all_t, all_v = encoder(P, prefix + ".encoder", cfg, t, v, mask_t, mask_v, drop, mask2)
    if output_all_encoded_layers:
        # :1098-1101 + :1388-1394 — in this mode the encoder returns only the per-connection-layer states and the poolers see
        # encoded_layers[-1], i.e. the output of the LAST CONNECTION LAYER (the tail layers' result is dropped)
        t, v = all_t[-1], all_v[-1]
    pooled_t = torch.relu(linear(P, prefix + ".t_pooler.dense", t[:, 0]))
    pooled_v = torch.relu(linear(P, prefix + ".v_pooler.dense", v[:, 0]))
    if output_all_encoded_layers:
        return all_t, all_v, pooled_t, pooled_v
    return t, v, pooled_t, pooled_v


# --------------------------------------------------------------------------- heads
def simple_classifier(P, pre, x):
    """SimpleClassifier, vilbert.py:1711-1722: Linear -> GeLU -> LayerNorm -> Linear."""
    h = gelu(linear(P, pre + ".logit_fc.0", x))
    h = layer_norm(h, P[pre + ".logit_fc.2.weight"], P[pre + ".logit_fc.2.bias"])
    return linear(P, pre + ".logit_fc.3", h)


def pretraining_heads(P, cfg, seq_t, seq_v, pooled_t, pooled_v, prefix="cls", drop=None):
    """BertPreTrainingHeads.forward, vilbert.py:1228-1243; LM head :1193-1196 (decoder tied to the word
    embeddings, :1190, + output-only bias); image head :1255-1258; transforms :1152-1156, :1172-1176."""
    pooled = pooled_t * pooled_v if cfg["fusion_method"] == "mul" else pooled_t + pooled_v
    pooled = _drop(pooled, drop, prefix + ".dropout", 0.1)      # BertPreTrainingHeads.dropout = nn.Dropout(0.1), :1233
    ht = layer_norm(gelu(linear(P, prefix + ".predictions.transform.dense", seq_t)),
                    P[prefix + ".predictions.transform.LayerNorm.weight"], P[prefix + ".predictions.transform.LayerNorm.bias"])
    dec_w = P.get(prefix + ".predictions.decoder.weight", P["bert.embeddings.word_embeddings.weight"])
    scores_t = F.linear(ht, dec_w) + P[prefix + ".predictions.bias"]
    seq_rel = linear(P, prefix + ".bi_seq_relationship", pooled)
    hv = layer_norm(gelu(linear(P, prefix + ".imagePredictions.transform.dense", seq_v)),
                    P[prefix + ".imagePredictions.transform.LayerNorm.weight"], P[prefix + ".imagePredictions.transform.LayerNorm.bias"])
    scores_v = linear(P, prefix + ".imagePredictions.decoder", hv)
    return scores_t, scores_v, seq_rel


def vilbert_for_vl_tasks(P, cfg, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None,
                         image_attention_mask=None, co_attention_mask=None, task_ids=None, drop=None):
    """VILBertForVLTasks.forward, vilbert.py:1638-1708 (eval mode: every dropout is identity). Returns
    the reference's tuple order (:1697-1708) minus all_attention_mask, preceded by the BertModel outputs:
    (seq_t, seq_v, pooled_t, pooled_v), (vil_prediction, vil_prediction_gqa, vil_logit,
    vil_binary_prediction, vil_tri_prediction, vision_prediction, vision_logit, linguisic_prediction,
    linguisic_logit)."""
    seq_t, seq_v, pooled_t, pooled_v = bert_model(P, cfg, input_txt, input_imgs, image_loc, token_type_ids, attention_mask,
                                                  image_attention_mask, co_attention_mask, task_ids, drop=drop)
    linguisic_prediction, vision_prediction, vil_binary_prediction = pretraining_heads(P, cfg, seq_t, seq_v, pooled_t, pooled_v, drop=drop)
    pooled = pooled_t * pooled_v if cfg["fusion_method"] == "mul" else pooled_t + pooled_v
    hp = drop.head_p if drop is not None else 0.0
    pooled = _drop(pooled, drop, "dropout.pooled", hp)           # self.dropout(pooled), :1677-1682
    vil_prediction = simple_classifier(P, "vil_prediction", pooled)
    vil_prediction_gqa = simple_classifier(P, "vil_prediction_gqa", pooled)
    if pooled.shape[0] % 2 == 0:  # :1686-1689 — pairs consecutive samples; odd B keeps the NSP-style output
        vil_binary_prediction = simple_classifier(P, "vil_binary_prediction", pooled.view(-1, pooled.shape[1] * 2))
    vil_logit = linear(P, "vil_logit", pooled)
    vil_tri_prediction = linear(P, "vil_tri_prediction", pooled)
    dt = seq_v.dtype
    vision_logit = linear(P, "vision_logit", _drop(seq_v, drop, "dropout.seq_v", hp)) + ((1.0 - image_attention_mask.to(dt)) * -10000.0).unsqueeze(2)
    linguisic_logit = linear(P, "linguisic_logit", _drop(seq_t, drop, "dropout.seq_t", hp))
    return (seq_t, seq_v, pooled_t, pooled_v), (vil_prediction, vil_prediction_gqa, vil_logit, vil_binary_prediction,
                                                vil_tri_prediction, vision_prediction, vision_logit, linguisic_prediction,
                                                linguisic_logit)


HEAD_NAMES = ("vil_prediction", "vil_prediction_gqa", "vil_logit", "vil_binary_prediction", "vil_tri_prediction",
              "vision_prediction", "vision_logit", "linguisic_prediction", "linguisic_logit")
BERT_OUT_NAMES = ("sequence_output_t", "sequence_output_v", "pooled_output_t", "pooled_output_v")


def nce_negative_indices(B, R, num_negative):
    """visual_target == 2, vilbert.py:1524-1557: per (sample, region) the flat indices (row * R + col into the [B * R] region
    table) of int(0.7 n) negatives from OTHER samples and int(0.3 n) from OTHER regions of the same sample. Draws from torch's
    global CPU generator in the reference's order (row_across, col_across, col_inside; Tensor.random_(0, hi) is [0, hi)), so
    that under the same torch.manual_seed it reproduces the reference's sample exactly."""
    n_across, n_inside = int(num_negative * 0.7), int(num_negative * 0.3)
    rows = torch.empty(B, R, n_across, dtype=torch.long).random_(0, B - 1)
    cols = torch.empty(B, R, n_across, dtype=torch.long).random_(0, R)
    own = torch.arange(B).view(B, 1, 1)
    rows = torch.where((rows == own) & (own < B - 1), torch.full_like(rows, B - 1), rows)       # a sample is never its own negative
    inside = torch.empty(B, R, n_inside, dtype=torch.long).random_(0, R - 1)
    reg = torch.arange(R).view(1, R, 1)
    inside = torch.where((inside == reg) & (reg < R - 1), torch.full_like(inside, R - 1), inside)   # nor a region its own
    return torch.cat((rows * R + cols, own * R + inside), dim=2)


def pretraining_losses(P, cfg, input_ids, image_feat, image_loc, token_type_ids, attention_mask, image_attention_mask,
                       masked_lm_labels, image_label, image_target, next_sentence_label, neg_index=None):
    """BertForMultiModalPreTraining.forward with labels, vilbert.py:1471-1590: masked-LM CE (ignore_index -1), NSP/alignment CE
    and the masked-region loss (global region dropped, :1506) selected by config.visual_target:
      0  KL-div against the soft class target, normalised by max(sum(image_label == 1), 0) as written at :1521
      1  MSE regression of the region feature, mean over the masked ELEMENTS (:1507-1513)
      2  cross-entropy of the true feature (index 0) against sampled negatives scored by the dot product with the prediction
         (:1523-1575); neg_index [B, R, n] = nce_negative_indices(...) (drawn here, in the reference's order, when None)."""
    seq_t, seq_v, pooled_t, pooled_v = bert_model(P, cfg, input_ids, image_feat, image_loc, token_type_ids, attention_mask,
                                                  image_attention_mask)
    scores_t, scores_v, seq_rel = pretraining_heads(P, cfg, seq_t, seq_v, pooled_t, pooled_v)
    scores_v = scores_v[:, 1:]
    masked = image_label == 1
    vt = cfg.get("visual_target", 0)
    if vt == 1:
        img_loss = F.mse_loss(scores_v, image_target, reduction="none")
        masked_img_loss = torch.sum(img_loss * masked.unsqueeze(2).float()) / max(torch.sum(masked.unsqueeze(2).expand_as(img_loss)), 1)
    elif vt == 2:
        B, R, _ = scores_v.shape
        if neg_index is None:
            neg_index = nce_negative_indices(B, R, cfg.get("num_negative", 128))
        neg_index = neg_index.to(scores_v.device)
        samples = torch.cat((image_target[masked].unsqueeze(1), image_target.reshape(B * R, -1)[neg_index[masked]]), dim=1)
        score = torch.bmm(samples, scores_v[masked].unsqueeze(2)).squeeze(2)
        masked_img_loss = F.cross_entropy(score, torch.zeros(score.shape[0], dtype=torch.long, device=score.device))
    else:
        img_loss = F.kl_div(F.log_softmax(scores_v, dim=2), image_target, reduction="none")
        masked_img_loss = torch.sum(img_loss * masked.unsqueeze(2).float()) / max(torch.sum(masked), 0)
    masked_lm_loss = F.cross_entropy(scores_t.view(-1, scores_t.shape[-1]), masked_lm_labels.view(-1), ignore_index=-1)
    nsp_loss = F.cross_entropy(seq_rel.view(-1, 2), next_sentence_label.view(-1), ignore_index=-1)
    return masked_lm_loss, masked_img_loss, nsp_loss


# --------------------------------------------------------------------------- parameters / inputs
def param_shapes(cfg, with_task_heads=True):
    """Reference state_dict names -> shapes. nn.Linear.weight is [out, in]."""
    Ht, It, Hv, Iv, Hb = cfg["hidden_size"], cfg["intermediate_size"], cfg["v_hidden_size"], cfg["v_intermediate_size"], cfg["bi_hidden_size"]
    S = {}

    def lin(n, o, i):
        S[n + ".weight"] = (o, i); S[n + ".bias"] = (o,)

    def ln(n, h):
        S[n + ".weight"] = (h,); S[n + ".bias"] = (h,)

    S["bert.embeddings.word_embeddings.weight"] = (cfg["vocab_size"], Ht)
    S["bert.embeddings.position_embeddings.weight"] = (cfg["max_position_embeddings"], Ht)
    S["bert.embeddings.token_type_embeddings.weight"] = (cfg["type_vocab_size"], Ht)
    ln("bert.embeddings.LayerNorm", Ht)
    if cfg["task_specific_tokens"]:
        S["bert.embeddings.task_embeddings.weight"] = (20, Ht)
    lin("bert.v_embeddings.image_embeddings", Hv, cfg["v_feature_size"])
    lin("bert.v_embeddings.image_location_embeddings", Hv, 5)
    ln("bert.v_embeddings.LayerNorm", Hv)
    for kind, n, H, I in (("layer", cfg["num_hidden_layers"], Ht, It), ("v_layer", cfg["v_num_hidden_layers"], Hv, Iv)):
        for i in range(n):
            p = f"bert.encoder.{kind}.{i}"
            for nm in ("query", "key", "value"):
                lin(f"{p}.attention.self.{nm}", H, H)
            if kind == "v_layer" and cfg.get("dynamic_attention"):      # BertImageSelfAttention.dyLinear_q / _k (:561-563)
                lin(f"{p}.attention.self.dyLinear_q", H, Ht); lin(f"{p}.attention.self.dyLinear_k", H, Ht)
            lin(f"{p}.attention.output.dense", H, H); ln(f"{p}.attention.output.LayerNorm", H)
            lin(f"{p}.intermediate.dense", I, H)
            lin(f"{p}.output.dense", H, I); ln(f"{p}.output.LayerNorm", H)
    for i in range(len(cfg["v_biattention_id"])):
        p = f"bert.encoder.c_layer.{i}"
        for nm in ("query1", "key1", "value1"):
            lin(f"{p}.biattention.{nm}", Hb, Hv)
        for nm in ("query2", "key2", "value2"):
            lin(f"{p}.biattention.{nm}", Hb, Ht)
        lin(f"{p}.biOutput.dense1", Hv, Hb); ln(f"{p}.biOutput.LayerNorm1", Hv); lin(f"{p}.biOutput.q_dense1", Hv, Hb)
        lin(f"{p}.biOutput.dense2", Ht, Hb); ln(f"{p}.biOutput.LayerNorm2", Ht); lin(f"{p}.biOutput.q_dense2", Ht, Hb)
        lin(f"{p}.v_intermediate.dense", Iv, Hv); lin(f"{p}.v_output.dense", Hv, Iv); ln(f"{p}.v_output.LayerNorm", Hv)
        lin(f"{p}.t_intermediate.dense", It, Ht); lin(f"{p}.t_output.dense", Ht, It); ln(f"{p}.t_output.LayerNorm", Ht)
    lin("bert.t_pooler.dense", Hb, Ht); lin("bert.v_pooler.dense", Hb, Hv)
    S["cls.predictions.bias"] = (cfg["vocab_size"],)
    lin("cls.predictions.transform.dense", Ht, Ht); ln("cls.predictions.transform.LayerNorm", Ht)
    S["cls.predictions.decoder.weight"] = (cfg["vocab_size"], Ht)   # tied to the word embeddings
    lin("cls.bi_seq_relationship", 2, Hb)
    lin("cls.imagePredictions.transform.dense", Hv, Hv); ln("cls.imagePredictions.transform.LayerNorm", Hv)
    lin("cls.imagePredictions.decoder", cfg["v_target_size"], Hv)
    if with_task_heads:
        for nm, i, o in (("vil_prediction", Hb, 3129), ("vil_prediction_gqa", Hb, 1533), ("vil_binary_prediction", 2 * Hb, 2)):
            lin(f"{nm}.logit_fc.0", 2 * Hb, i); ln(f"{nm}.logit_fc.2", 2 * Hb); lin(f"{nm}.logit_fc.3", o, 2 * Hb)
        lin("vil_logit", 1, Hb); lin("vil_tri_prediction", 3, Hb); lin("vision_logit", 1, Hv); lin("linguisic_logit", 1, Ht)
    return S


def synth_params(cfg, seed=0, dtype=torch.float32, device="cpu", with_task_heads=True, qk_scale=1.0):
    """Deterministic parameters independent of module construction order: each tensor is drawn from its
    own generator seeded by (seed, name). Weights ~ N(0, 0.02) like init_weights (vilbert.py:1274-1285) but
    biases and LayerNorm affine are also randomised so that they are exercised. qk_scale > 1 multiplies
    the query/key weights to create peaked softmax rows (SURVEY.md §8c adversarial case i)."""
    import zlib
    P = {}
    for name, shape in param_shapes(cfg, with_task_heads).items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if name == "cls.predictions.decoder.weight":
            continue
        if ".LayerNorm" in name or ".logit_fc.2." in name:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if name.endswith("weight") else 0.05 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
            if qk_scale != 1.0 and any(k in name for k in (".query", ".key")):
                t = t * qk_scale
        P[name] = t.to(dtype).to(device)
    P["cls.predictions.decoder.weight"] = P["bert.embeddings.word_embeddings.weight"]
    return P


def synth_inputs(cfg, B, Nv, Nt, seed=1234, device="cpu", ragged=True, task_id=None):
    """Synthetic (region-feature, token-id) batch, SURVEY.md §8d "Synthetic inputs": post-ReLU features,
    row 0 = mean of the valid rows / (0,0,1,1,1) box, prefix-valid masks."""
    g = torch.Generator().manual_seed(seed)
    V = cfg["vocab_size"]
    ids = torch.randint(0, V, (B, Nt), generator=g)
    ids[:, 0] = 101 % V
    if ragged:
        lt = torch.randint((Nt + 1) // 2, Nt + 1, (B,), generator=g)
        lv = torch.randint(min(10, Nv), Nv + 1, (B,), generator=g)
        lt[0] = 1 if B > 2 else lt[0]       # a length-1 text row (only CLS valid)
    else:
        lt = torch.full((B,), Nt); lv = torch.full((B,), Nv)
    am = (torch.arange(Nt)[None] < lt[:, None]).long()
    im = (torch.arange(Nv)[None] < lv[:, None]).long()
    feat = torch.relu(torch.randn(B, Nv, cfg["v_feature_size"], generator=g)) * im[..., None]
    denom = (im[:, 1:].sum(1, keepdim=True).clamp_min(1)).float()
    feat[:, 0] = (feat[:, 1:] * im[:, 1:, None]).sum(1) / denom
    xy = torch.rand(B, Nv, 2, generator=g) * 0.7
    wh = 0.05 + torch.rand(B, Nv, 2, generator=g) * 0.25
    loc = torch.cat([xy, xy + wh, (wh[..., :1] * wh[..., 1:])], dim=-1) * im[..., None]
    loc[:, 0] = torch.tensor([0.0, 0.0, 1.0, 1.0, 1.0])
    out = dict(input_txt=ids, input_imgs=feat, image_loc=loc, token_type_ids=torch.zeros_like(ids), attention_mask=am,
               image_attention_mask=im, co_attention_mask=torch.zeros(B, Nv, Nt), task_ids=None)
    if cfg["task_specific_tokens"]:
        out["task_ids"] = torch.full((B, 1), 1 if task_id is None else task_id, dtype=torch.long)
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in out.items()}


def vqa_loss(vil_prediction, target):
    """task_utils.py:325-327 — BCE-with-logits, mean, times the number of answers."""
    return F.binary_cross_entropy_with_logits(vil_prediction, target, reduction="mean") * target.shape[1]


def synth_vqa_target(B, n_ans=3129, seed=99, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    tgt = torch.zeros(B, n_ans)
    idx = torch.randint(0, n_ans, (B, 3), generator=g)
    val = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (B, 3), generator=g)]
    tgt.scatter_(1, idx, val)
    return tgt.to(device)


# --------------------------------------------------------------------------- reduced-precision operand modes
# The reference algorithm under the arithmetic contract of the engine's tensor-core modes: every F.linear / torch.matmul
# rounds its OPERANDS (fp32 accumulation, fp32 everything else), forward and backward. Forward operands (activations,
# weights) use _FWD_DT, gradient operands (dy, dS) use _GRAD_DT. Linear backward reads W and x in the forward format (the
# engine's dgrad / wgrad read them in place); the attention backward rounds Q/K/V/P to the gradient format like the
# engine's attention backward (its panels are converted to bf16 in shared memory). Separates the error inherent to the
# operand formats from implementation error; not used for fp32 parity.
_FWD_DT, _GRAD_DT = torch.bfloat16, torch.bfloat16


def _rf(t):
    return t.to(_FWD_DT).to(torch.float32)


def _rg(t):
    return t.to(_GRAD_DT).to(torch.float32)


class _LinearBF16(torch.autograd.Function):
    """y = r(x) r(W)^T + b with fp32 accumulation; backward also rounds its matmul operands (dy, x, W)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        y = _rf(x) @ _rf(w).t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyr = _rg(dy)
        dx = dyr @ _rf(w)
        dw = dyr.reshape(-1, dyr.shape[-1]).t() @ _rf(x).reshape(-1, x.shape[-1])
        return dx, dw, (dy.reshape(-1, dy.shape[-1]).sum(0) if ctx.has_bias else None)


class _MatmulBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _rf(a) @ _rf(b)

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        return _rg(dy) @ _rg(b).transpose(-1, -2), _rg(a).transpose(-1, -2) @ _rg(dy)


class operand_mode:
    """Context manager: F.linear / torch.matmul of this oracle round their operands (see above).
    operand_mode() = the engine's default "fp16" precision (fp16 forward operands, bf16 gradient operands)."""

    def __init__(self, fwd=torch.float16, grad=torch.bfloat16):
        self.fwd, self.grad = fwd, grad

    def __enter__(self):
        global _FWD_DT, _GRAD_DT
        self._saved = (F.linear, torch.matmul, _FWD_DT, _GRAD_DT)
        _FWD_DT, _GRAD_DT = self.fwd, self.grad
        F.linear = lambda x, w, b=None: _LinearBF16.apply(x, w, b)
        torch.matmul = lambda a, b: _MatmulBF16.apply(a, b)
        return self

    def __exit__(self, *exc):
        global _FWD_DT, _GRAD_DT
        F.linear, torch.matmul, _FWD_DT, _GRAD_DT = self._saved
        return False


def bf16_operand_mode():
    """Every operand bf16 (the engine's "bf16" precision; round-1 arithmetic)."""
    return operand_mode(torch.bfloat16, torch.bfloat16)
