"""Drop-in module surface: BertModel, VILBertForVLTasks, BertForMultiModalPreTraining with the
reference's constructor / forward signatures, output tuples and state_dict key names
(vilbert/vilbert.py:1288-1406, :1435-1597, :1600-1708), executing on the B200 engine
(engine.py -> libvilbert_b200.so). There is no PyTorch / CPU fallback: constructing a model on a
non-CUDA device or without the built extension raises.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .config import BertConfig
from .engine import BERT_OUT_NAMES, HEAD_NAMES, Engine


# Any torch.optim.Optimizer.step() (pytorch_transformers.AdamW and the reference's RAdam subclass it) may have rewritten
# parameters through p.data: bump a global epoch that every model compares with the epoch its 16-bit weight copy was made at.
_OPT_EPOCH = [0]


def _on_optimizer_step(optimizer, args, kwargs):
    _OPT_EPOCH[0] += 1


try:
    torch.optim.optimizer.register_optimizer_step_post_hook(_on_optimizer_step)
except AttributeError:      # very old torch: train-mode forwards refresh unconditionally anyway
    pass


class _Node(nn.Module):
    """Container whose children / parameters are registered under the reference's dotted names."""


class _BertNode(_Node):
    """The `bert` sub-module of a model with heads: owns the encoder parameters (names bert.*) and, when called like
    the reference's `model.bert(...)` (vilbert.py:1652), runs the same engine and returns the BertModel 5-tuple."""

    def forward(self, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None,
                co_attention_mask=None, task_ids=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        owner = self.__dict__["_owner_ref"]()
        return owner._bert_forward(input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids,
                                   output_all_encoded_layers, output_all_attention_masks)


def _register_tree(root, store):
    params = {}
    for name in store.entries:
        view = store.p(name)
        prm = nn.Parameter(view, requires_grad=True)
        prm.grad = store.g(name)
        params[name] = prm
        _attach(root, name, prm)
    # tied decoder: the same Parameter object under its reference name (vilbert.py:1190, 1634-1636)
    _attach(root, "cls.predictions.decoder.weight", params["bert.embeddings.word_embeddings.weight"])
    return params


def _attach(root, dotted, prm):
    import weakref
    parts = dotted.split(".")
    mod = root
    for i, p in enumerate(parts[:-1]):
        if p not in mod._modules:
            if i == 0 and p == "bert":
                node = _BertNode()
                node.__dict__["_owner_ref"] = weakref.ref(root)
            else:
                node = _Node()
            mod.add_module(p, node)
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], prm)


class _EngineFn(torch.autograd.Function):
    """Bridges the engine's static plans into torch.autograd: forward runs the plan's forward pass and returns
    its outputs; backward copies d(loss)/d(outputs) into the plan's static buffers and runs the hand-written
    backward pass, which accumulates parameter gradients directly into the flat gradient buffer (the
    Parameters' .grad are views of it). A plan is specialised on the set of outputs that receive a gradient;
    the set seen in the previous backward of a shape is used as the hint for the next forward, so in a
    steady training loop forward and backward share one plan and nothing is recomputed."""

    @staticmethod
    def forward(ctx, model, names, anchor, inputs):
        Nt = inputs["input_txt"].shape[1]
        B, Nv = inputs["input_imgs"].shape[:2]     # FAST_MODE: the text batch is 1, the image batch sets the plan
        train = bool(model.training)
        hint = model._grad_hint.get((B, Nt, Nv, names, train), ())
        plan = model.engine.plan(B, Nt, Nv, grad_outputs=hint, heads=model._heads_for(names), train=train)
        model._sync_weights()
        if train:
            model.engine.bump_dropout_step()      # fresh nn.Dropout masks for this forward
        plan.load_inputs(**inputs)
        if model.engine.auto_graph:
            plan.maybe_capture_passes()
        plan.run_forward()
        ctx.model, ctx.names, ctx.inputs, ctx.plan, ctx.fwd_id, ctx.train = model, names, inputs, plan, plan.fwd_id, train
        ctx.drop_step = int(model.engine.drop_step_host)     # the masks this forward used (needed if it has to be recomputed)
        model._last_plan = plan
        ctx.set_materialize_grads(False)
        return tuple(plan.outputs[n].clone() for n in names)

    @staticmethod
    def backward(ctx, *grads):
        model, names, inputs, plan = ctx.model, ctx.names, ctx.inputs, ctx.plan
        live = tuple(n for n, g in zip(names, grads) if g is not None)
        if live:
            Nt = inputs["input_txt"].shape[1]
            B, Nv = inputs["input_imgs"].shape[:2]
            clobbered = model.engine.arena is not None and model.engine.arena_owner != (plan, ctx.fwd_id)   # another plan used the shared arena
            if frozenset(live) != plan.grad_outputs or plan.fwd_id != ctx.fwd_id or clobbered:
                model._grad_hint[(B, Nt, Nv, names, ctx.train)] = live
                plan = model.engine.plan(B, Nt, Nv, grad_outputs=live, heads=model._heads_for(names), train=ctx.train)
                plan.load_inputs(**inputs)     # different plan (or overwritten activations): recompute the forward ...
                eng = model.engine
                now = int(eng.drop_step_host)
                if ctx.train and now != ctx.drop_step:
                    eng.set_dropout_step(ctx.drop_step)   # ... with the dropout masks of the forward the loss was computed on
                    plan.run_forward()
                    eng.set_dropout_step(now)
                else:
                    plan.run_forward()
            model._attach_grads()
            if model.engine.auto_graph:
                plan.maybe_capture_passes()
            for n, g in zip(names, grads):
                if g is not None:
                    plan.gout[n].copy_(g.reshape(plan.gout[n].shape))
            plan.run_backward()
            if model._ddp_reducer is not None:     # data parallel: average the flat gradient buffer over the ranks (apex DDP, delay_allreduce=True)
                model._ddp_reducer.allreduce()
        return None, None, None, None


class BertPreTrainedModel(nn.Module):
    """Weight handling of the reference's PreTrainedModel (vilbert/utils.py:703-1032) restricted to local
    files: state_dict with the reference key names, legacy gamma/beta renaming, eval mode after loading."""

    config_class = BertConfig
    _heads = "vl"          # which heads own parameters: "vl" | "pretraining" | "none"

    def __init__(self, config, device=None, precision=None):
        """precision: "fp16" (default: fp16 forward operands, bf16 gradient operands, fp32 accumulation / residual stream),
        "fp32" (split precision, matches the reference's fp32 outputs to 1e-3) or "bf16"; default from
        $VILBERT_B200_PRECISION. The reference's constructors have no such argument: it selects what `model.half()` /
        default fp32 select there."""
        super().__init__()
        if not isinstance(config, BertConfig):
            raise ValueError("Parameter config should be an instance of class `BertConfig`.")
        self.config = config
        precision = precision or os.environ.get("VILBERT_B200_PRECISION", "fp16")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise L.VBError("vilbert_b200 models run on sm_100a GPUs only; there is no CPU path")
        self.engine = Engine(config, dev, heads=self._heads, precision=precision)
        self._params = _register_tree(self, self.engine.ps)
        self._grad_hint = {}
        self._anchor = torch.zeros((), device=dev, requires_grad=True)
        self._shadow_version = None
        self._opt_epoch = -1
        self._ddp_reducer = None
        self.init_weights()

    # ---- reference init (vilbert.py:1274-1285): N(0, initializer_range) for Linear/Embedding weights, zero bias, LN 1/0
    def init_weights(self):
        std = self.config.initializer_range
        with torch.no_grad():
            for name, prm in self._params.items():
                if "LayerNorm" in name or ".logit_fc.2." in name:
                    prm.fill_(1.0) if name.endswith("weight") else prm.zero_()
                elif name.endswith(".bias"):
                    prm.zero_()
                else:
                    prm.normal_(mean=0.0, std=std)

    def tie_weights(self):
        pass  # the decoder weight IS the word-embedding Parameter (registered twice)

    def _heads_for(self, names):
        """BertModel-only outputs never need the heads' forward."""
        return "none" if all(n in BERT_OUT_NAMES for n in names) else self._heads

    def _sync_weights(self):
        """Refreshes the 16-bit operand copy of the weights from the fp32 master parameters. Optimizers write parameters
        through `p.data` (pytorch_transformers.AdamW: p.data.addcdiv_, vilbert/optimization.py RAdam: p.data.copy_), which no
        tensor version counter sees, so in train mode the copy is refreshed before EVERY forward (one cast kernel) unless the
        engine's own fused optimizer produced it (engine.shadow_trusted). In eval mode the version counter of the flat buffer,
        load_state_dict() and a global post-hook on every torch.optim.Optimizer.step() mark it stale; after writing parameters
        by hand through .data in eval mode call `model.engine.shadow_clean = False`."""
        eng = self.engine
        v = eng.ps.flat._version
        stale = (not eng.shadow_clean) or v != self._shadow_version
        if not eng.shadow_trusted and (self.training or _OPT_EPOCH[0] != self._opt_epoch):
            stale = True
        if stale:
            eng.refresh_weights()
        self._shadow_version, self._opt_epoch = v, _OPT_EPOCH[0]

    def load_state_dict(self, state_dict, strict=True):
        """strict is honoured (missing / unexpected keys raise like nn.Module). The tied decoder weight is one Parameter
        registered under two names: a checkpoint carrying only one of them is complete."""
        sd = dict(state_dict)
        w, d = "bert.embeddings.word_embeddings.weight", "cls.predictions.decoder.weight"
        if w in sd and d not in sd:
            sd[d] = sd[w]
        elif d in sd and w not in sd:
            sd[w] = sd[d]
        r = super().load_state_dict(sd, strict=strict)
        self.engine.shadow_clean = False
        return r

    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffer. The Parameters' .grad stay views of it (set_to_none is accepted and ignored: the
        engine accumulates into the flat buffer, dropping the views would only hide the gradients from the optimizer)."""
        self.engine.zero_grad()
        self._attach_grads(zero_if_detached=False)

    def _attach_grads(self, zero_if_detached=True):
        """torch.optim.Optimizer.zero_grad() defaults to set_to_none=True and detaches every .grad from the flat buffer.
        Before a backward the views are re-attached; if they had been dropped since the last backward the flat buffer (which
        the engine kept accumulating into) is zeroed first, which is what the caller asked for."""
        ps = self.engine.ps
        detached = [name for name, prm in self._params.items() if prm.grad is None]
        if not detached:
            return
        if zero_if_detached:
            self.engine.zero_grad()
        for name in detached:
            self._params[name].grad = ps.g(name)

    def _apply(self, fn, recurse=True):
        raise L.VBError("vilbert_b200 models own flat CUDA parameter buffers that cannot be moved or re-typed in place "
                        "(construct the model on the target GPU; reduced precision is the `precision=` argument)")

    # The calls the reference's drivers make on a freshly built model (train_tasks.py:486-500: model.to(device), model.cuda(),
    # model.half() under --fp16) are accepted when they ask for what the model already is: fp32 master parameters on its GPU.
    def to(self, *args, **kwargs):
        device, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        mine = self.engine.device
        if device is not None and (device.type != "cuda" or (device.index is not None and device.index != mine.index)):
            raise L.VBError(f"vilbert_b200 model lives on {mine}; it cannot be moved to {device}")
        if dtype is not None and dtype != torch.float32:
            raise L.VBError("parameters stay fp32 (master weights); the tensor-core operand precision is chosen with precision=")
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def float(self):
        return self

    def half(self):
        """The reference's --fp16 path (`model.half()` + apex FP16_Optimizer, train_concap.py:504-505): here the default precision
        already runs fp16 forward operands with fp32 master weights and accumulation, so this is a no-op for precision "fp16"."""
        if self.engine.precision != "fp16":
            raise L.VBError('model.half(): construct the model with precision="fp16" (the default)')
        return self

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, state_dict=None, **kwargs):
        """Local files only (no network): a directory containing pytorch_model.bin (+ config.json) or a .bin file.
        Returns an eval-mode model like the reference (vilbert/utils.py:1022)."""
        kwargs.pop("default_gpu", None)
        if config is None:
            cfg_path = os.path.join(pretrained_model_name_or_path, "config.json")
            config = BertConfig.from_json_file(cfg_path)
        model = cls(config, *model_args, **kwargs)
        if state_dict is None:
            path = pretrained_model_name_or_path
            if os.path.isdir(path):
                path = os.path.join(path, "pytorch_model.bin")
            state_dict = torch.load(path, map_location="cpu")
        renamed = {}
        for k, v in state_dict.items():
            nk = k.replace("gamma", "weight") if "gamma" in k else k
            nk = nk.replace("beta", "bias") if "beta" in nk else nk
            if nk.startswith("module."):
                nk = nk[len("module."):]
            renamed[nk] = v
        own = model.state_dict()
        if not any(k.startswith("bert.") for k in renamed) and any(("bert." + k) in own for k in renamed):
            renamed = {"bert." + k: v for k, v in renamed.items()}   # base-model checkpoint into a model with heads
        # like the reference's loader (vilbert/utils.py:960-1012) missing and unexpected keys are tolerated but REPORTED
        missing = sorted(k for k in own if k not in renamed and k != "cls.predictions.decoder.weight")
        unexpected = sorted(k for k in renamed if k not in own)
        model.load_state_dict({k: v for k, v in renamed.items() if k in own}, strict=False)
        model.loading_info = {"missing_keys": missing, "unexpected_keys": unexpected}
        if missing or unexpected:
            import logging
            logging.getLogger(__name__).warning("from_pretrained(%s): %d missing keys (kept at their initial values): %s; %d unexpected keys (ignored): %s",
                                                pretrained_model_name_or_path, len(missing), missing[:8], len(unexpected), unexpected[:8])
        model.eval()
        return model

    def _attention_masks(self, output_all_attention_masks):
        """all_attention_mask of the reference's outputs: empty lists unless asked for; with config.visualization the attn_data
        dicts of every layer, without it one None per layer (BertEncoder appends whatever the layer returned, vilbert.py:974-1061)."""
        if not output_all_attention_masks:
            return ([], [], [])
        plan = self._last_plan
        if plan.viz:
            return plan.attention_export()
        c = self.config
        return ([None] * c.num_hidden_layers, [None] * c.v_num_hidden_layers, [None] * len(c.v_biattention_id))

    def _bert_forward(self, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids,
                      output_all_encoded_layers, output_all_attention_masks=False):
        """BertModel.forward outputs (vilbert.py:1388-1406). With output_all_encoded_layers the encoded-layer entries are lists with
        one tensor per connection layer (:1075-1077); only the last entry is connected to autograd here."""
        o = self._run(BERT_OUT_NAMES, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids)
        seq_t, seq_v = o["sequence_output_t"], o["sequence_output_v"]
        if output_all_encoded_layers:
            plan = self._last_plan
            B = seq_t.shape[0]
            enc_t = [a.f32.view(B, plan.Nt, -1).clone() for a in plan.enc_t]
            enc_v = [a.f32.view(B, plan.Nv, -1).clone() for a in plan.enc_v]
            # Reference quirk (:1098-1101, :1388-1394): in this mode the encoder returns ONLY the per-connection-layer states, and
            # BertModel pools encoded_layers[-1], i.e. the output of the last connection layer, not of the tail layers. This
            # inspection path (not the hot path) reproduces that with two tiny fp32 ops on the pooler parameters.
            pt = F.relu(F.linear(enc_t[-1][:, 0], self._params["bert.t_pooler.dense.weight"], self._params["bert.t_pooler.dense.bias"]))
            pv = F.relu(F.linear(enc_v[-1][:, 0], self._params["bert.v_pooler.dense.weight"], self._params["bert.v_pooler.dense.bias"]))
            return (enc_t, enc_v, pt, pv, self._attention_masks(output_all_attention_masks))
        return (seq_t, seq_v, o["pooled_output_t"], o["pooled_output_v"], self._attention_masks(output_all_attention_masks))

    # ---- shared forward machinery
    def _run(self, names, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids):
        inputs = dict(input_txt=input_txt, input_imgs=input_imgs, image_loc=image_loc, token_type_ids=token_type_ids,
                      attention_mask=attention_mask, image_attention_mask=image_attention_mask, task_ids=task_ids)
        outs = _EngineFn.apply(self, tuple(names), self._anchor, inputs)
        return dict(zip(names, outs))


class BertModel(BertPreTrainedModel):
    """Reference: vilbert/vilbert.py:1288-1406. Parameters live under the bare names (embeddings.*, encoder.*)."""
    _heads = "none"

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        return type(sd)((key[len("bert."):], v) for key, v in sd.items() if key.startswith("bert."))

    def load_state_dict(self, state_dict, strict=True):
        return super().load_state_dict({"bert." + k: v for k, v in state_dict.items()}, strict=strict)

    def forward(self, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None,
                co_attention_mask=None, task_ids=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        return self._bert_forward(input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids,
                                  output_all_encoded_layers, output_all_attention_masks)


class VILBertForVLTasks(BertPreTrainedModel):
    """Reference: vilbert/vilbert.py:1600-1708. forward returns the same 10-tuple in the same order
    (:1697-1708). co_attention_mask is accepted and ignored exactly like the reference (:774-775, 796-797)."""
    _heads = "vl"

    def __init__(self, config, num_labels=1, dropout_prob=0.1, default_gpu=True, device=None, precision=None):
        super().__init__(config, device, precision)
        self.num_labels = num_labels
        self.dropout_prob = dropout_prob
        self.engine.head_dropout_prob = dropout_prob
        self.fusion_method = config.fusion_method

    def forward(self, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None,
                co_attention_mask=None, task_ids=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        if output_all_encoded_layers:
            raise NotImplementedError("VILBertForVLTasks(output_all_encoded_layers=True) is not supported; use model.bert(..., output_all_encoded_layers=True)")
        if image_attention_mask is None:
            raise TypeError("image_attention_mask is required by VILBertForVLTasks.forward (vilbert.py:1693)")
        o = self._run(HEAD_NAMES, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, task_ids)
        return tuple(o[n] for n in HEAD_NAMES) + (self._attention_masks(output_all_attention_masks),)


class BertForMultiModalPreTraining(BertPreTrainedModel):
    """Reference: vilbert/vilbert.py:1435-1597; the masked-region objective follows config.visual_target (0: KL divergence to the
    detector's class distribution, 1: feature regression, 2: noise-contrastive against sampled regions). The encoder
    and the three heads run on the engine; the three scalar losses are formed with torch on the head outputs
    exactly as the reference does (:1506-1590) and their gradients re-enter the engine through autograd."""
    _heads = "pretraining"

    def __init__(self, config, device=None, precision=None):
        super().__init__(config, device, precision)
        self.visual_target = config.visual_target
        self.num_negative = config.num_negative
        if self.visual_target not in (0, 1, 2):
            raise ValueError("visual_target must be 0, 1 or 2")

    def _nce_negatives(self, B, R, dev):
        """Flat region indices [B, R, n] of the negatives of visual_target == 2: 70 % from other samples, 30 % from other regions of
        the same sample, sampled on the device (`self.nce_sampler`, a callable (B, R, device) -> index, replaces it in tests)."""
        n_across, n_inside = int(self.num_negative * 0.7), int(self.num_negative * 0.3)
        rows = torch.randint(0, max(B - 1, 1), (B, R, n_across), device=dev)
        own = torch.arange(B, device=dev).view(B, 1, 1)
        rows = torch.where((rows == own) & (own < B - 1), torch.full_like(rows, B - 1), rows)     # never the sample itself
        across = rows * R + torch.randint(0, R, (B, R, n_across), device=dev)
        cols = torch.randint(0, max(R - 1, 1), (B, R, n_inside), device=dev)
        reg = torch.arange(R, device=dev).view(1, R, 1)
        cols = torch.where((cols == reg) & (reg < R - 1), torch.full_like(cols, R - 1), cols)     # never the region itself
        return torch.cat((across, own * R + cols), dim=2)

    def _nce_region_loss(self, pred, target, masked):
        """visual_target == 2: for every masked region, CE over [its own target feature, sampled negatives] scored by the dot
        product with the prediction (pinned against the reference with the reference's own sample: tiny_visual_target_2.json)."""
        B, R, _ = pred.shape
        dev = pred.device
        sampler = getattr(self, "nce_sampler", None)
        index = (sampler(B, R, dev) if sampler is not None else self._nce_negatives(B, R, dev))[masked]
        flat = target.reshape(B * R, -1)
        samples = torch.cat((target[masked].unsqueeze(1), flat[index]), dim=1)
        score = torch.bmm(samples, pred[masked].unsqueeze(2)).squeeze(2)
        return F.cross_entropy(score, torch.zeros(score.size(0), dtype=torch.long, device=dev))

    def forward(self, input_ids, image_feat, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None,
                masked_lm_labels=None, image_label=None, image_target=None, next_sentence_label=None, output_all_attention_masks=False):
        names = ("linguisic_prediction", "vision_prediction", "seq_relationship_score")
        o = self._run(names, input_ids, image_feat, image_loc, token_type_ids, attention_mask, image_attention_mask, None)
        prediction_scores_t, prediction_scores_v, seq_relationship_score = (o[n] for n in names)
        if masked_lm_labels is not None and next_sentence_label is not None and image_target is not None:
            prediction_scores_v = prediction_scores_v[:, 1:]
            masked = image_label == 1
            if self.visual_target == 0:      # KL to the soft class target (vilbert.py:1515-1521)
                img_loss = F.kl_div(F.log_softmax(prediction_scores_v, dim=2), image_target, reduction="none")
                masked_img_loss = torch.sum(img_loss * masked.unsqueeze(2).float()) / max(torch.sum(masked), 0)
            elif self.visual_target == 1:    # regression of the 2048-d region feature, mean over the masked elements (:1507-1513)
                img_loss = F.mse_loss(prediction_scores_v, image_target, reduction="none")
                masked_img_loss = torch.sum(img_loss * masked.unsqueeze(2).float()) / max(torch.sum(masked.unsqueeze(2).expand_as(img_loss)), 1)
            else:                            # contrastive: the true feature against num_negative sampled regions (:1523-1575)
                masked_img_loss = self._nce_region_loss(prediction_scores_v, image_target, masked)
            masked_lm_loss = F.cross_entropy(prediction_scores_t.view(-1, self.config.vocab_size), masked_lm_labels.view(-1), ignore_index=-1)
            next_sentence_loss = F.cross_entropy(seq_relationship_score.view(-1, 2), next_sentence_label.view(-1), ignore_index=-1)
            return masked_lm_loss.unsqueeze(0), masked_img_loss.unsqueeze(0), next_sentence_loss.unsqueeze(0)
        return prediction_scores_t, prediction_scores_v, seq_relationship_score, self._attention_masks(output_all_attention_masks)
