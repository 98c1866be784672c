"""Execution engine of the B200 ViLBERT hot path.

Mirrors, operation for operation, what the reference's eager modules do in
VILBertForVLTasks.forward -> BertModel.forward -> BertEncoder.forward (vilbert/vilbert.py:1638-1708,
:1309-1406, :934-1107) and what autograd derives from them, but as a static *plan*: for a given
(B, Nt, Nv) shape every activation buffer is allocated once and the forward and backward passes
become flat lists of C-ABI calls into libvilbert_b200.so (tcgen05 GEMMs, fused attention, row-wise
kernels). A plan can be replayed eagerly (a tight loop of ctypes calls) or captured into one CUDA
graph per pass. PyTorch is used for device memory, streams and (optionally) graph capture only.

Numerics. Accumulation is always fp32 (TMEM / registers); the residual stream, LayerNorm
statistics and outputs, softmax statistics, biases and every gradient accumulation are fp32; 16-bit copies of
activations and weights exist only as tensor-core operands. Three operand precisions (Engine(precision=...)):
  "fp16" (default)  forward operands (activations, weights) fp16 — 11 significant bits, the reference's own reduced
                    precision (model.half(), train_concap.py:504-505) — gradient operands (dy, dS, ...) bf16 for range.
                    tcgen05 faults on fp16 x bf16 (measured), so every forward operand the backward contracts with a
                    gradient (weights for dgrad, saved activations for wgrad) also has a bf16 copy, written by the kernel
                    that produces it (Act.bw, ParamStore.shadow_b);
  "fp32"            split precision: every forward operand is stored as fp16 hi + lo and every forward contraction
                    runs as three tensor-core passes (hi.hi + lo.hi + hi.lo), ~fp32 accuracy (north_star 1e-3);
  "bf16"            every operand bf16 (round-1 arithmetic; kept for A/B measurements).
"""
import ctypes as C
import math
import zlib
from collections import OrderedDict

import torch

from . import _lib as L

F32, BF16, F16, I64 = torch.float32, torch.bfloat16, torch.float16, torch.int64
PRECISIONS = ("fp16", "fp32", "bf16")


def _pad8(n):
    return (n + 7) // 8 * 8


def dropout_site_id(name):
    """Stable 32-bit id of a dropout layer, shared with the oracle's mask generator (crc32 of its canonical name)."""
    return zlib.crc32(name.encode()) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ parameters
class ParamStore:
    """All parameters live in ONE flat fp32 buffer (reference state_dict names map to views of it), with a
    flat fp32 gradient buffer of the same layout (what the data-parallel all-reduce operates on) and a flat
    bf16 shadow used as GEMM operands. query/key/value weights of one attention are allocated contiguously
    so that the fused [3H, H] QKV GEMM reads them in place."""

    def __init__(self, cfg, device, heads="vl", op_dtype=F16, split=False):
        """heads: "vl" = pre-training heads + the 7 task heads (VILBertForVLTasks), "pretraining" = cls.* only
        (BertForMultiModalPreTraining), "none" = bare BertModel. op_dtype: format of the 16-bit weight shadow;
        split: also keep the low parts (shadow_lo) for the split-precision mode."""
        assert heads in ("vl", "pretraining", "none")
        self.op_dtype, self.split = op_dtype, split
        self.heads = heads
        with_task_heads = heads == "vl"
        self.cfg = cfg
        self.device = device
        self.entries = OrderedDict()   # reference name -> (offset, shape)
        self.fused = {}                # fused qkv name -> (offset, shape)
        self._off = 0
        c = cfg
        Ht, It, Hv, Iv, Hb = c.hidden_size, c.intermediate_size, c.v_hidden_size, c.v_intermediate_size, c.bi_hidden_size
        for h in (Ht, It, Hv, Iv, Hb, c.v_feature_size):
            if h % 8:
                raise ValueError("vilbert_b200: hidden sizes must be multiples of 8 (TMA row pitch)")
        add, lin, ln = self._add, self._lin, self._ln
        add("bert.embeddings.word_embeddings.weight", (c.vocab_size, Ht))
        add("bert.embeddings.position_embeddings.weight", (c.max_position_embeddings, Ht))
        add("bert.embeddings.token_type_embeddings.weight", (c.type_vocab_size, Ht))
        ln("bert.embeddings.LayerNorm", Ht)
        if c.task_specific_tokens:
            add("bert.embeddings.task_embeddings.weight", (20, Ht))
        lin("bert.v_embeddings.image_embeddings", Hv, c.v_feature_size)
        lin("bert.v_embeddings.image_location_embeddings", Hv, 5)
        ln("bert.v_embeddings.LayerNorm", Hv)
        # Encoder parameters are laid out in EXECUTION order (the interleaving schedule of BertEncoder.forward,
        # vilbert.py:960-1096): backward then finishes the flat gradient buffer from its end towards its start, so the
        # data-parallel all-reduce can start on finished tail ranges while earlier layers are still in backward.
        def block(kind, i, H, I):
            p = f"bert.encoder.{kind}.{i}"
            self._qkv(f"{p}.attention.self", ("query", "key", "value"), H, H)
            if kind == "v_layer" and getattr(c, "dynamic_attention", False):
                # BertImageSelfAttention.dyLinear_q / dyLinear_k (vilbert.py:561-563), contiguous so that one [2Hv, Ht] GEMM computes both gates
                self._qkv(f"{p}.attention.self", ("dyLinear_q", "dyLinear_k"), H, Ht, fused="dy")
            lin(f"{p}.attention.output.dense", H, H); ln(f"{p}.attention.output.LayerNorm", H)
            lin(f"{p}.intermediate.dense", I, H)
            lin(f"{p}.output.dense", H, I); ln(f"{p}.output.LayerNorm", H)

        def conn(i):
            p = f"bert.encoder.c_layer.{i}"
            self._qkv(f"{p}.biattention", ("query1", "key1", "value1"), Hb, Hv, fused="qkv1")
            self._qkv(f"{p}.biattention", ("query2", "key2", "value2"), Hb, Ht, fused="qkv2")
            lin(f"{p}.biOutput.dense1", Hv, Hb); ln(f"{p}.biOutput.LayerNorm1", Hv); lin(f"{p}.biOutput.q_dense1", Hv, Hb)
            lin(f"{p}.biOutput.dense2", Ht, Hb); ln(f"{p}.biOutput.LayerNorm2", Ht); lin(f"{p}.biOutput.q_dense2", Ht, Hb)
            lin(f"{p}.v_intermediate.dense", Iv, Hv); lin(f"{p}.v_output.dense", Hv, Iv); ln(f"{p}.v_output.LayerNorm", Hv)
            lin(f"{p}.t_intermediate.dense", It, Ht); lin(f"{p}.t_output.dense", Ht, It); ln(f"{p}.t_output.LayerNorm", Ht)

        t_start = v_start = 0
        for count, (v_end, t_end) in enumerate(zip(c.v_biattention_id, c.t_biattention_id)):
            for i in range(t_start, t_end):
                block("layer", i, Ht, It)
            for i in range(v_start, v_end):
                block("v_layer", i, Hv, Iv)
            conn(count)
            v_start, t_start = v_end, t_end
        for i in range(v_start, c.v_num_hidden_layers):
            block("v_layer", i, Hv, Iv)
        for i in range(t_start, c.num_hidden_layers):
            block("layer", i, Ht, It)
        lin("bert.t_pooler.dense", Hb, Ht); lin("bert.v_pooler.dense", Hb, Hv)
        if heads != "none":
            add("cls.predictions.bias", (c.vocab_size,))
            lin("cls.predictions.transform.dense", Ht, Ht); ln("cls.predictions.transform.LayerNorm", Ht)
            lin("cls.bi_seq_relationship", 2, Hb)
            lin("cls.imagePredictions.transform.dense", Hv, Hv); ln("cls.imagePredictions.transform.LayerNorm", Hv)
            lin("cls.imagePredictions.decoder", c.v_target_size, Hv)
        self.with_task_heads = with_task_heads
        if with_task_heads:
            for nm, i, o in (("vil_prediction", Hb, 3129), ("vil_prediction_gqa", Hb, 1533), ("vil_binary_prediction", 2 * Hb, 2)):
                lin(f"{nm}.logit_fc.0", 2 * Hb, i); ln(f"{nm}.logit_fc.2", 2 * Hb); lin(f"{nm}.logit_fc.3", o, 2 * Hb)
            lin("vil_logit", 1, Hb); lin("vil_tri_prediction", 3, Hb); lin("vision_logit", 1, Hv); lin("linguisic_logit", 1, Ht)
        self.numel = self._off
        self.flat = torch.zeros(self.numel, dtype=F32, device=device)
        self.grad = torch.zeros(self.numel, dtype=F32, device=device)
        self.shadow = torch.zeros(self.numel, dtype=op_dtype, device=device)
        self.shadow_lo = torch.zeros(self.numel, dtype=op_dtype, device=device) if split else None
        # bf16 copy for the backward (dgrad: dy bf16 x W): the shadow itself when the operand format is already bf16
        self.shadow_b = self.shadow if op_dtype == BF16 else torch.zeros(self.numel, dtype=BF16, device=device)
        self.shadow_version = None

    def _add(self, name, shape):
        n = 1
        for s in shape:
            n *= s
        self.entries[name] = (self._off, tuple(shape))
        self._off += _pad8(n)

    def _lin(self, name, o, i):
        self._add(name + ".weight", (o, i)); self._add(name + ".bias", (o,))

    def _ln(self, name, h):
        self._add(name + ".weight", (h,)); self._add(name + ".bias", (h,))

    def _qkv(self, prefix, names, o, i, fused="qkv"):
        assert (o * i) % 8 == 0 and o % 8 == 0
        w0 = self._off
        for nm in names:
            self._add(f"{prefix}.{nm}.weight", (o, i))
        b0 = self._off
        for nm in names:
            self._add(f"{prefix}.{nm}.bias", (o,))
        self.fused[f"{prefix}.{fused}.weight"] = (w0, (len(names) * o, i))
        self.fused[f"{prefix}.{fused}.bias"] = (b0, (len(names) * o,))

    def _view(self, flat, name):
        off, shape = self.entries[name] if name in self.entries else self.fused[name]
        n = 1
        for s in shape:
            n *= s
        return flat[off:off + n].view(shape)

    def p(self, name):
        return self._view(self.flat, name)

    def g(self, name):
        return self._view(self.grad, name)

    def w16(self, name):
        return self._view(self.shadow, name)

    def w16lo(self, name):
        return self._view(self.shadow_lo, name) if self.split else None

    def w16b(self, name):
        return self._view(self.shadow_b, name)

    def refresh_shadow(self, stream):
        """16-bit operand copy (hi, and lo in split precision) of every parameter in one launch (belongs with the
        optimizer step in training)."""
        L.check(L.lib().vb_cast_f32_to_bf16(self.flat.data_ptr(), self.shadow.data_ptr(), self.numel, 1 if self.op_dtype == F16 else 0,
                                            self.shadow_lo.data_ptr() if self.split else None,
                                            self.shadow_b.data_ptr() if self.shadow_b is not self.shadow else None, stream), "vb_cast_f32_to_bf16")


# ------------------------------------------------------------------------------------------ plan
class _TrackedParams:
    """ParamStore proxy used while a plan is emitted: remembers, for every gradient range handed out during the
    backward emission, the index of the backward op about to write it (the last such index per range is kept)."""

    def __init__(self, ps, plan):
        self._ps, self._plan = ps, plan

    def __getattr__(self, name):
        return getattr(self._ps, name)

    def g(self, name):
        ps, plan = self._ps, self._plan
        off, shape = ps.entries[name] if name in ps.entries else ps.fused[name]
        n = 1
        for d in shape:
            n *= d
        if plan.cur is plan.bwd:
            plan.grad_touch[(off, n)] = len(plan.bwd)
        return ps.g(name)


class Act:
    """A residual-stream activation: fp32 values, 16-bit operand copy (b16; lo = its split-precision low part or None),
    fp32 gradient (lazily allocated)."""
    __slots__ = ("f32", "b16", "lo", "bw", "g32", "gw", "M", "H", "frozen")

    def __init__(self, f32, b16, M, H, lo=None, bw=None):
        """bw: bf16 copy of the operand for the backward's weight-gradient GEMM (== b16 when the operand format is bf16)."""
        self.f32, self.b16, self.lo, self.M, self.H = f32, b16, lo, M, H
        self.bw = b16 if bw is None else bw
        self.g32, self.gw = None, False
        self.frozen = False     # produced under the reference's torch.no_grad() (fixed_t_layer / fixed_v_layer): no gradient flows into it


# Objectives that can be fused into a plan, and the outputs each differentiates (task_utils.py:325-374, vilbert.py:1506-1590):
LOSS_HEADS = {
    "vqa": ("vil_prediction",),                 # VL-classifier: BCEWithLogits.mean() * 3129
    "gqa": ("vil_prediction_gqa",),             # VL-classifier-GQA: BCEWithLogits.mean() * 1533
    "vlogit_bce": ("vision_logit",),            # V-logit (refcoco*): BCEWithLogits(vision_logit [B,Nv,1], target).mean() * Nv
    "logit_ce": ("vil_logit",),                 # VL-logit (retrieval, VCR): CE over vil_logit.view(B / options, options)
    "binary_ce": ("vil_binary_prediction",),    # VL-binary-classifier (NLVR2): CE over the 2-way paired head
    "tri_ce": ("vil_tri_prediction",),          # VL-tri-classifier (SNLI-VE): CE over 3 classes
    "pretraining": ("linguisic_prediction", "vision_prediction", "seq_relationship_score"),   # masked-LM CE + masked-region KL + alignment CE
}

HEAD_NAMES = ("vil_prediction", "vil_prediction_gqa", "vil_logit", "vil_binary_prediction", "vil_tri_prediction",
              "vision_prediction", "vision_logit", "linguisic_prediction", "linguisic_logit")
BERT_OUT_NAMES = ("sequence_output_t", "sequence_output_v", "pooled_output_t", "pooled_output_v")


class Plan:
    """Static execution plan for one input shape. `grad_outputs` names the outputs that will receive a
    gradient in backward (dead branches are not emitted); `vqa_loss` fuses the VQA BCE objective
    (task_utils.py:325-327) and its gradient after the forward."""

    def __init__(self, engine, B, Nt, Nv, grad_outputs=(), vqa_loss=False, heads=None, train=False, loss=None):
        self.e, self.cfg = engine, engine.cfg
        self.grad_touch = {}           # (flat offset, numel) -> index of the last backward op writing that gradient range
        self.ps = _TrackedParams(engine.ps, self)
        self.lib = L.lib()
        self.dev = engine.device
        # in_batch_pairs (vilbert.py:1008-1040): at the first connection layer every (text i, image j) combination of the input
        # batch becomes one sample: the streams run at the input batch before it and at B^2 from there on
        self.pairs = bool(getattr(engine.cfg, "in_batch_pairs", False))
        self.Bin = B
        self.B, self.Nt_in, self.Nv = (B * B if self.pairs else B), Nt, Nv
        self.has_task = bool(self.cfg.task_specific_tokens)
        self.Nt = Nt + (1 if self.has_task else 0)
        # FAST_MODE (vilbert.py:1042-1053, eval_retrieval.py): one caption (text batch 1) against B images; inference only
        # config.visualization (vilbert.py:451-458, 610-617, 813-821): export attention probabilities, queries and keys per layer
        self.viz = bool(getattr(self.cfg, "visualization", False))
        self.attn_t, self.attn_v, self.attn_c = [], [], []
        if self.viz and train:
            raise ValueError("visualization exports the undropped attention probabilities: eval mode only")
        self.dyn = bool(getattr(self.cfg, "dynamic_attention", False))
        self.fast = bool(getattr(self.cfg, "fast_mode", False))
        self.Bt = 1 if self.fast else B
        if self.fast and (train or grad_outputs or vqa_loss or loss):
            raise ValueError("fast_mode is an inference path (text batch 1 broadcast to the image batch): no train mode / gradients")
        self.grad_outputs = frozenset(grad_outputs)
        # objective fused into the step (LOSS_HEADS): its scalar lands in self.loss (device) and its gradient goes straight into the
        # backward of the head(s) it reads; vqa_loss=True is the round-1 spelling of loss="vqa"
        self.loss_kind = "vqa" if vqa_loss else loss
        if self.loss_kind is not None and self.loss_kind not in LOSS_HEADS:
            raise ValueError(f"loss must be one of {sorted(LOSS_HEADS)}")
        self.vqa_loss = self.loss_kind == "vqa"
        self.train = bool(train)          # nn.Dropout layers active (model.train()); False = the reference's eval mode
        self.op_fp16 = 1 if engine.op_dtype == F16 else 0   # format of the forward operands (activations, weights)
        self.op_dtype, self.split = engine.op_dtype, engine.split
        self.head_dropout_prob = engine.head_dropout_prob
        self.heads = engine.ps.heads if heads is None else heads   # "vl" | "pretraining" | "none"
        self.fwd_id = 0
        self.fwd, self.bwd = [], []
        self.prologue = []       # optional per-step ops run before the forward (see enable_training_prologue)
        self.epilogue = []       # optional per-step ops run after the backward (the fused optimizer: enable_optimizer)
        self.cur = self.fwd
        self.sid = 0             # stream the next emitted op goes to: 0 = text/main stream, 1 = vision stream
        self.two_streams = engine.two_streams
        self.wgrad_streams = engine.wgrad_streams and engine.two_streams   # weight-gradient GEMMs off the critical chain
        self._streams = None
        self._n_events = 0
        self._scratch_epoch = 0
        self._keep = []          # ctypes structs / tensors referenced by raw pointer
        self._scratch = {}
        self._bwd_emitters = []
        self.n_kernels_fwd = self.n_kernels_bwd = 0
        self.graph_fwd = self.graph_bwd = self.graph_step = None
        self._eager_runs = [0, 0]      # eager forward / backward executions (maybe_capture_passes)
        self._arena_off = self.arena_bytes = 0
        self._build()

    # ------------------------------------------------------------------ infrastructure
    def buf(self, shape, dtype=F32, zero=False):
        """A static buffer of the plan. With the engine's shared activation arena enabled (Engine.enable_activation_arena), buffers
        that hold no state between runs (activations, scratch: everything a run writes before reading) are sub-allocated from the
        arena at the same offsets in every plan, so the plans of different shapes overlay each other; buffers that are
        initialised at build time or loaded from outside a run (zero=True: inputs, labels, output gradients, zero-padded
        operands) stay private."""
        arena = self.e.arena
        if arena is None or zero:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)
            self._keep.append(t)
            return t
        n = 1
        for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = self._arena_off
        if off + nbytes > arena.numel():
            raise L.VBError(f"activation arena of {arena.numel() / 2**30:.2f} GiB is too small for plan B={self.B} Nt={self.Nt} Nv={self.Nv} "
                            f"(needs more than {(off + nbytes) / 2**30:.2f} GiB): pass a larger size to Engine.enable_activation_arena")
        self._arena_off = (off + nbytes + 255) // 256 * 256
        self.arena_bytes = self._arena_off
        return arena[off:off + nbytes].view(dtype).view(shape)

    def buf16(self, shape, bw=True):
        """Forward-operand buffer in the engine's operand format: (hi, lo, bw) with lo = None unless split precision and
        bw = the bf16 copy the backward's weight-gradient GEMM reads (hi itself when the format is bf16, None if not wanted)."""
        hi = self.buf(shape, self.op_dtype)
        lo = self.buf(shape, self.op_dtype) if self.split else None
        b = None if not bw else (hi if self.op_dtype == BF16 else self.buf(shape, BF16))
        return hi, lo, b

    @staticmethod
    def _extra(bw, hi):
        """Pointer for a kernel's optional bf16-copy output: None when the copy IS the primary output."""
        return None if (bw is None or bw is hi) else bw

    def scratch(self, tag, shape, dtype):
        # with asynchronous weight-gradient streams a temporary may still be read after its layer's backward has moved on:
        # temporaries are then unique per backward emitter instead of being recycled by the next layer
        key = (tag, tuple(shape), dtype, self._scratch_epoch if self.wgrad_streams else 0)
        if key not in self._scratch:
            self._scratch[key] = self.buf(shape, dtype)
        return self._scratch[key]

    def emit(self, fn, *args):
        self.cur.append((fn, args, self.sid if self.two_streams else 0))

    def sync_streams(self, mirror=True):
        """Both streams wait for each other here. Between two connection layers the text and the vision segments are
        data-independent (vilbert.py:977-1006), so they run on two CUDA streams (and as parallel branches of the
        captured graph). mirror=True also places a barrier at the mirrored position of the backward pass."""
        if not self.two_streams:
            return
        self.cur.append((None, (), 0))
        if mirror and self.cur is self.fwd:
            self._bwd_emitters.append(None)

    class _On:
        def __init__(self, plan, sid):
            self.plan, self.sid = plan, sid

        def __enter__(self):
            self.prev = self.plan.sid
            self.plan.sid = self.sid

        def __exit__(self, *exc):
            self.plan.sid = self.prev
            return False

    def on(self, sid):
        return Plan._On(self, sid)

    def push_bwd(self, fn):
        """Registers a backward emitter; it will emit on the stream that is current now. Layers built while `self._no_grad` is set
        (the reference's `with torch.no_grad()` around the first fixed_t_layer / fixed_v_layer layers) register nothing."""
        if getattr(self, "_no_grad", False):
            return
        self._bwd_emitters.append((self.sid, fn))

    @staticmethod
    def _ptr(t):
        if t is None:
            return None
        return t.data_ptr() if torch.is_tensor(t) else int(t)

    def drop(self, name, p):
        """ctypes byref of a vb_dropout for the dropout layer `name` with probability p, or None when inactive."""
        if not self.train or p is None or p <= 0.0:
            return None
        d = L.Dropout()
        d.step, d.site, d.p = self.e.drop_step.data_ptr(), dropout_site_id(name), float(p)
        self._keep.append(d)
        return d

    @staticmethod
    def _ref(d):
        return C.byref(d) if d is not None else None

    def gemm(self, M, N, K, A, lda, B, ldb, a_mn=0, b_mn=0, bias=None, residual=None, ld_res=0, aux=None, ld_aux=0, act=0,
             out_f32=None, ld_of=0, out_bf16=None, ld_ob=0, out_pre=None, ld_op=0, atomic=0, split_k=1, alpha=1.0, out_colsum=None,
             dropout=None, a_lo=None, b_lo=None, out_lo=None, out_b16=None):
        """Operand formats follow the pass being emitted: in the forward pass A, B and the 16-bit output are forward
        operands (engine format, + low parts in split precision; out_b16 = optional bf16 copy of the output for the
        backward); in the backward pass everything is bf16: A is a gradient operand, B the bf16 copy of a weight (dgrad) or
        of a saved activation (wgrad), 16-bit outputs are gradients."""
        g = L.GemmArgs()
        fwd = self.cur is self.fwd
        g.a_fp16 = g.b_fp16 = g.out_fp16 = self.op_fp16 if fwd else 0
        if fwd:
            g.out_b16 = self._ptr(out_b16)
        if fwd and self.split:
            g.A_lo, g.B_lo, g.out_lo = self._ptr(a_lo), self._ptr(b_lo), self._ptr(out_lo)
        g.M, g.N, g.K = M, N, K
        g.A, g.lda, g.a_mn_major = self._ptr(A), lda, a_mn
        g.B, g.ldb, g.b_mn_major = self._ptr(B), ldb, b_mn
        g.alpha = alpha
        g.bias = self._ptr(bias)
        g.residual, g.ld_res = self._ptr(residual), ld_res
        g.aux, g.ld_aux = self._ptr(aux), ld_aux
        g.act = act
        g.out_f32, g.ld_out_f32 = self._ptr(out_f32), ld_of
        g.out_bf16, g.ld_out_bf16 = self._ptr(out_bf16), ld_ob
        g.out_pre, g.ld_out_pre = self._ptr(out_pre), ld_op
        g.atomic_out, g.split_k, g.block_n, g.max_ctas = atomic, split_k, 0, (0 if fwd else self.e.bwd_gemm_max_ctas)
        g.out_colsum = self._ptr(out_colsum)
        if dropout is not None:
            g.dropout = dropout
        self._keep.append(g)
        self.emit(self.lib.vb_gemm_bf16, C.byref(g))

    def attention(self, bwd, B, H, Nq, Nk, D, Q, ldq, K, ldk, V, ldv, mask, O, ldo, lse, dO=None, lddo=0, dQ=None, lddq=0,
                  dK=None, lddk=0, dV=None, lddv=0, delta=None, dbq=None, dbk=None, dbv=None, dropout=None, q_lo=None, k_lo=None,
                  v_lo=None, o_lo=None, o_b16=None):
        a = L.AttnArgs()
        a.qkv_fp16 = self.op_fp16
        a.O_b16 = self._ptr(o_b16)     # forward: written; backward: read for delta (consistent with the bf16 backward products)
        if not bwd and self.split:
            a.Q_lo, a.K_lo, a.V_lo, a.O_lo = self._ptr(q_lo), self._ptr(k_lo), self._ptr(v_lo), self._ptr(o_lo)
        a.B, a.H, a.Nq, a.Nk, a.D = B, H, Nq, Nk, D
        a.Q, a.ldq, a.K, a.ldk, a.V, a.ldv = self._ptr(Q), ldq, self._ptr(K), ldk, self._ptr(V), ldv
        a.mask, a.scale = self._ptr(mask), 1.0 / math.sqrt(D)
        a.O, a.ldo, a.lse = self._ptr(O), ldo, self._ptr(lse)
        a.dO, a.lddo, a.dQ, a.lddq = self._ptr(dO), lddo, self._ptr(dQ), lddq
        a.dK, a.lddk, a.dV, a.lddv, a.delta = self._ptr(dK), lddk, self._ptr(dV), lddv, self._ptr(delta)
        a.dbias_q, a.dbias_k, a.dbias_v = self._ptr(dbq), self._ptr(dbk), self._ptr(dbv)
        if dropout is not None:
            a.dropout = dropout
        self._keep.append(a)
        self.emit(self.lib.vb_attention_bwd if bwd else self.lib.vb_attention_fwd, C.byref(a))
        if not bwd and self.viz:
            # config.visualization: the probabilities (fp32 [B, heads, Nq, Nk]) and views of the queries / keys the reference returns
            probs = self.buf((B, H, Nq, Nk), F32)
            self.emit(self.lib.vb_attention_probs, C.byref(a), probs.data_ptr())
            self._last_attn = dict(attn=probs, q=Q, k=K, B=B, H=H, Nq=Nq, Nk=Nk, D=D)

    def ln_fwd(self, x, gamma, beta, M, H, want_f32=True, out_drop=None):
        y32 = self.buf((M, H), F32) if want_f32 else None
        y16, ylo, ybw = self.buf16((M, H))
        mean, rstd = self.buf((M,), F32), self.buf((M,), F32)
        self.emit(self.lib.vb_layernorm_fwd, x.data_ptr(), H, gamma.data_ptr(), beta.data_ptr(), 1e-12, self._ptr(y32), y16.data_ptr(), H,
                  mean.data_ptr(), rstd.data_ptr(), M, H, self._ref(out_drop), self.op_fp16, self._ptr(ylo), self._ptr(self._extra(ybw, y16)))
        return y32, y16, mean, rstd, ylo, ybw

    def ln_bwd(self, dy, x, gamma, mean, rstd, dx32, dx16, M, H, ggamma, gbeta, pre=None, gbias=None, out_drop=None, in_drop=None):
        """gbias: bias gradient of the Linear feeding this LayerNorm (column sums of dx), fused into the same pass."""
        self.emit(self.lib.vb_layernorm_bwd, dy.data_ptr(), H, x.data_ptr(), H, gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                  self._ptr(dx32), self._ptr(dx16), H, self._ptr(pre), H, ggamma.data_ptr(), gbeta.data_ptr(), self._ptr(gbias), M, H,
                  self._ref(out_drop), self._ref(in_drop))

    def colsum(self, X, ld, out, M, N):
        self.emit(self.lib.vb_colsum, X.data_ptr(), 1 if X.dtype == BF16 else 0, ld, out.data_ptr(), M, N)

    def grad_of(self, act):
        if act.g32 is None:
            act.g32 = self.buf((act.M, act.H), F32)
        return act.g32

    # dW, db of y = x W^T + b given dy (bf16 operand copy + a source for the bias column sums)
    def linear_wgrad(self, dy16, ld_dy, dy_bias, ld_dyb, x16, ld_x, M, N_out, K_in, wname, gw=None):
        """dW += dy^T x (split-K, atomics). Nothing on the critical chain depends on it, so with wgrad_streams it is issued on
        a side stream (2 = text chain, 3 = vision chain) right after an event marking that dy is ready; the side streams are
        joined at the data-parallel segment cuts and at the end of the backward pass."""
        if dy_bias is not None:
            self.colsum(dy_bias, ld_dyb, self.ps.g(wname + ".bias"), M, N_out)
        out = gw if gw is not None else self.ps.g(wname + ".weight")
        if not self.wgrad_streams:
            self.gemm(N_out, K_in, M, dy16, ld_dy, x16, ld_x, a_mn=1, b_mn=1, out_f32=out, ld_of=K_in, atomic=1, split_k=0)
            return
        ev = self._n_events
        self._n_events += 1
        chain = self.sid
        self.cur.append((None, ("rec", ev), chain))
        self.sid = 2 + chain
        self.cur.append((None, ("wait", ev), self.sid))
        self.gemm(N_out, K_in, M, dy16, ld_dy, x16, ld_x, a_mn=1, b_mn=1, out_f32=out, ld_of=K_in, atomic=1, split_k=0)
        self.sid = chain

    # act.g32 (+)= dy16 @ W (+ extra32)
    def dgrad_into(self, act, dy16, ld_dy, W16, M, N_out, K_in, extra32=None):
        if act.frozen:      # the producer ran under no_grad: the gradient stops here
            return
        g = self.grad_of(act)
        if not act.gw:
            self.gemm(M, K_in, N_out, dy16, ld_dy, W16, K_in, b_mn=1, residual=extra32, ld_res=K_in, out_f32=g, ld_of=K_in)
            act.gw = True
        else:
            if extra32 is not None:
                self.emit(self.lib.vb_axpy_f32, extra32.data_ptr(), g.data_ptr(), M * K_in, 1.0)
            self.gemm(M, K_in, N_out, dy16, ld_dy, W16, K_in, b_mn=1, residual=g, ld_res=K_in, out_f32=g, ld_of=K_in)

    def add_grad(self, act, src32):
        if act.frozen:
            return
        g = self.grad_of(act)
        if not act.gw:
            self.emit(self.lib.vb_memset_zero, g.data_ptr(), g.numel() * 4)
            act.gw = True
        self.emit(self.lib.vb_axpy_f32, src32.data_ptr(), g.data_ptr(), g.numel(), 1.0)

    # ------------------------------------------------------------------ blocks
    def dense_res_ln(self, a16, K_in, res, wname, lnname, tag, drop=None, a_lo=None, a_bw=None):
        """LN(dense(a) + residual)  — BertSelfOutput / BertOutput / BertBiOutput halves (vilbert.py:470-474, 513-517, 844-855)."""
        ps, M, H = self.ps, res.M, res.H
        y = self.buf((M, H), F32)
        self.gemm(M, H, K_in, a16, K_in, ps.w16(wname + ".weight"), K_in, bias=ps.p(wname + ".bias"), residual=res.f32, ld_res=H,
                  out_f32=y, ld_of=H, dropout=drop, a_lo=a_lo, b_lo=ps.w16lo(wname + ".weight"))
        o32, o16, mean, rstd, olo, obw = self.ln_fwd(y, ps.p(lnname + ".weight"), ps.p(lnname + ".bias"), M, H)
        out = Act(o32, o16, M, H, lo=olo, bw=obw)
        a_bw = a16 if a_bw is None else a_bw

        def bwd():
            """returns (dy16, dy32) of the dense output (== grad of the LN input); adds dy32 to res."""
            if not out.gw:
                return None
            dy32 = self.scratch(tag + ".dy32", (M, H), F32)
            dy16 = self.scratch(tag + ".dy16", (M, H), BF16)
            self.ln_bwd(out.g32, y, ps.p(lnname + ".weight"), mean, rstd, dy32, dy16, M, H, ps.g(lnname + ".weight"), ps.g(lnname + ".bias"),
                        gbias=ps.g(wname + ".bias"), in_drop=drop)
            self.linear_wgrad(dy16, H, None, 0, a_bw, K_in, M, H, K_in, wname)
            return dy16, dy32
        return out, bwd

    def ffn(self, x, I, w1, w2, lnname, tag, drop=None):
        """LN(dense2(gelu(dense1(x))) + x) — BertIntermediate + BertOutput (vilbert.py:500-503, 513-517)."""
        ps, M, H = self.ps, x.M, x.H
        pre16 = self.buf((M, I), BF16)          # gelu'(pre), bf16 in every mode (only the backward reads it)
        f16, flo, fbw = self.buf16((M, I))
        self.gemm(M, I, H, x.b16, H, ps.w16(w1 + ".weight"), H, bias=ps.p(w1 + ".bias"), act=L.VB_ACT_GELU, out_bf16=f16, ld_ob=I,
                  out_pre=pre16, ld_op=I, a_lo=x.lo, b_lo=ps.w16lo(w1 + ".weight"), out_lo=flo, out_b16=self._extra(fbw, f16))
        out, out_bwd = self.dense_res_ln(f16, I, x, w2, lnname, tag + ".o", drop=drop, a_lo=flo, a_bw=fbw)

        def bwd():
            r = out_bwd()
            if r is None:
                return
            dy16, dy32 = r
            dpre16 = self.scratch(tag + ".dpre16", (M, I), BF16)
            # d pre = (dy W2) * gelu'(pre)
            self.gemm(M, I, H, dy16, H, ps.w16b(w2 + ".weight"), I, b_mn=1, aux=pre16, ld_aux=I, act=L.VB_ACT_DGELU, out_bf16=dpre16, ld_ob=I,
                      out_colsum=ps.g(w1 + ".bias"))
            self.linear_wgrad(dpre16, I, None, 0, x.bw, H, M, I, H, w1)
            self.dgrad_into(x, dpre16, I, ps.w16b(w1 + ".weight"), M, I, H, extra32=dy32)
        self.push_bwd(bwd)
        return out

    def self_attention_block(self, x, B, N, nh, mask, prefix, tag, p_attn=0.0, p_hidden=0.0, pool=None):
        """BertAttention (self-attention + output), text or image stream (vilbert.py:424-474, 571-633). pool: the pooled text states
        (text_pool) when config.dynamic_attention gates this image layer's queries and keys (:577-586)."""
        ps, M, H = self.ps, x.M, x.H
        D = H // nh
        qkv, qkvl, _ = self.buf16((M, 3 * H), bw=False)     # the attention backward converts its Q/K/V panels in shared memory
        self.gemm(M, 3 * H, H, x.b16, H, ps.w16(prefix + ".self.qkv.weight"), H, bias=ps.p(prefix + ".self.qkv.bias"), out_bf16=qkv, ld_ob=3 * H,
                  a_lo=x.lo, b_lo=ps.w16lo(prefix + ".self.qkv.weight"), out_lo=qkvl)
        if pool is not None:
            # z = dyLinear_q | dyLinear_k (pool) as one [B, 2H] GEMM; Q and K are scaled in place by 1 + sigmoid(z)
            Kp = pool.H
            z = self.buf((B, 2 * H), F32)
            self.gemm(B, 2 * H, Kp, pool.b16, Kp, ps.w16(prefix + ".self.dy.weight"), Kp, bias=ps.p(prefix + ".self.dy.bias"), out_f32=z, ld_of=2 * H,
                      a_lo=pool.lo, b_lo=ps.w16lo(prefix + ".self.dy.weight"))
            self.emit(self.lib.vb_gate_scale_fwd, qkv.data_ptr(), self._ptr(qkvl), 3 * H, z.data_ptr(), B, N, 2 * H, self.op_fp16)
        ctx, ctxl, ctxb = self.buf16((M, H))
        lse = self.buf((B, nh, N), F32)
        q, k, v = qkv[:, 0:H], qkv[:, H:2 * H], qkv[:, 2 * H:]
        ql, kl, vl = (qkvl[:, 0:H], qkvl[:, H:2 * H], qkvl[:, 2 * H:]) if self.split else (None, None, None)
        adrop = self.drop(prefix + ".self.dropout", p_attn)
        self.attention(False, B, nh, N, N, D, q, 3 * H, k, 3 * H, v, 3 * H, mask, ctx, H, lse, dropout=adrop, q_lo=ql, k_lo=kl, v_lo=vl, o_lo=ctxl,
                       o_b16=self._extra(ctxb, ctx))
        if self.viz:
            (self.attn_t if tag == "t" else self.attn_v).append(self._last_attn)
        out, out_bwd = self.dense_res_ln(ctx, H, x, prefix + ".output.dense", prefix + ".output.LayerNorm", tag + ".ao",
                                         drop=self.drop(prefix + ".output.dropout", p_hidden), a_lo=ctxl, a_bw=ctxb)

        def bwd():
            r = out_bwd()
            if r is None:
                return
            dy16, dy32 = r
            dctx = self.scratch(tag + ".dctx", (M, H), BF16)
            self.gemm(M, H, H, dy16, H, ps.w16b(prefix + ".output.dense.weight"), H, b_mn=1, out_bf16=dctx, ld_ob=H)
            dqkv = self.scratch(tag + ".dqkv", (M, 3 * H), BF16)
            delta = self.scratch(tag + ".delta", (B, nh, N), F32)
            gb = ps.g(prefix + ".self.qkv.bias")     # bias gradients = column sums of dQ|dK|dV, fused into the attention backward
            gated = pool is not None                 # ... except under the gate, where the biases sit before the scaling
            self.attention(True, B, nh, N, N, D, q, 3 * H, k, 3 * H, v, 3 * H, mask, ctx, H, lse, dO=dctx, lddo=H,
                           dQ=dqkv[:, 0:H], lddq=3 * H, dK=dqkv[:, H:2 * H], lddk=3 * H, dV=dqkv[:, 2 * H:], lddv=3 * H, delta=delta,
                           dbq=None if gated else gb[0:H], dbk=None if gated else gb[H:2 * H], dbv=gb[2 * H:], dropout=adrop,
                           o_b16=self._extra(ctxb, ctx))
            if gated:
                dz32 = self.scratch(tag + ".dz32", (B, 2 * H), F32)
                dz16 = self.scratch(tag + ".dz16", (B, 2 * H), BF16)
                self.emit(self.lib.vb_gate_scale_bwd, dqkv.data_ptr(), 3 * H, qkv.data_ptr(), self._ptr(qkvl), 3 * H, z.data_ptr(), dz32.data_ptr(),
                          dz16.data_ptr(), B, N, 2 * H, self.op_fp16)
                self.colsum(dqkv, 3 * H, gb[0:2 * H], M, 2 * H)
                self.linear_wgrad(dz16, 2 * H, dz32, 2 * H, pool.bw, pool.H, B, 2 * H, pool.H, prefix + ".self.dy")
                self.dgrad_into(pool, dz16, 2 * H, ps.w16b(prefix + ".self.dy.weight"), B, 2 * H, pool.H)
            self.linear_wgrad(dqkv, 3 * H, None, 0, x.bw, H, M, 3 * H, H, prefix + ".self.qkv")
            self.dgrad_into(x, dqkv, 3 * H, ps.w16b(prefix + ".self.qkv.weight"), M, 3 * H, H, extra32=dy32)
        self.push_bwd(bwd)
        return out

    def connection_layer(self, v, t, idx):
        """BertConnectionLayer.forward (vilbert.py:871-900): co-attention both ways + dual FFN."""
        ps, c, B = self.ps, self.cfg, self.B
        p = f"bert.encoder.c_layer.{idx}"
        Hb, nh = c.bi_hidden_size, c.bi_num_attention_heads
        D = Hb // nh
        Mv, Mt, Hv, Ht, Nv, Nt = v.M, t.M, v.H, t.H, self.Nv, self.Nt
        qkv1, qkv1l, _ = self.buf16((Mv, 3 * Hb), bw=False)
        qkv2, qkv2l, _ = self.buf16((Mt, 3 * Hb), bw=False)
        with self.on(1):
            self.gemm(Mv, 3 * Hb, Hv, v.b16, Hv, ps.w16(p + ".biattention.qkv1.weight"), Hv, bias=ps.p(p + ".biattention.qkv1.bias"), out_bf16=qkv1, ld_ob=3 * Hb,
                      a_lo=v.lo, b_lo=ps.w16lo(p + ".biattention.qkv1.weight"), out_lo=qkv1l)
        self.gemm(Mt, 3 * Hb, Ht, t.b16, Ht, ps.w16(p + ".biattention.qkv2.weight"), Ht, bias=ps.p(p + ".biattention.qkv2.bias"), out_bf16=qkv2, ld_ob=3 * Hb,
                  a_lo=t.lo, b_lo=ps.w16lo(p + ".biattention.qkv2.weight"), out_lo=qkv2l)
        self.sync_streams()      # each direction needs the other stream's keys / values
        q1, k1, v1 = qkv1[:, 0:Hb], qkv1[:, Hb:2 * Hb], qkv1[:, 2 * Hb:]
        q2, k2, v2 = qkv2[:, 0:Hb], qkv2[:, Hb:2 * Hb], qkv2[:, 2 * Hb:]
        ctx1, ctx1l, ctx1b = self.buf16((Mt, Hb)); lse1 = self.buf((B, nh, Nt), F32)   # text queries over vision keys/values
        ctx2, ctx2l, ctx2b = self.buf16((Mv, Hb)); lse2 = self.buf((B, nh, Nv), F32)   # vision queries over text keys/values
        L3 = 3 * Hb
        if self.split:
            q1l, k1l, v1l = qkv1l[:, 0:Hb], qkv1l[:, Hb:2 * Hb], qkv1l[:, 2 * Hb:]
            q2l, k2l, v2l = qkv2l[:, 0:Hb], qkv2l[:, Hb:2 * Hb], qkv2l[:, 2 * Hb:]
        else:
            q1l = k1l = v1l = q2l = k2l = v2l = None
        # dropout1 acts on attention_probs1 (text queries over regions), dropout2 on attention_probs2 (vilbert.py:730, 738, 778, 800)
        adrop1 = self.drop(p + ".biattention.dropout1", c.v_attention_probs_dropout_prob)
        adrop2 = self.drop(p + ".biattention.dropout2", c.attention_probs_dropout_prob)
        self.attention(False, B, nh, Nt, Nv, D, q2, L3, k1, L3, v1, L3, self.mask_v, ctx1, Hb, lse1, dropout=adrop1,
                       q_lo=q2l, k_lo=k1l, v_lo=v1l, o_lo=ctx1l, o_b16=self._extra(ctx1b, ctx1))
        a1 = self._last_attn if self.viz else None
        with self.on(1):
            self.attention(False, B, nh, Nv, Nt, D, q1, L3, k2, L3, v2, L3, self.mask_t, ctx2, Hb, lse2, dropout=adrop2,
                           q_lo=q1l, k_lo=k2l, v_lo=v2l, o_lo=ctx2l, o_b16=self._extra(ctx2b, ctx2))
            a2 = self._last_attn if self.viz else None
        if self.viz:
            self.attn_c.append((a1, a2))
        # biOutput: ctx2 -> vision stream (dense1 / LayerNorm1), ctx1 -> text stream (dense2 / LayerNorm2) (:890-892)
        with self.on(1):
            v1o, v1_bwd = self.dense_res_ln(ctx2, Hb, v, p + ".biOutput.dense1", p + ".biOutput.LayerNorm1", "c.v.bo",
                                            drop=self.drop(p + ".biOutput.dropout1", c.v_hidden_dropout_prob), a_lo=ctx2l, a_bw=ctx2b)
        t1o, t1_bwd = self.dense_res_ln(ctx1, Hb, t, p + ".biOutput.dense2", p + ".biOutput.LayerNorm2", "c.t.bo",
                                        drop=self.drop(p + ".biOutput.dropout2", c.hidden_dropout_prob), a_lo=ctx1l, a_bw=ctx1b)

        def bwd():
            if not (v1o.gw or t1o.gw):
                return
            for a in (v1o, t1o):   # a stream without downstream gradient contributes zeros
                if not a.gw:
                    g = self.grad_of(a)
                    self.emit(self.lib.vb_memset_zero, g.data_ptr(), g.numel() * 4)
                    a.gw = True
            rv, rt = v1_bwd(), t1_bwd()
            dqkv1 = self.scratch("c.dqkv1", (Mv, L3), BF16)
            dqkv2 = self.scratch("c.dqkv2", (Mt, L3), BF16)
            dctx2 = self.scratch("c.dctx2", (Mv, Hb), BF16)
            dctx1 = self.scratch("c.dctx1", (Mt, Hb), BF16)
            dyv16, dyv32 = rv
            dyt16, dyt32 = rt
            self.gemm(Mv, Hb, Hv, dyv16, Hv, ps.w16b(p + ".biOutput.dense1.weight"), Hb, b_mn=1, out_bf16=dctx2, ld_ob=Hb)
            self.gemm(Mt, Hb, Ht, dyt16, Ht, ps.w16b(p + ".biOutput.dense2.weight"), Hb, b_mn=1, out_bf16=dctx1, ld_ob=Hb)
            gb1, gb2 = ps.g(p + ".biattention.qkv1.bias"), ps.g(p + ".biattention.qkv2.bias")
            d1 = self.scratch("c.delta1", (B, nh, Nt), F32)
            d2 = self.scratch("c.delta2", (B, nh, Nv), F32)
            self.attention(True, B, nh, Nt, Nv, D, q2, L3, k1, L3, v1, L3, self.mask_v, ctx1, Hb, lse1, dO=dctx1, lddo=Hb,
                           dQ=dqkv2[:, 0:Hb], lddq=L3, dK=dqkv1[:, Hb:2 * Hb], lddk=L3, dV=dqkv1[:, 2 * Hb:], lddv=L3, delta=d1,
                           dbq=gb2[0:Hb], dbk=gb1[Hb:2 * Hb], dbv=gb1[2 * Hb:], dropout=adrop1, o_b16=self._extra(ctx1b, ctx1))
            self.attention(True, B, nh, Nv, Nt, D, q1, L3, k2, L3, v2, L3, self.mask_t, ctx2, Hb, lse2, dO=dctx2, lddo=Hb,
                           dQ=dqkv1[:, 0:Hb], lddq=L3, dK=dqkv2[:, Hb:2 * Hb], lddk=L3, dV=dqkv2[:, 2 * Hb:], lddv=L3, delta=d2,
                           dbq=gb1[0:Hb], dbk=gb2[Hb:2 * Hb], dbv=gb2[2 * Hb:], dropout=adrop2, o_b16=self._extra(ctx2b, ctx2))
            self.linear_wgrad(dqkv1, L3, None, 0, v.bw, Hv, Mv, L3, Hv, p + ".biattention.qkv1")
            self.linear_wgrad(dqkv2, L3, None, 0, t.bw, Ht, Mt, L3, Ht, p + ".biattention.qkv2")
            self.dgrad_into(v, dqkv1, L3, ps.w16b(p + ".biattention.qkv1.weight"), Mv, L3, Hv, extra32=dyv32)
            self.dgrad_into(t, dqkv2, L3, ps.w16b(p + ".biattention.qkv2.weight"), Mt, L3, Ht, extra32=dyt32)
        # the cross-modal backward touches both streams' tensors: it runs on the main stream between two barriers
        self._bwd_emitters.append(None)
        self.push_bwd(bwd)
        self._bwd_emitters.append(None)
        with self.on(1):
            v2o = self.ffn(v1o, c.v_intermediate_size, p + ".v_intermediate.dense", p + ".v_output.dense", p + ".v_output.LayerNorm", "c.v.ffn",
                           drop=self.drop(p + ".v_output.dropout", c.v_hidden_dropout_prob))
        t2o = self.ffn(t1o, c.intermediate_size, p + ".t_intermediate.dense", p + ".t_output.dense", p + ".t_output.LayerNorm", "c.t.ffn",
                       drop=self.drop(p + ".t_output.dropout", c.hidden_dropout_prob))
        return v2o, t2o

    def broadcast_text(self, t):
        """FAST_MODE: t [1*Nt, H] -> [B*Nt, H] (fp32 values and operand copies), and the text mask [1, Nt] -> [B, Nt]."""
        B, M1, H = self.B, t.M, t.H
        f32 = self.buf((B * M1, H), F32)
        b16, lo, _ = self.buf16((B * M1, H), bw=False)
        self.emit(self.lib.vb_broadcast_rows, t.f32.data_ptr(), f32.data_ptr(), M1 * H * 4, B)
        self.emit(self.lib.vb_broadcast_rows, t.b16.data_ptr(), b16.data_ptr(), M1 * H * 2, B)
        if lo is not None:
            self.emit(self.lib.vb_broadcast_rows, t.lo.data_ptr(), lo.data_ptr(), M1 * H * 2, B)
        nt4 = (self.Nt * 4 + 15) // 16 * 16           # the mask row is padded to 16 bytes for the broadcast kernel
        if nt4 == self.Nt * 4:
            m = self.buf((B, self.Nt), F32)
            self.emit(self.lib.vb_broadcast_rows, self.mask_t.data_ptr(), m.data_ptr(), self.Nt * 4, B)
        else:
            m = self.mask_t.new_empty((B, self.Nt)); self._keep.append(m)
            self.emit(self.lib.vb_mask_to_additive, self.in_amask_b.data_ptr(), m.data_ptr(), B, self.Nt_in, 1 if self.has_task else 0)
        self.mask_t = m
        return Act(f32, b16, B * M1, H, lo=lo)

    def expand_pairs(self, t, v):
        """in_batch_pairs (vilbert.py:1008-1040): sample p = i * b + j of the expanded batch pairs text i with image j —
        txt.unsqueeze(1).expand(b, b, ...) (every text repeated b times) and img.unsqueeze(0).expand(b, b, ...) (the image batch
        tiled b times). Backward: the gradient of an item is the sum over its b copies (vb_sum_strided)."""
        b, lib = self.Bin, self.lib
        outs = []
        self.sync_streams()      # the image stream's tensors are expanded on the main stream
        for act, N, is_text in ((t, self.Nt, True), (v, self.Nv, False)):
            n = N * act.H
            f32 = self.buf((b * b * N, act.H), F32)
            b16, lo, bw = self.buf16((b * b * N, act.H))
            bufs = [(act.f32, f32, 4), (act.b16, b16, 2)] + ([(act.lo, lo, 2)] if lo is not None else []) + ([(act.bw, bw, 2)] if bw is not b16 else [])
            for src, dst, sz in bufs:
                if is_text:
                    self.emit(lib.vb_repeat_rows, src.data_ptr(), dst.data_ptr(), n * sz, b, b)
                else:
                    self.emit(lib.vb_broadcast_rows, src.data_ptr(), dst.data_ptr(), b * n * sz, b)
            out = Act(f32, b16, b * b * N, act.H, lo=lo, bw=bw)
            outs.append(out)

            def bwd(act=act, out=out, n=n, is_text=is_text):
                if not out.gw or act.frozen:
                    return
                g = self.grad_of(act)
                acc = 1 if act.gw else 0
                if is_text:   # g[i] = sum_j out.g32[i * b + j]
                    self.emit(lib.vb_sum_strided, out.g32.data_ptr(), g.data_ptr(), n, b, b * n, b, n, acc)
                else:         # g[j] = sum_i out.g32[i * b + j]
                    self.emit(lib.vb_sum_strided, out.g32.data_ptr(), g.data_ptr(), n, b, n, b, b * n, acc)
                act.gw = True
            self._bwd_emitters.append(None)
            self.push_bwd(bwd)
            self._bwd_emitters.append(None)
        # masks: text mask rows repeated, image mask tiled (4-byte rows: plain torch-free kernels need 16-byte items -> host-side views)
        mt = self.buf((b * b, self.Nt), F32); mv = self.buf((b * b, self.Nv), F32)
        self._pair_masks = (self.mask_t, self.mask_v, mt, mv)
        self.emit(lib.vb_mask_to_additive, self.in_amask_pairs.data_ptr(), mt.data_ptr(), b * b, self.Nt_in, 1 if self.has_task else 0)
        self.emit(lib.vb_mask_to_additive, self.in_imask_pairs.data_ptr(), mv.data_ptr(), b * b, self.Nv, 0)
        self.mask_t, self.mask_v = mt, mv
        self.sync_streams()
        return outs[0], outs[1]

    def text_layer(self, x, i):
        p = f"bert.encoder.layer.{i}"
        c = self.cfg
        h1 = self.self_attention_block(x, x.M // self.Nt, self.Nt, c.num_attention_heads, self.mask_t, p + ".attention", "t",
                                       p_attn=c.attention_probs_dropout_prob, p_hidden=c.hidden_dropout_prob)
        return self.ffn(h1, c.intermediate_size, p + ".intermediate.dense", p + ".output.dense", p + ".output.LayerNorm", "t.ffn",
                        drop=self.drop(p + ".output.dropout", c.hidden_dropout_prob))

    def text_pool(self, t):
        """dynamic_attention: masked mean of the text states over the tokens (vilbert.py:578-579) as an Act [B, Ht] with GEMM operand
        copies; its backward adds d pool to the gradient of the text states. Emitted on the text stream; the caller places a
        barrier before the image layers that read it (their backward accumulates d pool, mirrored barrier, then this backward)."""
        B, Ht, lib = self.B, t.H, self.lib
        p32 = self.buf((B, Ht), F32)
        p16, plo, pbw = self.buf16((B, Ht))
        self.emit(lib.vb_masked_mean_fwd, t.f32.data_ptr(), self.mask_t.data_ptr(), p32.data_ptr(), p16.data_ptr(), self._ptr(plo),
                  self._ptr(self._extra(pbw, p16)), self.op_fp16, B, self.Nt, Ht)
        pool = Act(p32, p16, B, Ht, lo=plo, bw=pbw)
        mask = self.mask_t

        def bwd():
            if not pool.gw or t.frozen:
                return
            g = self.grad_of(t)
            self.emit(lib.vb_masked_mean_bwd, pool.g32.data_ptr(), mask.data_ptr(), g.data_ptr(), 1 if t.gw else 0, B, self.Nt, Ht)
            t.gw = True
        self.push_bwd(bwd)
        return pool

    def image_layer(self, x, i, pool=None):
        p = f"bert.encoder.v_layer.{i}"
        c = self.cfg
        h1 = self.self_attention_block(x, x.M // self.Nv, self.Nv, c.v_num_attention_heads, self.mask_v, p + ".attention", "v",
                                       p_attn=c.v_attention_probs_dropout_prob, p_hidden=c.v_hidden_dropout_prob, pool=pool)
        return self.ffn(h1, c.v_intermediate_size, p + ".intermediate.dense", p + ".output.dense", p + ".output.LayerNorm", "v.ffn",
                        drop=self.drop(p + ".output.dropout", c.v_hidden_dropout_prob))

    # ------------------------------------------------------------------ embeddings
    def embeddings(self):
        ps, c, B, lib = self.ps, self.cfg, self.Bin, self.lib     # the embeddings run at the INPUT batch
        Ht, Hv, Nt, Nv, Fv = c.hidden_size, c.v_hidden_size, self.Nt, self.Nv, c.v_feature_size
        Bt = self.Bt
        Mt, Mv = Bt * Nt, B * Nv
        # static inputs (the text side has batch 1 in FAST_MODE)
        self.in_ids = self.buf((Bt, self.Nt_in), I64, zero=True)
        self.in_tt = self.buf((Bt, self.Nt_in), I64, zero=True)
        self.in_task = self.buf((Bt,), I64, zero=True) if self.has_task else None
        self.in_amask = self.buf((Bt, self.Nt_in), I64, zero=True)
        self.in_amask_b = self.in_amask.expand(B, self.Nt_in).contiguous() if self.fast else self.in_amask   # refreshed in load_inputs
        if self.pairs:   # 0/1 masks of the expanded batch (text rows repeated, image rows tiled); refreshed in load_inputs
            self.in_amask_pairs = self.buf((B * B, self.Nt_in), I64, zero=True)
            self.in_imask_pairs = self.buf((B * B, Nv), I64, zero=True)
        self.in_imask = self.buf((B, Nv), I64, zero=True)
        self.in_feat = self.buf((B, Nv, Fv), F32, zero=True)
        self.in_loc = self.buf((B, Nv, 5), F32, zero=True)
        self.mask_t = self.buf((Bt, Nt), F32)
        self.mask_v = self.buf((B, Nv), F32)
        self.emit(lib.vb_mask_to_additive, self.in_amask.data_ptr(), self.mask_t.data_ptr(), Bt, self.Nt_in, 1 if self.has_task else 0)
        self.emit(lib.vb_mask_to_additive, self.in_imask.data_ptr(), self.mask_v.data_ptr(), B, Nv, 0)
        self.sync_streams()
        # text: gather-sum (+ task row) then LayerNorm (vilbert.py:346-367)
        xe = self.buf((Mt, Ht), F32)
        e = "bert.embeddings"
        self.emit(lib.vb_embed_text_fwd, self.in_ids.data_ptr(), self.in_tt.data_ptr(), self._ptr(self.in_task), ps.p(e + ".word_embeddings.weight").data_ptr(),
                  ps.p(e + ".position_embeddings.weight").data_ptr(), ps.p(e + ".token_type_embeddings.weight").data_ptr(),
                  ps.p(e + ".task_embeddings.weight").data_ptr() if self.has_task else None, xe.data_ptr(), Bt, self.Nt_in, Ht)
        tdrop = self.drop(e + ".dropout", c.hidden_dropout_prob)
        t32, t16, tmean, trstd, tlo, tbw = self.ln_fwd(xe, ps.p(e + ".LayerNorm.weight"), ps.p(e + ".LayerNorm.bias"), Mt, Ht, out_drop=tdrop)
        t = Act(t32, t16, Mt, Ht, lo=tlo, bw=tbw)

        def bwd_text():
            if t.gw:
                dxe = self.scratch("emb.dxe", (Mt, Ht), F32)
                self.ln_bwd(t.g32, xe, ps.p(e + ".LayerNorm.weight"), tmean, trstd, dxe, None, Mt, Ht, ps.g(e + ".LayerNorm.weight"), ps.g(e + ".LayerNorm.bias"),
                            out_drop=tdrop)
                self.emit(lib.vb_embed_text_bwd, dxe.data_ptr(), self.in_ids.data_ptr(), self.in_tt.data_ptr(), self._ptr(self.in_task),
                          ps.g(e + ".word_embeddings.weight").data_ptr(), ps.g(e + ".position_embeddings.weight").data_ptr(),
                          ps.g(e + ".token_type_embeddings.weight").data_ptr(),
                          ps.g(e + ".task_embeddings.weight").data_ptr() if self.has_task else None, B, self.Nt_in, Ht)
        self.push_bwd(bwd_text)
        # image: region features fp32 -> bf16 ingest, 2048 -> Hv GEMM with the 5 -> Hv box projection as residual, LayerNorm (:1421-1432)
        ve = "bert.v_embeddings"
        with self.on(1):
            feat16, featl, featb = self.buf16((Mv, Fv))
            self.emit(lib.vb_cast_f32_to_bf16, self.in_feat.data_ptr(), feat16.data_ptr(), Mv * Fv, self.op_fp16, self._ptr(featl),
                      self._ptr(self._extra(featb, feat16)))
            locp = self.buf((Mv, Hv), F32)
            self.emit(lib.vb_loc_proj_fwd, self.in_loc.data_ptr(), ps.p(ve + ".image_location_embeddings.weight").data_ptr(),
                      ps.p(ve + ".image_location_embeddings.bias").data_ptr(), locp.data_ptr(), Mv, Hv)
            yv = self.buf((Mv, Hv), F32)
            self.gemm(Mv, Hv, Fv, feat16, Fv, ps.w16(ve + ".image_embeddings.weight"), Fv, bias=ps.p(ve + ".image_embeddings.bias"),
                      residual=locp, ld_res=Hv, out_f32=yv, ld_of=Hv, a_lo=featl, b_lo=ps.w16lo(ve + ".image_embeddings.weight"))
            vdrop = self.drop(ve + ".dropout", c.hidden_dropout_prob)     # BertImageEmbeddings uses hidden_dropout_prob (vilbert.py:1419)
            v32, v16, vmean, vrstd, vlo, vbw = self.ln_fwd(yv, ps.p(ve + ".LayerNorm.weight"), ps.p(ve + ".LayerNorm.bias"), Mv, Hv, out_drop=vdrop)
            v = Act(v32, v16, Mv, Hv, lo=vlo, bw=vbw)

            def bwd_image():
                if v.gw:
                    dyv32 = self.scratch("emb.dyv32", (Mv, Hv), F32)
                    dyv16 = self.scratch("emb.dyv16", (Mv, Hv), BF16)
                    self.ln_bwd(v.g32, yv, ps.p(ve + ".LayerNorm.weight"), vmean, vrstd, dyv32, dyv16, Mv, Hv, ps.g(ve + ".LayerNorm.weight"), ps.g(ve + ".LayerNorm.bias"),
                                gbias=ps.g(ve + ".image_embeddings.bias"), out_drop=vdrop)
                    self.linear_wgrad(dyv16, Hv, None, 0, featb, Fv, Mv, Hv, Fv, ve + ".image_embeddings")
                    self.emit(lib.vb_loc_proj_bwd, dyv32.data_ptr(), self.in_loc.data_ptr(), ps.g(ve + ".image_location_embeddings.weight").data_ptr(),
                              ps.g(ve + ".image_location_embeddings.bias").data_ptr(), Mv, Hv)
            self.push_bwd(bwd_image)
        return t, v

    # ------------------------------------------------------------------ poolers and heads
    def pooler(self, seq, N, wname):
        """Linear + ReLU on token 0 (vilbert.py:1116-1122, 1131-1137); A is read with row pitch N*H."""
        ps, B, H, Hb = self.ps, self.B, seq.H, self.cfg.bi_hidden_size
        p32 = self.buf((B, Hb), F32)
        p16, plo, _ = self.buf16((B, Hb), bw=False)
        self.gemm(B, Hb, H, seq.b16, N * H, ps.w16(wname + ".weight"), H, bias=ps.p(wname + ".bias"), act=L.VB_ACT_RELU, out_f32=p32, ld_of=Hb,
                  out_bf16=p16, ld_ob=Hb, a_lo=seq.lo, b_lo=ps.w16lo(wname + ".weight"), out_lo=plo)
        pooled = Act(p32, p16, B, Hb, lo=plo)

        def bwd():
            if not pooled.gw:
                return
            dpre = self.scratch("pool.dpre", (B, Hb), BF16)
            dpre32 = self.scratch("pool.dpre32", (B, Hb), F32)
            self.emit(self.lib.vb_relu_bwd, pooled.g32.data_ptr(), p32.data_ptr(), dpre.data_ptr(), dpre32.data_ptr(), B * Hb)
            self.linear_wgrad(dpre, Hb, dpre32, Hb, seq.bw, N * H, B, Hb, H, wname)
            g = self.grad_of(seq)
            if not seq.gw:
                self.emit(self.lib.vb_memset_zero, g.data_ptr(), g.numel() * 4)
                seq.gw = True
            # rows b*N of the sequence gradient += dpre @ W
            self.gemm(B, H, Hb, dpre, Hb, ps.w16b(wname + ".weight"), H, b_mn=1, residual=g, ld_res=N * H, out_f32=g, ld_of=N * H)
        self.push_bwd(bwd)
        return pooled

    def out_grad_buffer(self, name, shape):
        """Static fp32 buffer the caller's d(loss)/d(output) is copied into before backward."""
        if name not in self.gout:
            self.gout[name] = self.buf(shape, F32, zero=True)
        return self.gout[name]

    def big_head(self, name, x16, ld_x, x_act, M, K_in, N_out, wname, bias_name, w16=None, gw=None, w16lo=None, w16b=None):
        """Wide linear head (N_out in the thousands): logits = x W^T + b as fp32 [M, N_out]; backward from a
        caller-supplied fp32 d(logits) (cast to a bf16 operand with an 8-padded row pitch)."""
        ps = self.ps
        W16 = w16 if w16 is not None else ps.w16(wname + ".weight")
        W16lo = w16lo if w16 is not None else ps.w16lo(wname + ".weight")
        W16b = w16b if w16 is not None else ps.w16b(wname + ".weight")
        logits = self.buf((M, N_out), F32)
        self.gemm(M, N_out, K_in, x16, ld_x, W16, K_in, bias=ps.p(bias_name), out_f32=logits, ld_of=N_out, a_lo=x_act.lo, b_lo=W16lo)
        self.outputs[name] = logits

        def bwd():
            if name not in self.grad_outputs:
                return None
            ldp = _pad8(N_out)
            if self.vqa_loss and name == "vil_prediction":
                dl32, dl16 = self.vqa_dl32, self.vqa_dl16
            else:
                dl32 = self.out_grad_buffer(name, (M, N_out))
                dl16 = self.scratch("head.dl16." + name, (M, ldp), BF16)
                self.emit(self.lib.vb_cast2d_f32_to_bf16, dl32.data_ptr(), N_out, dl16.data_ptr(), ldp, M, N_out, 1.0)
            self.colsum(dl32, N_out, ps.g(bias_name), M, N_out)
            self.linear_wgrad(dl16, ldp, None, 0, x_act.bw, ld_x, M, N_out, K_in, wname, gw=gw)
            return dl16, ldp, W16b
        return bwd

    def lm_head_compact(self, ht, ht_bwd):
        """Masked-LM head of the fused pre-training objective without the [tokens, vocab] logits: the rows with a label
        (masked_lm_labels != -1, 15 % of the tokens; vilbert.py:1578-1583) are compacted on the device into a fixed-capacity
        operand (engine.lm_capacity of the rows), the tied decoder GEMM, the cross-entropy and both backward GEMMs run on those
        rows, and the gradient is scattered back to the token rows. More labelled rows than the capacity poison the loss (NaN)."""
        ps, c, lib = self.ps, self.cfg, self.lib
        M, Ht, V = ht.M, ht.H, c.vocab_size
        cap = min(_pad8(M), _pad8(max(64, int(math.ceil(self.e.lm_capacity * M)))))
        labels = self.buf((M,), I64, zero=True)      # a plan input (loaded from outside a run): private, never in the shared arena
        labels.fill_(-1)
        self.loss_inputs["masked_lm_labels"] = labels
        idx, cnt, lab_c = self.buf((cap,), torch.int32), self.buf((1,), torch.int32, zero=True), self.buf((cap,), I64)
        self.emit(lib.vb_compact_rows, labels.data_ptr(), -1, M, cap, idx.data_ptr(), cnt.data_ptr(), lab_c.data_ptr())
        hc, hclo, hcbw = self.buf16((cap, Ht))
        self.emit(lib.vb_gather_rows16, ht.b16.data_ptr(), hc.data_ptr(), self._ptr(self._extra(ht.bw, ht.b16)), self._ptr(self._extra(hcbw, hc)),
                  idx.data_ptr(), cap, Ht)
        if hclo is not None:
            self.emit(lib.vb_gather_rows16, ht.lo.data_ptr(), hclo.data_ptr(), None, None, idx.data_ptr(), cap, Ht)
        wn = "bert.embeddings.word_embeddings.weight"
        logits = self.buf((cap, V), F32)
        self.gemm(cap, V, Ht, hc, Ht, ps.w16(wn), Ht, bias=ps.p("cls.predictions.bias"), out_f32=logits, ld_of=V, a_lo=hclo, b_lo=ps.w16lo(wn))
        ldp = _pad8(V)
        self.lm_c = dict(cap=cap, idx=idx, count=cnt, labels=lab_c, logits=logits, dl32=self.buf((cap, V), F32),
                         dl16=self.buf((cap, ldp), BF16, zero=True), ldp=ldp)

        def bwd():
            lc = self.lm_c
            self.colsum(lc["dl32"], V, ps.g("cls.predictions.bias"), cap, V)
            self.linear_wgrad(lc["dl16"], ldp, None, 0, hcbw, Ht, cap, V, Ht, None, gw=ps.g(wn))
            if ht.frozen:
                return
            gc = self.scratch("lm.gc", (cap, Ht), F32)
            self.gemm(cap, Ht, V, lc["dl16"], ldp, ps.w16b(wn), Ht, b_mn=1, out_f32=gc, ld_of=Ht)
            g = self.grad_of(ht)
            self.emit(lib.vb_memset_zero, g.data_ptr(), g.numel() * 4)
            self.emit(lib.vb_scatter_rows_f32, gc.data_ptr(), g.data_ptr(), idx.data_ptr(), cap, Ht, cnt.data_ptr(), self.loss.data_ptr())
            ht.gw = True
            ht_bwd()
        return bwd

    def lm_rows(self):
        """(labelled rows of the last step, capacity) of the compacted masked-LM head (device sync)."""
        return int(self.lm_c["count"].item()), self.lm_c["cap"]

    def transform(self, x, wdense, lnname, tag):
        """Linear -> GELU -> LayerNorm (BertPredictionHeadTransform / BertImgPredictionHeadTransform / the first three
        stages of SimpleClassifier; vilbert.py:1152-1156, 1172-1176, 1714-1718). x: Act or (b16, M, K) for a plain operand."""
        ps = self.ps
        M, K = x.M, x.H
        Hh = ps.p(wdense + ".weight").shape[0]
        g32, pre16 = self.buf((M, Hh), F32), self.buf((M, Hh), BF16)
        self.gemm(M, Hh, K, x.b16, K, ps.w16(wdense + ".weight"), K, bias=ps.p(wdense + ".bias"), act=L.VB_ACT_GELU, out_f32=g32, ld_of=Hh,
                  out_pre=pre16, ld_op=Hh, a_lo=x.lo, b_lo=ps.w16lo(wdense + ".weight"))
        _, h16, mean, rstd, hlo, hbw = self.ln_fwd(g32, ps.p(lnname + ".weight"), ps.p(lnname + ".bias"), M, Hh, want_f32=False)
        hn = Act(None, h16, M, Hh, lo=hlo, bw=hbw)

        def bwd():
            if not hn.gw:
                return
            dpre16 = self.scratch(tag + ".dpre16", (M, Hh), BF16)
            self.ln_bwd(hn.g32, g32, ps.p(lnname + ".weight"), mean, rstd, None, dpre16, M, Hh, ps.g(lnname + ".weight"), ps.g(lnname + ".bias"), pre=pre16,
                        gbias=ps.g(wdense + ".bias"))
            self.linear_wgrad(dpre16, Hh, None, 0, x.bw, K, M, Hh, K, wdense)
            self.dgrad_into(x, dpre16, Hh, ps.w16b(wdense + ".weight"), M, Hh, K)
        return hn, bwd

    def small_head(self, name, x, wname, N_out, addend=None, x32=None, M=None, K=None, in_drop=None):
        ps = self.ps
        M = x.M if M is None else M
        K = x.H if K is None else K
        xin = x.f32 if x32 is None else x32
        y = self.buf((M, N_out), F32)
        self.emit(self.lib.vb_small_linear_fwd, xin.data_ptr(), K, ps.p(wname + ".weight").data_ptr(), ps.p(wname + ".bias").data_ptr(),
                  self._ptr(addend), y.data_ptr(), M, K, N_out, self._ref(in_drop))
        self.outputs[name] = y

        def bwd():
            if name not in self.grad_outputs:
                return
            dy = self.out_grad_buffer(name, (M, N_out))
            g = self.grad_of(x)
            acc = 1 if x.gw else 0
            x.gw = True
            self.emit(self.lib.vb_small_linear_bwd, dy.data_ptr(), xin.data_ptr(), K, ps.p(wname + ".weight").data_ptr(), g.data_ptr(), K, acc,
                      ps.g(wname + ".weight").data_ptr(), ps.g(wname + ".bias").data_ptr(), M, K, N_out, self._ref(in_drop))
        self.push_bwd(bwd)

    def build_heads(self, seq_t, seq_v, pooled_t, pooled_v):
        """VILBertForVLTasks.forward after self.bert (vilbert.py:1673-1708) + BertPreTrainingHeads (:1228-1243).
        Dropout layers are identity (eval-mode / p = 0 parity protocol)."""
        ps, c, B, lib = self.ps, self.cfg, self.B, self.lib
        Hb, Ht, Hv, Nt, Nv = c.bi_hidden_size, c.hidden_size, c.v_hidden_size, self.Nt, self.Nv
        mul = 1 if c.fusion_method == "mul" else 0
        def fuse(drop):
            f32 = self.buf((B, Hb), F32)
            f16, flo, fbw = self.buf16((B, Hb))
            self.emit(lib.vb_fuse_pooled_fwd, pooled_t.f32.data_ptr(), pooled_v.f32.data_ptr(), f32.data_ptr(), f16.data_ptr(), B * Hb, mul, self._ref(drop),
                      self.op_fp16, self._ptr(flo), self._ptr(self._extra(fbw, f16)))
            act = Act(f32, f16, B, Hb, lo=flo, bw=fbw)

            def fuse_bwd():
                if not act.gw:
                    return
                for a in (pooled_t, pooled_v):
                    g = self.grad_of(a)
                    if not a.gw:
                        self.emit(lib.vb_memset_zero, g.data_ptr(), g.numel() * 4)
                        a.gw = True
                self.emit(lib.vb_fuse_pooled_bwd, act.g32.data_ptr(), pooled_t.f32.data_ptr(), pooled_v.f32.data_ptr(), pooled_t.g32.data_ptr(),
                          pooled_v.g32.data_ptr(), B * Hb, mul, self._ref(drop))
            self.push_bwd(fuse_bwd)
            return act
        # VILBertForVLTasks.dropout on the fused vector (vilbert.py:1677-1682); BertPreTrainingHeads has its own nn.Dropout(0.1)
        # on its own fused vector (:1233-1241) — a different mask, needed only where the alignment score is an output
        fused = fuse(self.drop("dropout.pooled", self.head_dropout_prob)) if self.heads == "vl" else None
        need_cls_fused = self.heads == "pretraining" or (B % 2 == 1)
        cls_drop = self.drop("cls.dropout", 0.1)
        if need_cls_fused:
            fused_cls = fuse(cls_drop) if (cls_drop is not None or fused is None) else fused
        else:
            fused_cls = None
        f32, f16, f16lo, f16bw = (fused.f32, fused.b16, fused.lo, fused.bw) if fused is not None else (None, None, None, None)

        # --- cls: masked-LM head (decoder tied to the word embeddings), image-region head, alignment head
        ht, ht_bwd = self.transform(seq_t, "cls.predictions.transform.dense", "cls.predictions.transform.LayerNorm", "lm.tr")
        # fused pre-training objective: only the masked rows enter the LM cross-entropy, so the tied decoder runs on those alone
        self.lm_c = None
        if self.loss_kind == "pretraining" and self.e.lm_compact:
            lm_bwd = None
            lm_compact_bwd = self.lm_head_compact(ht, ht_bwd)
        else:
            lm_bwd = self.big_head("linguisic_prediction", ht.b16, Ht, ht, B * Nt, Ht, c.vocab_size, None, "cls.predictions.bias",
                                   w16=ps.w16("bert.embeddings.word_embeddings.weight"), gw=ps.g("bert.embeddings.word_embeddings.weight"),
                                   w16lo=ps.w16lo("bert.embeddings.word_embeddings.weight"), w16b=ps.w16b("bert.embeddings.word_embeddings.weight"))
        hv, hv_bwd = self.transform(seq_v, "cls.imagePredictions.transform.dense", "cls.imagePredictions.transform.LayerNorm", "im.tr")
        im_bwd = self.big_head("vision_prediction", hv.b16, Hv, hv, B * Nv, Hv, c.v_target_size, "cls.imagePredictions.decoder",
                               "cls.imagePredictions.decoder.bias")

        def wide_bwd(head_bwd, hn, tr_bwd, K, N_out):
            def f():
                r = head_bwd()
                if r is None:
                    return
                dl16, ldp, W16 = r
                g = self.grad_of(hn)
                self.gemm(hn.M, K, N_out, dl16, ldp, W16, K, b_mn=1, out_f32=g, ld_of=K)
                hn.gw = True
                tr_bwd()
            return f
        self.push_bwd(wide_bwd(lm_bwd, ht, ht_bwd, Ht, c.vocab_size) if lm_bwd is not None else lm_compact_bwd)
        self.push_bwd(wide_bwd(im_bwd, hv, hv_bwd, Hv, c.v_target_size))

        if self.heads == "pretraining":
            # BertForMultiModalPreTraining returns the alignment score of self.cls (vilbert.py:1497)
            self.small_head("seq_relationship_score", fused_cls, "cls.bi_seq_relationship", 2)
            return
        if B % 2 == 0:
            # vil_binary_prediction pairs consecutive samples: pooled.view(-1, 2*Hb) (:1686-1689)
            pair = Act(f32.view(B // 2, 2 * Hb), f16.view(B // 2, 2 * Hb), B // 2, 2 * Hb, lo=f16lo.view(B // 2, 2 * Hb) if f16lo is not None else None, bw=f16bw.view(B // 2, 2 * Hb))
            hb, hb_bwd = self.transform(pair, "vil_binary_prediction.logit_fc.0", "vil_binary_prediction.logit_fc.2", "bin.tr")
            # LayerNorm output is needed in fp32 for the 2-way linear: recompute it from the bf16 copy is lossy, so run the small
            # linear on an fp32 LayerNorm output
            hb32 = self.buf((B // 2, 2 * Hb), F32)
            # re-emit LN with an fp32 output (cheap: B/2 rows)
            fwd_fn, fwd_args, fwd_sid = self.fwd[-1]
            args = list(fwd_args); args[5] = hb32.data_ptr(); self.fwd[-1] = (fwd_fn, tuple(args), fwd_sid)
            hb.f32 = hb32

            def bin_bwd():
                if not hb.gw:
                    return
                hb_bwd()
                if pair.gw:   # gradient landed in pair.g32 [B/2, 2Hb] == [B, Hb]
                    self.add_grad(fused, pair.g32.view(B, Hb))
            self.push_bwd(bin_bwd)   # registered first => runs after the 2-way linear's backward
            self.small_head("vil_binary_prediction", hb, "vil_binary_prediction.logit_fc.3", 2)
        else:
            # odd batch: the reference returns the [B, 2] alignment output of self.cls here (:1673, 1686)
            self.small_head("vil_binary_prediction", fused_cls, "cls.bi_seq_relationship", 2)

        for nm, n_out in (("vil_prediction", 3129), ("vil_prediction_gqa", 1533)):
            hh, hh_bwd = self.transform(fused, nm + ".logit_fc.0", nm + ".logit_fc.2", nm + ".tr")
            head_bwd = self.big_head(nm, hh.b16, 2 * Hb, hh, B, 2 * Hb, n_out, nm + ".logit_fc.3", nm + ".logit_fc.3.bias")
            self.push_bwd(wide_bwd(head_bwd, hh, hh_bwd, 2 * Hb, n_out))
        self.small_head("vil_logit", fused, "vil_logit", 1)
        self.small_head("vil_tri_prediction", fused, "vil_tri_prediction", 3)
        self.small_head("vision_logit", seq_v, "vision_logit", 1, addend=self.mask_v, in_drop=self.drop("dropout.seq_v", self.head_dropout_prob))
        self.small_head("linguisic_logit", seq_t, "linguisic_logit", 1, in_drop=self.drop("dropout.seq_t", self.head_dropout_prob))

    # ------------------------------------------------------------------ whole model
    def _build(self):
        c, B = self.cfg, self.B
        self.outputs, self.gout = OrderedDict(), {}
        self.loss_inputs = {}
        self.loss = self.buf((1,), F32, zero=True) if (self.loss_kind is not None and not self.vqa_loss) else None
        self.enc_t, self.enc_v = [], []
        t, v = self.embeddings()
        # BertEncoder.forward interleaving schedule (vilbert.py:960-1096)
        t_start = v_start = 0
        for count, (v_end, t_end) in enumerate(zip(c.v_biattention_id, c.t_biattention_id)):
            # fixed_t_layer / fixed_v_layer (vilbert.py:968-1003): the first layers of a stream run under no_grad — their backward is
            # not emitted and their output stops the gradient (embeddings and the frozen layers' parameters receive none)
            for i in range(t_start, t_end):
                frozen = i < getattr(c, "fixed_t_layer", 0)
                self._no_grad = frozen
                t = self.text_layer(t, i)
                self._no_grad = False
                t.frozen = frozen
            pool = None
            if self.dyn and v_end > v_start:
                # dynamic_attention: this segment's image layers read the pooled text states of the segment's END (the text layers
                # run first in the reference, vilbert.py:977-1004), so the two streams cannot overlap here
                pool = self.text_pool(t)
                self.sync_streams()
            with self.on(1):
                for i in range(v_start, v_end):
                    frozen = i < getattr(c, "fixed_v_layer", 0)
                    self._no_grad = frozen
                    v = self.image_layer(v, i, pool)
                    self._no_grad = False
                    v.frozen = frozen
            if count == 0 and self.fast:
                t = self.broadcast_text(t)
            if count == 0 and self.pairs:
                t, v = self.expand_pairs(t, v)
            if c.with_coattention:
                v, t = self.connection_layer(v, t, count)
            v_start, t_start = v_end, t_end
            self.enc_t.append(t); self.enc_v.append(v)     # output_all_encoded_layers: one entry per connection layer (:1075-1077)
        pool = None
        if self.dyn and c.v_num_hidden_layers > v_start:
            pool = self.text_pool(t)      # the trailing image layers see the text states BEFORE the trailing text layers (:1079-1092)
            self.sync_streams()
        with self.on(1):
            for i in range(v_start, c.v_num_hidden_layers):
                v = self.image_layer(v, i, pool)
        for i in range(t_start, c.num_hidden_layers):
            t = self.text_layer(t, i)
        self.sync_streams()      # poolers and heads read both streams; they run on the main stream
        self.seq_t, self.seq_v = t, v
        self.pooled_t = self.pooler(t, self.Nt, "bert.t_pooler.dense")
        self.pooled_v = self.pooler(v, self.Nv, "bert.v_pooler.dense")
        self.outputs["sequence_output_t"] = t.f32.view(B, self.Nt, -1)
        self.outputs["sequence_output_v"] = v.f32.view(B, self.Nv, -1)
        self.outputs["pooled_output_t"] = self.pooled_t.f32
        self.outputs["pooled_output_v"] = self.pooled_v.f32
        if self.heads != "none":
            self.build_heads(t, v, self.pooled_t, self.pooled_v)
            for nm in ("vision_prediction", "vision_logit"):
                if nm in self.outputs:
                    self.outputs[nm] = self.outputs[nm].view(B, self.Nv, -1)
            for nm in ("linguisic_prediction", "linguisic_logit"):
                if nm in self.outputs:
                    self.outputs[nm] = self.outputs[nm].view(B, self.Nt, -1)
        self.sync_streams(mirror=False)
        self.n_kernels_fwd = sum(1 for op in self.fwd if op[0] is not None)

        # ---------------- backward
        self.cur = self.bwd
        if self.vqa_loss:
            lg = self.outputs["vil_prediction"]
            self.vqa_target = self.buf(tuple(lg.shape), F32, zero=True)
            self.loss = self.buf((1,), F32, zero=True)
            self.vqa_dl32 = self.buf(tuple(lg.shape), F32)
            self.vqa_dl16 = self.buf((lg.shape[0], _pad8(lg.shape[1])), BF16, zero=True)
            self.emit(self.lib.vb_bce_logits_loss, lg.data_ptr(), self.vqa_target.data_ptr(), self.loss.data_ptr(), self.vqa_dl32.data_ptr(),
                      self.vqa_dl16.data_ptr(), _pad8(lg.shape[1]), lg.shape[0], lg.shape[1], 1.0)
        elif self.loss_kind is not None:
            self._emit_loss()
        # gradients flowing into the BertModel outputs themselves
        for nm, act in (("sequence_output_t", self.seq_t), ("sequence_output_v", self.seq_v), ("pooled_output_t", self.pooled_t),
                        ("pooled_output_v", self.pooled_v)):
            if nm in self.grad_outputs:
                self.add_grad(act, self.out_grad_buffer(nm, (act.M, act.H)))
        self.sync_streams()
        for entry in reversed(self._bwd_emitters):
            if entry is None:
                self.sync_streams()
            else:
                self.sid, emitter = entry
                self._scratch_epoch += 1
                emitter()
        self.sid = 0
        self.cur.append((None, ("all",), 0))     # join every stream (incl. the weight-gradient side streams)
        self.n_kernels_bwd = sum(1 for op in self.bwd if op[0] is not None)
        self.cur = self.fwd

    def attention_export(self):
        """The reference's all_attention_mask triple (BertEncoder.forward, vilbert.py:1098-1107) for config.visualization: lists of
        attn_data dicts in layer order — text / image: {"attn", "queries", "keys"}; connection layers: {"attn1", "queries1",
        "keys1", "attn2", "querues2" (the reference's spelling), "keys2"}. Tensors are fp32 [B, heads, N, ...] copies."""
        def qk(d, which, N):
            x = d[which]
            return x.float().reshape(d["B"], N, d["H"], d["D"]).permute(0, 2, 1, 3).contiguous()
        def one(d):
            return {"attn": d["attn"].clone(), "queries": qk(d, "q", d["Nq"]), "keys": qk(d, "k", d["Nk"])}
        ts, vs = [one(d) for d in self.attn_t], [one(d) for d in self.attn_v]
        cs = [{"attn1": a1["attn"].clone(), "queries1": qk(a1, "q", a1["Nq"]), "keys1": qk(a1, "k", a1["Nk"]),
               "attn2": a2["attn"].clone(), "querues2": qk(a2, "q", a2["Nq"]), "keys2": qk(a2, "k", a2["Nk"])} for a1, a2 in self.attn_c]
        return ts, vs, cs

    def _emit_loss(self):
        """Fused objectives other than "vqa": one loss kernel per head writes the scalar (self.loss, fp32 device) and the fp32
        d(loss)/d(head output) into the plan's output-gradient buffer, from where the head's backward proceeds as for a
        caller-supplied gradient. Labels / targets are static plan inputs (self.loss_inputs)."""
        lib, B, k = self.lib, self.B, self.loss_kind
        missing = [n for n in LOSS_HEADS[k] if n not in self.grad_outputs]
        if missing:
            raise ValueError(f"loss={k!r} differentiates {LOSS_HEADS[k]}: add them to grad_outputs")
        li = self.loss_inputs

        def ce(name, rows, cols, label_key, acc):
            lg = self.outputs[name]
            li[label_key] = self.buf((rows,), I64, zero=True)
            d = self.out_grad_buffer(name, tuple(lg.shape))
            self.emit(lib.vb_ce_loss, lg.data_ptr(), cols, li[label_key].data_ptr(), -1, self.loss.data_ptr(), d.data_ptr(), cols, None, 0,
                      rows, cols, 1.0, 1 if acc else 0)

        def bce(name, rows, cols):
            lg = self.outputs[name]
            li["target"] = self.buf((rows, cols), F32, zero=True)
            d = self.out_grad_buffer(name, tuple(lg.shape))
            self.emit(lib.vb_bce_logits_loss, lg.data_ptr(), li["target"].data_ptr(), self.loss.data_ptr(), d.data_ptr(), None, 0, rows, cols, 1.0)

        if k == "gqa":
            bce("vil_prediction_gqa", B, 1533)
        elif k == "vlogit_bce":
            bce("vision_logit", B, self.Nv)
        elif k == "logit_ce":
            opts = self.e.loss_options
            if B % opts:
                raise ValueError(f"loss='logit_ce': batch {B} is not a multiple of {opts} options")
            ce("vil_logit", B // opts, opts, "labels", False)
        elif k == "binary_ce":
            ce("vil_binary_prediction", self.outputs["vil_binary_prediction"].shape[0], 2, "labels", False)
        elif k == "tri_ce":
            ce("vil_tri_prediction", B, 3, "labels", False)
        elif k == "pretraining":
            # vilbert.py:1578-1590 (+ train_concap.py: loss = masked_loss_t + masked_loss_v + next_sentence_loss)
            V, C = self.cfg.vocab_size, self.cfg.v_target_size
            if self.lm_c is not None:
                lc = self.lm_c
                self.emit(lib.vb_ce_loss, lc["logits"].data_ptr(), V, lc["labels"].data_ptr(), -1, self.loss.data_ptr(), lc["dl32"].data_ptr(), V,
                          lc["dl16"].data_ptr(), lc["ldp"], lc["cap"], V, 1.0, 0)
            else:
                ce("linguisic_prediction", B * self.Nt, V, "masked_lm_labels", False)
            sv = self.outputs["vision_prediction"]
            li["image_target"] = self.buf((B, self.Nv - 1, C), F32, zero=True)
            li["image_label"] = self.buf((B, self.Nv - 1), I64, zero=True)
            dv = self.out_grad_buffer("vision_prediction", tuple(sv.shape))
            self.emit(lib.vb_kl_masked_loss, sv.data_ptr(), li["image_target"].data_ptr(), li["image_label"].data_ptr(), self.loss.data_ptr(),
                      dv.data_ptr(), None, 0, B, self.Nv, C, 1.0, 1)
            ce("seq_relationship_score", B, 2, "next_sentence_label", True)

    # ------------------------------------------------------------------ execution
    def load_inputs(self, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None,
                    task_ids=None, non_blocking=True):
        """Host (ideally pinned) or device tensors -> the plan's static input buffers."""
        self.in_ids.copy_(input_txt, non_blocking=non_blocking)
        if token_type_ids is None:
            self.in_tt.zero_()
        else:
            self.in_tt.copy_(token_type_ids, non_blocking=non_blocking)
        if attention_mask is None:
            self.in_amask.fill_(1)
        else:
            self.in_amask.copy_(attention_mask, non_blocking=non_blocking)
        if self.fast:
            self.in_amask_b.copy_(self.in_amask.expand_as(self.in_amask_b))
        if image_attention_mask is None:
            self.in_imask.fill_(1)
        else:
            self.in_imask.copy_(image_attention_mask, non_blocking=non_blocking)
        if self.pairs:
            b = self.Bin
            self.in_amask_pairs.view(b, b, -1).copy_(self.in_amask.unsqueeze(1).expand(b, b, -1))
            self.in_imask_pairs.view(b, b, -1).copy_(self.in_imask.unsqueeze(0).expand(b, b, -1))
        self.in_feat.copy_(input_imgs, non_blocking=non_blocking)
        self.in_loc.copy_(image_loc, non_blocking=non_blocking)
        if self.has_task:
            if task_ids is None:
                raise ValueError("task_specific_tokens is set: task_ids is required")
            self.in_task.copy_(task_ids.reshape(-1), non_blocking=non_blocking)

    def _run(self, ops):
        """Issues the ops on their streams. Markers: (None, ()) = barrier between the text and vision streams;
        (None, ("all",)) = join of every stream; (None, ("rec"|"wait", id)) = event edge to a weight-gradient side stream."""
        main = torch.cuda.current_stream()
        if self._streams is None:
            n = 4 if self.wgrad_streams else (2 if self.two_streams else 1)
            self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(n - 1)]
        streams = [main] + self._streams
        handles = [st.cuda_stream for st in streams]
        check = L.check
        events = {}
        for fn, args, sid in ops:
            if fn is None:
                if not args:                      # text <-> vision barrier
                    if len(streams) > 1:
                        aux = streams[1]
                        e1 = torch.cuda.Event(); e1.record(main); aux.wait_event(e1)
                        e2 = torch.cuda.Event(); e2.record(aux); main.wait_event(e2)
                elif args[0] == "all":
                    # full barrier over every stream. Fork first (main -> side streams), then join (side -> main): inside a
                    # graph capture a side stream only belongs to the capture once it has waited on a captured event
                    e = torch.cuda.Event(); e.record(main)
                    for st in streams[1:]:
                        st.wait_event(e)
                    for st in streams[1:]:
                        e2 = torch.cuda.Event(); e2.record(st); main.wait_event(e2)
                elif args[0] == "rec":
                    e = torch.cuda.Event(); e.record(streams[sid]); events[args[1]] = e
                elif args[0] == "wait":
                    streams[sid].wait_event(events[args[1]])
                continue
            st = fn(*args, handles[sid] if sid < len(handles) else handles[0])
            if st:
                check(st, fn.__name__)

    def run_forward(self):
        self.fwd_id += 1
        self.e.arena_owner = (self, self.fwd_id)
        if self.graph_fwd is not None:
            self.graph_fwd.replay()
        else:
            self._run(self.fwd)
            self._eager_runs[0] += 1

    def run_backward(self):
        if self.e.arena is not None and self.e.arena_owner != (self, self.fwd_id):
            raise L.VBError("shared activation arena: another plan's forward ran between this plan's forward and backward "
                            "(its saved activations are gone); run forward + backward per batch, or disable the arena")
        self.e.grad_clean = False
        if self.graph_bwd is not None:
            self.graph_bwd.replay()
        else:
            self._run(self.bwd)
            self._eager_runs[1] += 1

    def maybe_capture_passes(self, after=2):
        """Module-surface path: once a pass of this plan has run eagerly `after` times (kernels loaded, attributes set), capture
        it into its own CUDA graph — without a warm-up run, which would accumulate into the gradient buffer — so that the
        ~600 ctypes launches of a step become two graph replays."""
        if self.graph_fwd is None and self._eager_runs[0] >= after:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(self.fwd)
            self.graph_fwd = g
        if self.graph_bwd is None and self._eager_runs[1] >= after:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(self.bwd)
            self.graph_bwd = g

    def enable_training_prologue(self, zero_grad=True, refresh_weights=True):
        """Makes run_step a complete training-step body: bump the dropout step counter (train mode), zero the flat
        gradient buffer and refresh the bf16 weight shadow from the fp32 master parameters (what an optimizer step
        invalidates) before the forward. With two streams only the weights the first text layers need (everything laid out
        before the first connection layer) are cast on the main stream; the rest of the cast and the gradient memset run on
        the vision stream underneath those text layers (the vision stream's own first consumer, the image embedding, is
        queued behind them and is not needed before the first connection layer)."""
        ps, lib = self.ps, self.lib
        self.prologue = []
        if self.train:
            self.prologue.append((lib.vb_step_counter_bump, (self.e.drop_step.data_ptr(),), 0))
        n = ps.numel
        first_c = [off for name, (off, _) in ps.entries.items() if ".c_layer." in name]
        split = min(first_c) if (self.two_streams and first_c) else n
        if refresh_weights and split > 0:
            sb = ps.shadow_b if ps.shadow_b is not ps.shadow else None
            self.prologue.append((lib.vb_cast_f32_to_bf16, (ps.flat.data_ptr(), ps.shadow.data_ptr(), split, self.op_fp16,
                                                            ps.shadow_lo.data_ptr() if self.split else None, sb.data_ptr() if sb is not None else None), 0))
        tail = []
        if zero_grad:
            tail.append((lib.vb_memset_zero, (ps.grad.data_ptr(), ps.grad.numel() * 4), 1 if self.two_streams else 0))
        if refresh_weights and split < n:
            tail.append((lib.vb_cast_f32_to_bf16, (ps.flat.data_ptr() + 4 * split, ps.shadow.data_ptr() + 2 * split, n - split, self.op_fp16,
                                                   ps.shadow_lo.data_ptr() + 2 * split if self.split else None,
                                                   ps.shadow_b.data_ptr() + 2 * split if ps.shadow_b is not ps.shadow else None), 1))
        if self.two_streams:
            self.prologue += [(None, (), 0)] + tail       # barrier: the vision stream starts after the main-stream part
        else:
            self.prologue += tail
        self.graph_step = None

    def enable_optimizer(self, opt, dropout_bump=True):
        """Training step with the fused optimizer (optim.FusedAdamW): the step body becomes [dropout step bump] + forward + loss +
        backward + ONE AdamW launch that also rewrites the 16-bit weight copy and zeroes the gradients — no weight cast and no
        gradient memset in the step. (The host-side lr table of `opt` is refreshed by opt.step(); here the launch alone is
        replayed, e.g. inside the step graph, with the table currently on the device.)"""
        self.prologue = []
        if self.train and dropout_bump:
            self.prologue.append((self.lib.vb_step_counter_bump, (self.e.drop_step.data_ptr(),), 0))
        fn, args = opt.op()
        self.epilogue = [(fn, args, 0)]
        self.graph_step = None

    @property
    def n_launches_step(self):
        return sum(1 for op in self.prologue + self.epilogue if op[0] is not None) + self.n_kernels_fwd + self.n_kernels_bwd

    def run_step(self):
        """(prologue) + forward + (loss) + backward (+ epilogue); gradients accumulate into ParamStore.grad."""
        self.fwd_id += 1
        self.e.arena_owner = (self, self.fwd_id)
        self.e.grad_clean = False
        if self.graph_step is not None:
            self.graph_step.replay()
        else:
            self._run(self.prologue)
            self._run(self.fwd)
            self._run(self.bwd)
            self._run(self.epilogue)

    def ddp_segments(self, n_segments=4, tail_cut=True):
        """Cuts the backward op list at stream barriers into `n_segments` pieces and returns
        [(bwd_op_lo, bwd_op_hi, grad_lo, grad_hi)]: after piece i has run, the flat gradient range [grad_lo, grad_hi) is
        final (no later op writes it) and may be all-reduced while the remaining pieces execute. The ranges tile the
        whole flat buffer from its end (heads, last layers) to its start (embeddings). With `tail_cut` one extra piece holds
        only the last few kernels, so up to n_segments + 1 pieces are returned."""
        n_ops = len(self.bwd)
        # legal cut positions: a piece ends with a join of every stream and the next one starts with a fork, so any position
        # works as long as no event recorded in one piece is waited for in a later one (rec/wait markers of the side streams)
        rec_pos, last_wait = {}, {}
        for i, op in enumerate(self.bwd):
            if op[0] is None and len(op[1]) == 2:
                if op[1][0] == "rec": rec_pos[op[1][1]] = i
                elif op[1][0] == "wait": last_wait[op[1][1]] = i
        straddle = [0] * (n_ops + 1)
        for ev, r in rec_pos.items():
            for i in range(r + 1, last_wait.get(ev, r) + 1):
                straddle[i] += 1
        n_kern = [0] * (n_ops + 1)                      # kernels in bwd[:i]
        for i, op in enumerate(self.bwd):
            n_kern[i + 1] = n_kern[i] + (op[0] is not None)
        cand = [i for i in range(1, n_ops) if straddle[i] == 0 and self.bwd[i - 1][0] is not None or
                (straddle[i] == 0 and self.bwd[i - 1][0] is None and not self.bwd[i - 1][1])]
        cuts = []
        for k in range(1, n_segments):
            want = n_kern[n_ops] * k // n_segments     # equal kernel counts per piece
            free = [c for c in cand if c not in cuts]
            if not free:
                break
            cuts.append(min(free, key=lambda c: abs(n_kern[c] - want)))
        if tail_cut and cand:
            # one more cut just before the last kernels (the embedding backward): everything except the embedding tables, whose
            # gradients only the very last kernels produce, leaves the exposed final all-reduce
            late = [c for c in cand if n_kern[n_ops] - n_kern[c] >= 3]
            if late and (not cuts or late[-1] > max(cuts)):
                cuts.append(late[-1])
        cuts = sorted(set(cuts)) + [n_ops]
        # every piece must contain at least one kernel (an empty CUDA graph is legal but pointless)
        kept, prev = [], 0
        for cpos in cuts:
            if n_kern[cpos] > n_kern[prev]:
                kept.append(cpos)
                prev = cpos
        if not kept or kept[-1] != n_ops:               # trailing markers join the last piece that has kernels
            if kept: kept[-1] = n_ops
            else: kept = [n_ops]
        cuts = kept
        ranges = sorted(self.grad_touch.items())          # by flat offset
        numel = self.e.ps.numel
        segs, lo_op, hi_grad = [], 0, numel
        for cut in cuts:
            ready_lo = 0 if cut == n_ops else hi_grad
            if cut != n_ops:
                for (off, n), touch in reversed(ranges):
                    if off >= hi_grad:
                        continue
                    if touch >= cut:
                        break
                    ready_lo = off
            segs.append((lo_op, cut, ready_lo, hi_grad))
            lo_op, hi_grad = cut, ready_lo
        return segs

    def capture_segments(self, n_segments=4, tail_cut=True):
        """Captures the step as CUDA graphs, one per backward piece of ddp_segments (graph 0 = prologue + forward + first
        backward piece), for the data-parallel step: see run_step_overlapped."""
        self.segments = self.ddp_segments(n_segments, tail_cut=tail_cut)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._run(self.prologue); self._run(self.fwd); self._run(self.bwd)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        barrier = [(None, ("all",), 0)]
        self.segment_graphs = []
        for i, (lo, hi, _, _) in enumerate(self.segments):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if i == 0:
                    self._run(self.prologue); self._run(self.fwd)
                self._run(barrier + self.bwd[lo:hi] + barrier)
            self.segment_graphs.append(g)
        torch.cuda.synchronize()

    def live_ranges(self, lo, hi):
        """Sub-ranges of the flat gradient range [lo, hi) that the backward of THIS plan writes (coalesced, 1024-element
        granularity): parameters whose gradient is identically zero for this plan (heads outside the objective, the never-called
        q_dense1/2 of BertBiOutput, vilbert.py:834,841) are not exchanged — zeros average to zeros on every rank."""
        spans = sorted((off, off + n) for (off, n) in self.grad_touch if off + n > lo and off < hi)
        out = []
        for a, b in spans:
            a, b = max(a, lo) // 1024 * 1024, min(-(-min(b, hi) // 1024) * 1024, hi)
            a = max(a, lo)
            if out and a <= out[-1][1] + 4096:
                out[-1][1] = max(out[-1][1], b)
            else:
                out.append([a, b])
        return [(a, b) for a, b in out if b > a]

    def capture_step_ddp(self, allreduce_range, n_segments=8, tail_cut=True, skip_dead=True):
        """Data-parallel step as ONE CUDA graph: prologue + forward + the backward pieces of ddp_segments, with the all-reduce of
        each finished gradient range captured on a communication stream inside the same graph (NCCL collectives are capturable),
        forked after its piece and joined at the end. Compared with capture_segments / run_step_overlapped there is a single
        graph launch per step and no host-side event bookkeeping between pieces. `allreduce_range(lo, hi)` must enqueue the
        collective on the current stream (async_op=False semantics)."""
        self.segments = self.ddp_segments(n_segments, tail_cut=tail_cut)
        torch.cuda.synchronize()
        barrier = [(None, ("all",), 0)]
        g = torch.cuda.CUDAGraph()
        comm = torch.cuda.Stream(device=self.dev)
        self._ddp_comm = comm
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            self._run(self.prologue); self._run(self.fwd)
            for (lo_op, hi_op, lo, hi) in self.segments:
                self._run(barrier + self.bwd[lo_op:hi_op] + barrier)
                if hi > lo:
                    ev = torch.cuda.Event(); ev.record(main)
                    comm.wait_event(ev)
                    with torch.cuda.stream(comm):
                        for (a, b) in (self.live_ranges(lo, hi) if skip_dead else [(lo, hi)]):
                            allreduce_range(a, b)
            ev = torch.cuda.Event(); ev.record(comm)
            main.wait_event(ev)
            self._run(self.epilogue)
        self.graph_step_ddp = g
        torch.cuda.synchronize()

    def run_step_ddp(self):
        self.fwd_id += 1
        self.e.arena_owner = (self, self.fwd_id)
        self.e.grad_clean = False
        self.graph_step_ddp.replay()

    def run_step_overlapped(self, allreduce_range, comm_stream, skip_dead=True):
        """Replays the segment graphs; after each one the finished tail range of the flat gradient buffer is handed to
        `allreduce_range(lo, hi)` (issued under `comm_stream`, which first waits for that segment) so the collective
        overlaps the rest of the backward. Ranges no backward op of this plan writes (live_ranges) are not exchanged.
        Returns the list of whatever allreduce_range returned (async work handles)."""
        self.fwd_id += 1
        self.e.grad_clean = False
        main = torch.cuda.current_stream()
        works = []
        for g, (_, _, lo, hi) in zip(self.segment_graphs, self.segments):
            g.replay()
            if hi > lo:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(ev)
                    for (a, b) in (self._live_cache(lo, hi) if skip_dead else [(lo, hi)]):
                        works.append(allreduce_range(a, b))
        return works

    def _live_cache(self, lo, hi):
        c = self.__dict__.setdefault("_live_ranges_cache", {})
        if (lo, hi) not in c:
            c[(lo, hi)] = self.live_ranges(lo, hi)
        return c[(lo, hi)]

    def capture(self, separate=False):
        """Captures the plan into CUDA graphs (one for the whole step, or one per pass)."""
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):   # warm-up outside capture (module load, smem attribute calls)
            self._run(self.prologue)
            self._run(self.fwd)
            self._run(self.bwd)
            self._run(self.epilogue)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if separate:
            self.graph_fwd, self.graph_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fwd):
                self._run(self.fwd)
            with torch.cuda.graph(self.graph_bwd):
                self._run(self.bwd)
        else:
            self.graph_step = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_step):
                self._run(self.prologue)
                self._run(self.fwd)
                self._run(self.bwd)
                self._run(self.epilogue)
        torch.cuda.synchronize()


class Engine:
    """Owns the parameters and the per-shape plans."""

    def __init__(self, cfg, device="cuda", heads="vl", _build_only=False, two_streams=True, wgrad_streams=True, precision="fp16"):
        """_build_only=True (tests) allows a CPU device: plans can be constructed and inspected but never run.
        precision: "fp16" | "fp32" | "bf16" (module docstring)."""
        cfg.check_supported()
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        self.op_dtype = BF16 if precision == "bf16" else F16
        self.split = precision == "fp32"
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda" and not _build_only:
            raise L.VBError("vilbert_b200 runs on sm_100a GPUs only; there is no CPU path (device=%s)" % device)
        L.lib()  # fail loudly now if the extension is missing
        self.ps = ParamStore(cfg, self.device, heads, self.op_dtype, self.split)
        self.two_streams = two_streams   # text / vision segments on two CUDA streams (parallel graph branches)
        self.wgrad_streams = wgrad_streams   # weight-gradient GEMMs on two more side streams (off the backward critical chain)
        self.plans = OrderedDict()       # LRU cache of per-shape plans (each owns its activation buffers)
        self.max_plans = 16              # a 12-in-1 mix has ~12 shapes x {train} x one gradient set in steady state
        self.head_dropout_prob = 0.1     # VILBertForVLTasks(dropout_prob=0.1), vilbert.py:1601
        self.drop_step = torch.zeros(1, dtype=torch.int32, device=self.device)   # dropout step counter (uint32 on the device)
        self.drop_step_host = 0          # host mirror of the counter for the eager module path (bump / set below)
        self.shadow_clean = False        # the 16-bit weight copy matches the fp32 master parameters
        self.shadow_trusted = False      # True while the engine's own fused optimizer is the only writer of the parameters
        self.grad_clean = False          # the flat gradient buffer is all zeros (set by zero_grad / the fused optimizer)
        self.loss_options = 4            # answer options per question of the VL-logit objective (retrieval / VCR: 4)
        self.auto_graph = True           # module surface: capture a plan's passes into CUDA graphs after two eager runs
        self.arena = None                # optional shared activation arena (enable_activation_arena)
        self.arena_owner = None          # (plan, forward id) whose activations the arena currently holds
        self.bwd_gemm_max_ctas = 0       # persistent CTAs of the backward GEMMs (0 = one per SM); data parallel: leave SMs to NCCL
        self.lm_compact = True           # fused pre-training objective: masked-LM decoder + CE on the labelled rows only (Plan.lm_head_compact)
        self.lm_capacity = 0.25          # ... with room for this fraction of the token rows (15 % are masked; more poisons the loss with NaN)

    def plan(self, B, Nt, Nv, grad_outputs=(), vqa_loss=False, heads=None, train=False, loss=None):
        loss = "vqa" if vqa_loss else loss
        key = (B, Nt, Nv, frozenset(grad_outputs), loss, heads, bool(train), (self.lm_compact, self.lm_capacity) if loss == "pretraining" else None)
        if key in self.plans:
            self.plans.move_to_end(key)
            return self.plans[key]
        while len(self.plans) >= self.max_plans:   # evict the least recently used plan: its buffers go back to the allocator
            self.plans.popitem(last=False)
        self.plans[key] = Plan(self, B, Nt, Nv, grad_outputs, False, heads, train, loss=loss)
        return self.plans[key]

    def enable_activation_arena(self, nbytes):
        """One activation arena shared by all plans built afterwards (12-in-1 training holds a plan per task shape, but runs one
        forward + backward at a time, vilbert/task_utils.py:313-374 + train_tasks.py:545-551): their activation / scratch buffers
        overlay each other in `nbytes` of device memory instead of adding up. A plan's outputs must be consumed (the module
        surface copies the ones it returns) before another plan runs; run_backward refuses to run on clobbered activations."""
        if self.plans:
            raise L.VBError("enable_activation_arena must be called before the first plan is built")
        self.arena = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)

    def release_plans(self):
        """Drops every cached plan (and its activation / scratch buffers)."""
        self.plans.clear()

    def bump_dropout_step(self):
        """New dropout masks for the next forward (plans with a training prologue do this inside their graph)."""
        L.check(L.lib().vb_step_counter_bump(self.drop_step.data_ptr(), torch.cuda.current_stream().cuda_stream), "vb_step_counter_bump")
        self.drop_step_host = (self.drop_step_host + 1) & 0xFFFFFFFF

    def set_dropout_step(self, step):
        self.drop_step_host = int(step) & 0xFFFFFFFF
        self.drop_step.fill_(self.drop_step_host if self.drop_step_host < 2 ** 31 else self.drop_step_host - 2 ** 32)

    def refresh_weights(self):
        self.ps.refresh_shadow(torch.cuda.current_stream().cuda_stream)
        self.shadow_clean = True

    def zero_grad(self, force=False):
        """Zeroes the flat gradient buffer unless it is known to be clean (the fused optimizer zeroes it in its own pass)."""
        if self.grad_clean and not force:
            return
        L.check(L.lib().vb_memset_zero(self.ps.grad.data_ptr(), self.ps.grad.numel() * 4, torch.cuda.current_stream().cuda_stream))
        self.grad_clean = True
