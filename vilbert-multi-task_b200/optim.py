"""Fused multi-tensor AdamW on the engine's flat parameter / gradient buffers.

Drop-in for the optimizer the reference builds at train_tasks.py:401-426:

    optimizer = AdamW(optimizer_grouped_parameters, lr=base_lr, correct_bias=False)        # pytorch_transformers 1.0.0
 -> optimizer = FusedAdamW(optimizer_grouped_parameters, lr=base_lr, correct_bias=False, model=model)

Same constructor arguments and defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-6, weight_decay 0, correct_bias True), same
`param_groups` list of dicts (one group per tensor with its own lr / weight_decay in the reference; the warm-up schedulers
of train_tasks.py:431-437 mutate group["lr"] exactly as before). step() is ONE kernel launch over the flat buffers
(csrc/vb_optim.cu) that also writes the 16-bit tensor-core operand copy of the updated weights and zeroes the gradients,
so the training step needs no separate weight cast and the `model.zero_grad()` that follows optimizer.step() in the
reference (train_tasks.py:551) finds the buffer already clean.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_GROUP_DT = np.dtype([("lr", "<f4"), ("beta1", "<f4"), ("beta2", "<f4"), ("eps", "<f4"), ("weight_decay", "<f4"), ("correct_bias", "<i4")])


def build_chunks(ranges, chunk=32768):
    """ranges: [(flat offset, numel, group index)] -> (start int64[], count int32[], group int32[]): contiguous pieces of at
    most `chunk` elements that never cross a tensor boundary (pure host logic, unit-tested on CPU)."""
    assert chunk % 4 == 0
    st, cn, gr = [], [], []
    for off, n, gi in ranges:
        if off % 4:
            raise ValueError("parameter tensors must start on a 4-element boundary of the flat buffer")
        for s in range(0, n, chunk):
            st.append(off + s); cn.append(min(chunk, n - s)); gr.append(gi)
    return np.asarray(st, np.int64), np.asarray(cn, np.int32), np.asarray(gr, np.int32)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True, model=None, engine=None,
                 zero_grad=True, chunk=32768):
        if engine is None:
            if model is None:
                raise ValueError("FusedAdamW needs model= (a vilbert_b200 model) or engine=")
            engine = model.engine
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {} - should be in [0.0, 1.0[".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self.engine = engine
        self.fused_zero_grad = bool(zero_grad)
        ps = engine.ps
        dev = ps.flat.device
        self.exp_avg = torch.zeros_like(ps.flat)
        self.exp_avg_sq = torch.zeros_like(ps.flat)
        base, numel = ps.flat.data_ptr(), ps.numel
        ranges, seen = [], set()
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if id(p) in seen:
                    continue     # the tied decoder / word-embedding Parameter is one tensor
                seen.add(id(p))
                if not p.requires_grad:
                    continue
                off = (p.data_ptr() - base) // 4
                if (p.data_ptr() - base) % 4 or off < 0 or off + p.numel() > numel or not p.is_contiguous() or p.dtype != torch.float32:
                    raise ValueError("FusedAdamW: every parameter must be a contiguous fp32 view of the engine's flat buffer")
                ranges.append((off, p.numel(), gi))
                self.state[p] = dict(step=0, exp_avg=self.exp_avg[off:off + p.numel()].view(p.shape),
                                     exp_avg_sq=self.exp_avg_sq[off:off + p.numel()].view(p.shape))
        st, cn, gr = build_chunks(ranges, chunk)
        self.n_chunks = len(st)
        self._chunk_start = torch.from_numpy(st).to(dev)
        self._chunk_count = torch.from_numpy(cn).to(dev)
        self._chunk_group = torch.from_numpy(gr).to(dev)
        self._groups_host = torch.zeros(len(self.param_groups) * _GROUP_DT.itemsize, dtype=torch.uint8).pin_memory() if dev.type == "cuda" else \
            torch.zeros(len(self.param_groups) * _GROUP_DT.itemsize, dtype=torch.uint8)
        self._groups_np = self._groups_host.numpy().view(_GROUP_DT)
        self._groups_dev = torch.zeros_like(self._groups_host, device=dev)
        self._groups_last = None
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        self.grad_scale = 1.0
        self._upload_groups()
        if dev.type == "cuda":
            engine.refresh_weights()         # frozen tensors (not in any group) keep this copy; updated ones are rewritten every step
            engine.shadow_trusted = True

    # ------------------------------------------------------------------ hyper-parameter table
    def _upload_groups(self):
        g = self._groups_np
        for i, grp in enumerate(self.param_groups):
            g[i] = (grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"], 1 if grp["correct_bias"] else 0)
        key = g.tobytes()
        if key != self._groups_last:
            self._groups_dev.copy_(self._groups_host, non_blocking=True)
            self._groups_last = key

    # ------------------------------------------------------------------ stepping
    def launch(self, stream=None):
        """The kernel launch alone (capturable in a CUDA graph): uses the hyper-parameter table and step counter currently on the
        device. `step()` = advance the counter + refresh the table + launch."""
        ps = self.engine.ps
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        L.check(L.lib().vb_adamw_step(ps.flat.data_ptr(), ps.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                      ps.shadow.data_ptr(), ps.shadow_lo.data_ptr() if ps.split else None,
                                      ps.shadow_b.data_ptr() if ps.shadow_b is not ps.shadow else None, 1 if ps.op_dtype == torch.float16 else 0,
                                      self._chunk_start.data_ptr(), self._chunk_count.data_ptr(), self._chunk_group.data_ptr(), self.n_chunks,
                                      self._groups_dev.data_ptr(), self._step_dev.data_ptr(), C.c_float(self.grad_scale),
                                      1 if self.fused_zero_grad else 0, stream), "vb_adamw_step")

    def op(self):
        """(fn, args) for an engine op list (Plan.epilogue): the launch as a plan operation."""
        ps = self.engine.ps
        return (L.lib().vb_adamw_step, (ps.flat.data_ptr(), ps.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        ps.shadow.data_ptr(), ps.shadow_lo.data_ptr() if ps.split else None,
                                      ps.shadow_b.data_ptr() if ps.shadow_b is not ps.shadow else None, 1 if ps.op_dtype == torch.float16 else 0,
                                        self._chunk_start.data_ptr(), self._chunk_count.data_ptr(), self._chunk_group.data_ptr(), self.n_chunks,
                                        self._groups_dev.data_ptr(), self._step_dev.data_ptr(), C.c_float(self.grad_scale),
                                        1 if self.fused_zero_grad else 0))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.step_count += 1
        self._step_dev.add_(1)
        for st in self.state.values():
            st["step"] = self.step_count
        self._upload_groups()
        self.launch()
        eng = self.engine
        eng.shadow_clean = True
        if self.fused_zero_grad:
            eng.grad_clean = True
        return loss

    def zero_grad(self, set_to_none=False):
        """The gradients live in the engine's flat buffer and were zeroed by step(); the views stay attached."""
        self.engine.zero_grad()

    def load_state_dict(self, state_dict):
        """Copies exp_avg / exp_avg_sq INTO the flat state buffers (the default implementation would replace the views)."""
        groups = state_dict["param_groups"]
        params = [p for g in self.param_groups for p in g["params"]]
        ids = [i for g in groups for i in g["params"]]
        for pid, p in zip(ids, params):
            s = state_dict["state"].get(pid)
            if s is None or p not in self.state:
                continue
            self.state[p]["exp_avg"].copy_(s["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
            self.state[p]["step"] = int(s["step"])
            self.step_count = max(self.step_count, int(s["step"]))
        self._step_dev.fill_(self.step_count)
        for g_new, g_old in zip(groups, self.param_groups):
            for k, v in g_new.items():
                if k != "params":
                    g_old[k] = v
        self._upload_groups()
