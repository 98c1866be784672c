"""2-rank worker of tests/test_ddp_gpu.py (launched with torch.distributed.run, one rank per GPU, NCCL).
Checks the data-parallel step of the engine numerically: (A) backward + flat all-reduce after it, (B) the overlapped step
(CUDA-graph segments, tail ranges all-reduced while backward continues) and (C) one GPU on the concatenated batch must agree
(reference semantics: apex DDP averages the gradients over the world, train_tasks.py:490-497)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vilbert_oracle as O          # noqa: E402  (synthetic parameters / inputs)
from vilbert_b200.config import BertConfig     # noqa: E402
from vilbert_b200.ddp import FlatGradAllReducer  # noqa: E402
from vilbert_b200.engine import Engine         # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfgj = json.load(open(os.path.join(ROOT, "tests", "golden", "tiny_b4.json")))["config"]
    cfg = O.make_config(cfgj)
    Bl, Nv, Nt = 4, 33, 24
    P = O.synth_params(cfg, seed=0, device=dev)
    glob = O.synth_inputs(cfg, Bl * world, Nv, Nt, seed=77, device=dev)
    tgt = O.synth_vqa_target(Bl * world, 3129, device=dev)
    keys = ("input_txt", "input_imgs", "image_loc", "token_type_ids", "attention_mask", "image_attention_mask")

    def engine():
        eng = Engine(BertConfig.from_dict(cfgj), dev)
        for k in eng.ps.entries:
            eng.ps.p(k).copy_(P[k])
        eng.refresh_weights()
        return eng

    eng = engine()
    sl = slice(rank * Bl, (rank + 1) * Bl)
    plan = eng.plan(Bl, Nt, Nv, grad_outputs=("vil_prediction",), vqa_loss=True)
    plan.load_inputs(*(glob[k][sl] for k in keys))
    plan.vqa_target.copy_(tgt[sl])
    red = FlatGradAllReducer(eng.ps.grad, n_buckets=4)
    # (A) backward, then all-reduce
    eng.zero_grad(force=True); plan.run_step(); red.allreduce(); torch.cuda.synchronize()
    gA = eng.ps.grad.clone()
    # (B) overlapped: graph segments + tail-range all-reduce on a communication stream
    plan.capture_segments(4)
    comm = torch.cuda.Stream()
    res = {}
    for rep in range(2):
        eng.zero_grad(force=True)
        for w in plan.run_step_overlapped(red.allreduce_range, comm):
            if w is not None:
                w.wait()
        torch.cuda.synchronize()
        gB = eng.ps.grad.clone()
        res[f"A_vs_B_{rep}"] = ((gA - gB).abs().max() / gA.abs().max()).item()
    # the segment ranges tile the flat buffer
    segs = plan.segments
    covered = sorted((lo, hi) for (_, _, lo, hi) in segs if hi > lo)
    res["tiles"] = covered[0][0] == 0 and covered[-1][1] == eng.ps.numel and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # every rank holds the same averaged gradients
    g0 = gA.clone(); dist.broadcast(g0, 0)
    res["ranks_equal"] = bool(torch.equal(g0, gA))
    if rank == 0:
        # (C) one GPU, concatenated batch
        e1 = engine()
        p1 = e1.plan(Bl * world, Nt, Nv, grad_outputs=("vil_prediction",), vqa_loss=True)
        p1.load_inputs(*(glob[k] for k in keys)); p1.vqa_target.copy_(tgt)
        e1.zero_grad(force=True); p1.run_step(); torch.cuda.synchronize()
        gC = e1.ps.grad
        res["A_vs_C_max"] = ((gA - gC).abs().max() / gC.abs().max()).item()
        res["A_vs_C_l2"] = ((gA - gC).norm() / gC.norm()).item()
        worst = 0.0
        for k in e1.ps.entries:
            a, c = eng.ps.g(k), e1.ps.g(k)
            if c.abs().max() > 1e-3 * gC.abs().max():
                worst = max(worst, ((a - c).norm() / c.norm()).item())
        res["A_vs_C_worst_tensor_l2"] = worst
        res["loss_mean_of_ranks_vs_global"] = None
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, plan.loss.clone())
    if rank == 0:
        res["loss_mean_of_ranks_vs_global"] = abs(sum(x.item() for x in losses) / world - p1.loss.item()) / abs(p1.loss.item())
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
