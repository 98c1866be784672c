"""GPU probe: full model (engine plan) vs the oracle (fp32 and bf16-operand mode), outputs and gradients. Development tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _gpu_util import model_case
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def show(name, r):
    for mode in ("fp32", "bf16"):
        oe = r["out_" + mode]
        print(f"--- {name} [{mode} oracle] outputs worst={max(oe.values()):.2e}  " + " ".join(f"{k.replace('_prediction','_p').replace('sequence_output','seq').replace('pooled_output','pool')}={v:.1e}" for k, v in oe.items()))
        if "grad_" + mode in r:
            ge = r["grad_" + mode]
            mx = sorted(((v[0], k) for k, v in ge.items()), reverse=True)
            l2 = sorted(v[1] for v in ge.values())
            print(f"    loss mine {r['loss']:.5f} oracle {r['loss_' + mode]:.5f}; grads: worst max-rel {mx[0][0]:.2e} ({mx[0][1]}), median max-rel {mx[len(mx)//2][0]:.2e}, "
                  f"median rel-L2 {l2[len(l2)//2]:.2e}, worst rel-L2 {l2[-1]:.2e}; top5 " + ", ".join(f"{k.split('encoder.')[-1]}={e:.1e}" for e, k in mx[:5]))


if __name__ == "__main__":
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    tiny = json.load(open(os.path.join(gold, "tiny_b4.json")))["config"]
    base22 = json.load(open(os.path.join(gold, "base_2layer_2conect_cfg1.json")))["config"]
    base66 = json.load(open(os.path.join(gold, "base_6layer_6conect_b4.json")))["config"]
    t0 = time.time()
    show("tiny B4", model_case(tiny, 4, 11, 9))
    show("tiny tasktok odd B3", model_case(dict(tiny, task_specific_tokens=True), 3, 7, 12, seed=1))
    show("tiny peaked", model_case(tiny, 2, 37, 21, seed=2, qk_scale=8.0))
    show("tiny vqa-only", model_case(tiny, 4, 11, 9, names=("vil_prediction",)))
    show("base22 cfg1 B2", model_case(base22, 2, 36, 20))
    show("base66 B8", model_case(base66, 8, 100, 36))
    show("base66 B32 tasktok", model_case(dict(base66, task_specific_tokens=True), 32, 101, 23, seed=3))
    print("elapsed", time.time() - t0)
