"""GPU probe: full model (engine plan) vs the oracle (fp32 and the engine's operand-rounding mode), outputs and gradients,
for each operand precision. Development tool; its log is the evidence behind the tolerances in tests/test_model_gpu.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _gpu_util import model_case
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def show(name, r):
    for mode in ("fp32", "op"):
        if "out_" + mode not in r:
            continue
        oe = r["out_" + mode]
        print(f"--- {name} [{mode} oracle] outputs worst={max(oe.values()):.2e}  " + " ".join(f"{k.replace('_prediction','_p').replace('sequence_output','seq').replace('pooled_output','pool')}={v:.1e}" for k, v in oe.items()))
        if "grad_" + mode in r:
            ge = r["grad_" + mode]
            mx = sorted(((v[0], k) for k, v in ge.items()), reverse=True)
            l2 = sorted(v[1] for v in ge.values())
            print(f"    loss mine {r['loss']:.5f} oracle {r['loss_' + mode]:.5f}; grads: worst max-rel {mx[0][0]:.2e} ({mx[0][1]}), median max-rel {mx[len(mx)//2][0]:.2e}, "
                  f"median rel-L2 {l2[len(l2)//2]:.2e}, p90 rel-L2 {l2[int(len(l2)*0.9)]:.2e}, worst rel-L2 {l2[-1]:.2e}; top5 " + ", ".join(f"{k.split('encoder.')[-1]}={e:.1e}" for e, k in mx[:5]))
    sys.stdout.flush()


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = os.path.join(root, "tests", "golden")
    tiny = json.load(open(os.path.join(gold, "tiny_b4.json")))["config"]
    base22 = json.load(open(os.path.join(gold, "base_2layer_2conect_cfg1.json")))["config"]
    base66 = json.load(open(os.path.join(gold, "base_6layer_6conect_b4.json")))["config"]
    large = json.load(open(os.path.join(root, "vilbert-multi-task_b200", "configs", "bert_large_6layer_6conect.json")))
    which = sys.argv[1:] or ["small", "sizes"]
    t0 = time.time()
    for prec in ("fp16", "fp32", "bf16"):
        print(f"=========== precision {prec}")
        if "small" in which:
            show(f"{prec} tiny B4", model_case(tiny, 4, 11, 9, precision=prec))
            show(f"{prec} tiny tasktok odd B3", model_case(dict(tiny, task_specific_tokens=True), 3, 7, 12, seed=1, precision=prec))
            show(f"{prec} tiny peaked", model_case(tiny, 2, 37, 21, seed=2, qk_scale=8.0, precision=prec))
            show(f"{prec} tiny train step 3", model_case(tiny, 4, 11, 9, precision=prec, train_step=3))
            show(f"{prec} base22 cfg1 B2", model_case(base22, 2, 36, 20, precision=prec))
            show(f"{prec} base66 B8", model_case(base66, 8, 100, 36, precision=prec))
        if "sizes" in which and prec != "bf16":
            show(f"{prec} base66 cfg2 B64 100x36", model_case(base66, 64, 100, 36, precision=prec, oracle_modes=("fp32",)))
            show(f"{prec} base66 cfg3 B64 37x36", model_case(base66, 64, 37, 36, precision=prec, oracle_modes=("fp32",), seed=3))
            show(f"{prec} large cfg4 B32 100x60", model_case(large, 32, 100, 60, precision=prec, oracle_modes=("fp32",), seed=2, names=("vil_logit", "vil_prediction")))
            show(f"{prec} base66 cfg5 B2 306x256 tasktok", model_case(dict(base66, task_specific_tokens=True), 2, 306, 256, precision=prec, oracle_modes=("fp32",), seed=5))
        print("elapsed", time.time() - t0); sys.stdout.flush()
