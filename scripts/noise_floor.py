"""How far is the CUDA path from exact arithmetic, compared with the reference's own fp32 (torch CPU)?
Prints relative errors of step-0 gradients / logits against the fp64 oracle for both."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from oracle import adapter_math as am, cases

gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kernel_cases.npz"))
def rel(a, b): return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
for name in ("rn50_b699", "vitl_b256"):
    c = cases.make_case(name)
    D, H = c["D"], c["H"]
    p64 = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in c["p_old"].items()}
    r64 = am.train_step_single(c["X"][0], c["Y"][0], p64, None, am.normalize_text(c["T_class"].astype(np.float64)), 0.01, c["lr"], dtype=np.float64)
    ad = ops.AdapterTensors.from_numpy(c["p_old"]); buf = ops.TrainBuffers(D, H); st = ops.BatchStatsBuffers(1, 4)
    dev = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).cuda()
    That = ops.normalize_text(dev(c["T_class"]))
    ops.train_step(dev(c["X"][0]), dev(c["Y"][0], torch.int32), dev(c["G"][0], torch.int32), ad, That, 100.0, buf, c["lr"], st)
    g = buf.grads.cpu().numpy(); sl = ops.flat_param_slices(D, H)
    key = {"W1": "layers.0.weight", "gamma": "layers.1.weight", "beta": "layers.1.bias", "W2": "layers.3.weight", "b2": "layers.3.bias"}
    for k, tk in key.items():
        ref32 = gold[f"{name}/s1_grad0/{tk}"].reshape(-1)
        g64 = r64["grads"][k].reshape(-1)
        ours = g[sl[k]]
        if ref32.size != g64.size: g64s, ours_s = g64[::41], ours[::41]
        else: g64s, ours_s = g64, ours
        print(f"{name} d{k}: reference-fp32 vs fp64 {rel(ref32, g64s):.2e}   cuda vs fp64 {rel(ours_s, g64s):.2e}")
    print(f"{name} logits0: reference-fp32 vs fp64 {rel(gold[f'{name}/s1_logits0'], r64['logits']):.2e}")
    # our eval logits with the same (initial) params in eval mode vs fp64
    ad0 = ops.AdapterTensors.from_numpy(c["p_old"])
    lo, _ = ops.eval_fwd(dev(c["Xe"]), None, None, ad0, That, 100.0, None, 128, want_logits=True)
    l64 = am.eval_logits(c["Xe"], p64 | {"running_mean": c["p_old"]["running_mean"].astype(np.float64), "running_var": c["p_old"]["running_var"].astype(np.float64)}, am.normalize_text(c["T_class"].astype(np.float64)), 0.01, dtype=np.float64)
    l32 = am.eval_logits(c["Xe"], c["p_old"], am.normalize_text(c["T_class"]), 0.01, dtype=np.float32)
    print(f"{name} eval logits: numpy-fp32 vs fp64 {rel(l32, l64):.2e}   cuda vs fp64 {rel(lo.cpu().numpy(), l64):.2e}")
