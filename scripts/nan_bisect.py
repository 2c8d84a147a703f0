"""Debug helper: train_epoch (fused tail) vs stepwise on a small case under env switches; prints NaN / deviation per configuration."""
import os, subprocess, sys, json
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)

def child():
    sys.path.insert(0, ROOT)
    import importlib, numpy as np, torch
    ops = importlib.import_module("debiasing-multi-modal_b200.ops")
    from oracle import cases
    c = cases.make_case(os.environ.get("CASE", "vitl_b256"))
    D, H = c["D"], c["H"]
    dev = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    X = dev(np.concatenate(c["X"])); y = dev(np.concatenate(c["Y"]), torch.int32); g = dev(np.concatenate(c["G"]), torch.int32)
    n = X.shape[0]
    order = torch.randperm(n, generator=torch.Generator().manual_seed(1)).to(torch.int32).cuda()
    That = ops.normalize_text(dev(c["T_class"]))
    bs = int(os.environ.get("BS", "200")); steps = (n + bs - 1) // bs
    lrs = [0.3, 0.2, 0.1, 0.05, 0.05, 0.05, 0.05, 0.05][:steps]
    a1 = ops.AdapterTensors.from_numpy(c["p_old"]); b1 = ops.TrainBuffers(D, H); s1 = ops.BatchStatsBuffers(steps, 4)
    ops.train_epoch(X, order, bs, y, g, a1, That, 100.0, b1, lrs, s1)
    torch.cuda.synchronize()
    a2 = ops.AdapterTensors.from_numpy(c["p_old"]); b2 = ops.TrainBuffers(D, H); s2 = ops.BatchStatsBuffers(steps, 4)
    for s in range(steps):
        idx = order[s * bs:(s + 1) * bs].contiguous()
        ops.train_step(X, y, g, a2, That, 100.0, b2, lrs[s], s2, slot=s, idx=idx)
    torch.cuda.synchronize()
    out = {}
    p1, p2 = a1.to_numpy(), a2.to_numpy()
    for k in ("W1", "gamma", "W2", "b2", "running_mean"):
        d = np.abs(p1[k].astype(np.float64) - p2[k]).max() / (np.abs(p2[k]).max() + 1e-30)
        out[k] = float(d)
    out["nan_epoch"] = bool(any(np.isnan(v).any() for v in p1.values() if hasattr(v, "dtype") and v.dtype.kind == "f"))
    out["nan_steps"] = bool(any(np.isnan(v).any() for v in p2.values() if hasattr(v, "dtype") and v.dtype.kind == "f"))
    print("RESULT", json.dumps(out))

if __name__ == "__main__":
    if os.environ.get("NB_CHILD"):
        child(); sys.exit(0)
    configs = [{}, {"DBMM_GRAPH": "0"}, {"DBMM_TAIL": "split"}, {"DBMM_TAIL": "serial"}, {"DBMM_PDL": "0"},
               {"DBMM_GRAPH": "0", "DBMM_PDL": "0"}, {"DBMM_GRAPH": "0", "DBMM_TAIL": "serial"},
               {"DBMM_GRAPH": "0", "DBMM_TAIL": "serial", "DBMM_PDL": "0"}, {"BS": "256"}, {"CASE": "rn50_b699", "BS": "699"}]
    for cfg in configs:
        env = dict(os.environ); env.update(cfg); env["NB_CHILD"] = "1"
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(cfg, line[0] if line else ("FAILED rc=%d %s" % (r.returncode, (r.stdout + r.stderr)[-600:])), flush=True)
