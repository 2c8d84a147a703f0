"""Sweep members trained concurrently on ONE GPU (BASELINE config 5): M independent adapters (own weights, momentum,
batch order, statistics, workspace) over the same resident embedding matrix, each member's epoch graph on its own stream.
A single member's step is a chain of latency-bound kernels that leaves most SMs idle; concurrent members fill them.
Prints aggregate embeddings/s for M = 1, 2, 4, 8."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from dbmm.modules import Adapter
N, D, H, G = int(os.environ.get("N", 162770)), 1024, 128, 4
bs, epochs = int(os.environ.get("BS", 1024)), int(os.environ.get("EPOCHS", 3))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
steps = (N + bs - 1) // bs


class Member:
    def __init__(self, seed):
        torch.manual_seed(seed)
        self.ad = Adapter(D, H).to(dev).tensors()
        self.st = ops.BatchStatsBuffers(steps, G, device=dev)
        self.buf = ops.TrainBuffers(D, H, device=dev)
        self.order = torch.randperm(N, device=dev).to(torch.int32)
        self.stream = torch.cuda.Stream(device=dev)
        self.ws = torch.empty(ops.train_workspace_bytes(bs, D, H, 2), dtype=torch.uint8, device=dev)

    def epoch(self):
        with torch.cuda.stream(self.stream):
            ops.train_epoch(X, self.order, bs, y, g, self.ad, That, 100.0, self.buf, [0.01] * steps, self.st, ws=self.ws)


for M in (1, 2, 4, 8):
    members = [Member(s) for s in range(M)]
    for _ in range(2):
        for m in members: m.epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for m in members: m.stream.wait_event(e0)
    for _ in range(epochs):
        for m in members: m.epoch()
    for m in members: torch.cuda.current_stream().wait_stream(m.stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / epochs
    print(f"M={M}: {ms:.3f} ms per round of {M} epochs, {M * N / ms / 1e3:.2f} M emb/s aggregate, {1e3 * ms / steps:.1f} us per step-round")
