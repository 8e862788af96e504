"""contrastLoss / InfoNCE forward + backward time on the batch sizes of the named configs (CUDA events, median of 9)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypergraph_diffusion_for_recommendation_b200 import loss_torch  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for n, m in ((30000, 3600), (41000, 4000), (1250000, 65536)):
        e1 = torch.randn(n, 64, device=dev, requires_grad=True)
        e2 = torch.randn(n, 64, device=dev, requires_grad=True)
        nodes = torch.randperm(n, device=dev)[:m]
        ts = []
        for it in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss = loss_torch.contrastLoss(e1, e2, nodes, 0.2)
            loss.backward()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            e1.grad = e2.grad = None
        ms = sorted(ts[3:])[len(ts[3:]) // 2]
        flops = 2.0 * m * m * 64 * (1 + 2 + 2)  # logits once forward, logits + weighted sum in each backward pass
        print("contrastLoss fwd+bwd: table %d x 64, %d picked rows | %.3f ms | %.1f TFLOP/s fp32" % (n, m, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
