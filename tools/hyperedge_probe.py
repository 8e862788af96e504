"""HGNNLayer.forward / backward of HCCF: libhgr tall-and-skinny kernels vs torch.mm (cuBLAS) on the same GPU.

    python tools/hyperedge_probe.py [--n 30000,41000,1250000] [--k 128] [--d 64]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypergraph_diffusion_for_recommendation_b200 import ops  # noqa: E402


def timed(fn, flush, iters=9):
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="30000,41000,1250000")
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--d", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    for n in (int(v) for v in args.n.split(",")):
        h = (torch.randn(n, args.k, device=dev) * 0.1).requires_grad_(True)
        e = torch.randn(n, args.d, device=dev).requires_grad_(True)
        g = torch.randn(n, args.d, device=dev)

        def ours():
            y = ops.hyperedge(h, e)
            return torch.autograd.grad(y, (h, e), g) + (y.detach(),)

        def lib():
            y = torch.mm(h, torch.mm(h.T, e))
            return torch.autograd.grad(y, (h, e), g) + (y.detach(),)

        def fwd_ours():
            with torch.no_grad():
                return ops.hyperedge(h, e)

        def fwd_lib():
            with torch.no_grad():
                return torch.mm(h, torch.mm(h.T, e))

        for _ in range(3):
            ours(), lib()
        t1, o1 = timed(ours, flush)
        t2, o2 = timed(lib, flush)
        f1, _ = timed(fwd_ours, flush)
        f2, _ = timed(fwd_lib, flush)
        rel = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(o1, o2)]
        print("n %8d K %d D %d | fwd libhgr %.3f ms  torch.mm %.3f ms | fwd+bwd libhgr %.3f ms  torch.mm %.3f ms | max rel diff dH %.1e dE %.1e Y %.1e" % (
            n, args.k, args.d, f1, f2, t1, t2, rel[0], rel[1], rel[2]), flush=True)


if __name__ == "__main__":
    main()
