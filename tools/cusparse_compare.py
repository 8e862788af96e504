"""Propagation kernel vs the library path the reference uses on a GPU (torch.sparse.mm -> cuSPARSE), same matrix, same X.

    python tools/cusparse_compare.py [--shapes UxIxE,...]
Prints per shape: libhgr ms, torch COO ms (the reference's tensor type, base/torch_interface.py:8-12), torch CSR ms, and the
relative difference of the results (they differ only by summation order)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypergraph_diffusion_for_recommendation_b200 import graph, ops  # noqa: E402
from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions_device  # noqa: E402


def timeit(fn, flush, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="52000x92000x3000000,1250000x250000x125000000")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for shape in args.shapes.split(","):
        U, I, E = (int(v) for v in shape.split("x"))
        u, i = powerlaw_interactions_device(U, I, E, dev)
        adj = graph.build_norm_adj(u, i, U, I, device=dev)
        n = U + I
        x = torch.randn(n, 64, device=dev)
        rows = torch.repeat_interleave(torch.arange(n, device=dev), adj.indptr[1:] - adj.indptr[:-1])
        coo = torch.sparse_coo_tensor(torch.stack([rows, adj.indices.long()]), adj.values, (n, n))  # uncoalesced flag, like the reference
        csr = torch.sparse_csr_tensor(adj.indptr, adj.indices.long(), adj.values, (n, n))
        y = ops.spmm_raw(adj, x)
        y_coo = torch.sparse.mm(coo, x)
        err = float((y - y_coo).abs().max() / y_coo.abs().max())
        t_h = timeit(lambda: ops.spmm_raw(adj, x), flush)
        t_coo = timeit(lambda: torch.sparse.mm(coo, x), flush)
        t_csr = timeit(lambda: torch.sparse.mm(csr, x), flush)
        print("%s nnz %d | libhgr %.3f ms | torch COO (reference path) %.3f ms (%.1fx) | torch CSR %.3f ms (%.1fx) | max rel diff %.2e" % (
            shape, adj._nnz(), t_h, t_coo, t_coo / t_h, t_csr, t_csr / t_h, err), flush=True)
        del coo, csr, adj, x, y, y_coo


if __name__ == "__main__":
    main()
