"""Time full-ranking evaluation (csrc/eval_topk.cu) on a synthetic power-law graph of a BASELINE shape.

    python tools/eval_probe.py --shape c4 --k 20 --iters 5 [--engine tensor|simt] [--users N]

Prints one JSON line: users/s over the whole C-ABI call (pack + tcgen05 candidates + rescoring + fallback),
CUDA-event timed on the launching stream, L2 flushed between iterations, plus the candidate statistics.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {"c2": (6040, 3706, 750_000), "c3": (30_000, 41_000, 1_000_000), "c4": (52_000, 92_000, 3_000_000),
          "c5s": (1_250_000, 2_000_000, 125_000_000), "c5e": (1_250_000, 250_000, 125_000_000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c4")
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--engine", default="tensor")
    ap.add_argument("--mode", default="exact")
    ap.add_argument("--users", type=int, default=0, help="evaluate only the first N test users")
    ap.add_argument("--trained", action="store_true", help="make training items score high (mask-heavy case)")
    ap.add_argument("--torch-baseline", action="store_true",
                    help="also time the batched library path: cuBLAS fp32 GEMM + masked fill + torch.topk, 4096 users at a time")
    args = ap.parse_args()
    import torch

    from hypergraph_diffusion_for_recommendation_b200 import evaluation as E
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions_device

    dev = torch.device("cuda", 0)
    U, I, n_train = SHAPES[args.shape]
    u, i = powerlaw_interactions_device(U, I, n_train, dev, seed=1234)
    key = torch.sort(u.to(torch.int64) * I + i.to(torch.int64)).values
    indptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(torch.div(key, I, rounding_mode="floor"), minlength=U), 0, out=indptr[1:])
    indices = (key % I).to(torch.int32)
    torch.manual_seed(0)
    ue = torch.nn.init.xavier_uniform_(torch.empty(U, 64, device=dev))
    ie = torch.nn.init.xavier_uniform_(torch.empty(I, 64, device=dev))
    if args.trained:
        # one LightGCN-like smoothing step: users move towards the mean of their training items
        rows = torch.repeat_interleave(torch.arange(U, device=dev), indptr[1:] - indptr[:-1])
        agg = torch.zeros_like(ue).index_add_(0, rows, ie[indices.long()])
        ue = ue + 4 * agg / (indptr[1:] - indptr[:-1]).clamp(min=1).unsqueeze(1)
    n_test = args.users or U
    test_users = torch.randperm(U, device=dev)[:n_test].to(torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times, stats = [], None
    for it in range(args.iters + 2):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids, sc, stats = E.fullrank_topk(ue, ie, test_users, indptr, indices, args.k, mode=args.mode, engine=args.engine,
                                         return_stats=True)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    base_ms = None
    if args.torch_baseline:
        rows = torch.repeat_interleave(torch.arange(U, device=dev), indptr[1:] - indptr[:-1])
        tt = []
        for it in range(3):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs = []
            for c0 in range(0, n_test, 4096):
                us = test_users[c0:c0 + 4096].long()
                sc_ = ue[us] @ ie.T
                # mask the training items of these users: positions of (user in chunk, item)
                lo, hi = indptr[us], indptr[us + 1]
                cnt = hi - lo
                r = torch.repeat_interleave(torch.arange(us.numel(), device=dev), cnt)
                off = torch.arange(int(cnt.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
                sc_[r, indices[(torch.repeat_interleave(lo, cnt) + off)].long()] = -10e8
                outs.append(torch.topk(sc_, args.k, dim=1).indices)
            e1.record()
            torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1))
        base_ms = sorted(tt)[1]
        agree = float((torch.cat(outs).sort(1).values == ids.long().sort(1).values).all(1).float().mean())
    s = stats.tolist()
    flops = 2.0 * 64 * I * n_test
    print(json.dumps({"shape": args.shape, "n_test": n_test, "n_items": I, "k": args.k, "engine": args.engine, "mode": args.mode,
                      "ms": ms, "users_per_s": n_test / ms * 1e3, "tflops_bf16_equiv": flops / ms / 1e9,
                      "candidates_per_user": s[0] / n_test, "rescored_per_user": s[1] / n_test, "fallback_users": s[2],
                      **({"torch_gemm_topk_ms": base_ms, "speedup_vs_torch": base_ms / ms, "same_item_sets": agree} if base_ms else {})}))


if __name__ == "__main__":
    main()
