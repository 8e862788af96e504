"""How fast can this B200 gather 256-byte embedding rows at all?  (measure the ceiling on-box.)

Runs tools/gather_probe.cu (built here with nvcc into tools/libgather_probe.so; not part of libhgr.so):
  * `gp_stream`: coalesced reads of an L2-resident buffer (L2 -> SM fabric peak) and of a 4 GB buffer (DRAM read peak);
  * `gp_gather`: pure row gathers over (a) the indices array of the benchmark graph in stored order - the column
    distribution and L2 hit rate of the product kernel, no CSR walk, no FMA, perfectly balanced groups - and (b)
    uniformly random ids over tables of 64 MB (L2-resident) and the graph's own size (DRAM-resident);
  * the product kernel `hgr_spmm_f32` on the same graph for comparison.

    python tools/gather_probe.py --build-only          # CPU box: compile
    python tools/gather_probe.py --shapes 1250000x250000x125000000
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
SO = os.path.join(HERE, "libgather_probe.so")
SRC = os.path.join(HERE, "gather_probe.cu")

RING = {0: "ring 2x4 rows, 5 blocks/SM (product geometry)", 1: "ring 3x4, 4 blocks", 2: "ring 4x4, 3 blocks", 3: "ring 2x8, 3 blocks",
        4: "ring 3x8, 2 blocks", 5: "ring 4x2, 6 blocks", 6: "ring 2x2, 8 blocks", 7: "ring 2x4, 7 blocks",
        10: "register loads 4 x 6 blocks", 11: "register loads 8 x 4 blocks", 12: "register loads 16 x 2 blocks"}


def build():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                               "-shared", "-o", SO, SRC])
    return SO


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-only", action="store_true")
    ap.add_argument("--shapes", default="52000x92000x3000000,1250000x250000x125000000")
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,10,11,12")
    ap.add_argument("--per-group", default="256,1024")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tma", default="", help="comma list of TMA gather4 variants (0: 2 stages x 5 blocks, 1: 3x4, 2: 4x3, 3: 2x6)")
    ap.add_argument("--skip-streams", action="store_true")
    args = ap.parse_args()
    build()
    if args.build_only:
        return
    import torch

    from hypergraph_diffusion_for_recommendation_b200 import ops
    from hypergraph_diffusion_for_recommendation_b200.synth import norm_adj_from_pairs_torch, powerlaw_interactions_device

    lib = C.CDLL(SO)
    lib.gp_gather.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.gp_gather_tma.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.gp_stream.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 0

    def timeit(fn, iters=args.iters, do_flush=True):
        ts = []
        for _ in range(iters + 1):
            if do_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[1:])
        return ts[len(ts) // 2]

    sink = torch.empty(64, device=dev)
    print("== coalesced streaming reads (ld.global.cg.v4, 148 x 8 blocks of 256 threads)")
    for mb, rep in (() if args.skip_streams else ((32, 64), (64, 32), (96, 24), (4096, 1))):
        buf = torch.empty(mb << 18, device=dev)  # mb MiB of floats
        buf.normal_()
        ms = timeit(lambda: lib.gp_stream(buf.data_ptr(), buf.numel(), rep, 148 * 8, sink.data_ptr(), st), do_flush=mb > 126)
        print("  %5d MiB x %2d passes: %.3f ms -> %.0f GB/s" % (mb, rep, ms, buf.numel() * 4 * rep / ms / 1e6), flush=True)
        del buf

    for shape in args.shapes.split(","):
        U, I, E = (int(v) for v in shape.split("x"))
        u, i = powerlaw_interactions_device(U, I, E, dev)
        adj = norm_adj_from_pairs_torch(u, i, U, I)
        del u, i
        n, nnz = U + I, adj._nnz()
        x = torch.randn(n, 64, device=dev)
        print("== shape %s: %d rows, nnz %d, table %.0f MB, gather bytes %.2f GB" % (shape, n, nnz, n * 256 / 1e6, nnz * 256 / 1e9), flush=True)
        ms = timeit(lambda: ops.spmm_raw(adj, x))
        print("  product hgr_spmm_f32 (schedule %s): %.3f ms -> %.2f TB/s of gathered rows, %.1f Gnnz/s" % (adj.schedule, ms, nnz * 256 / ms / 1e9, nnz / ms / 1e6), flush=True)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        streams = {"graph ids (stored order)": (adj.indices, x)}
        small_rows = (64 << 20) // 256
        streams["uniform ids over a 64 MB table"] = (torch.randint(0, small_rows, (nnz,), device=dev, dtype=torch.int32, generator=gen), x[:small_rows])
        streams["uniform ids over the %d MB table" % (n * 256 // 1000000)] = (torch.randint(0, n, (nnz,), device=dev, dtype=torch.int32, generator=gen), x)
        for name, (ids, tab) in streams.items():
            print("  -- %s" % name)
            for pg in (int(v) for v in args.per_group.split(",")):
                out = torch.empty(((nnz + pg - 1) // pg + 16) * 64, device=dev)
                for v in (int(v) for v in args.variants.split(",")):
                    rc = lib.gp_gather(ids.data_ptr(), nnz, tab.data_ptr(), out.data_ptr(), v, pg, st)
                    if rc:
                        print("     variant %d failed (%d)" % (v, rc))
                        continue
                    torch.cuda.synchronize()
                    ms = timeit(lambda: lib.gp_gather(ids.data_ptr(), nnz, tab.data_ptr(), out.data_ptr(), v, pg, st))
                    print("     %-46s per_group %4d: %.3f ms -> %.2f TB/s, %.1f Gnnz/s" % (RING[v], pg, ms, nnz * 256 / ms / 1e9, nnz / ms / 1e6), flush=True)
                if args.tma:
                    want = torch.empty_like(out)
                    lib.gp_gather(ids.data_ptr(), nnz, tab.data_ptr(), want.data_ptr(), 0, pg, st)
                    for v in (int(v) for v in args.tma.split(",")):
                        for box in (1, 4):
                            out.zero_()
                            rc = lib.gp_gather_tma(ids.data_ptr(), nnz, tab.data_ptr(), tab.shape[0], out.data_ptr(), v, pg, box, st)
                            try:
                                torch.cuda.synchronize()
                            except RuntimeError as e:
                                print("     TMA gather4 variant %d box %d: launch died: %s" % (v, box, str(e)[:200]))
                                raise
                            if rc:
                                print("     TMA gather4 variant %d box %d failed (%d)" % (v, box, rc))
                                continue
                            n_full = (nnz // pg) * 64
                            same = torch.equal(out[:n_full], want[:n_full])
                            close = torch.allclose(out[:n_full], want[:n_full], rtol=1e-4, atol=1e-4)
                            ms = timeit(lambda: lib.gp_gather_tma(ids.data_ptr(), nnz, tab.data_ptr(), tab.shape[0], out.data_ptr(), v, pg, box, st))
                            print("     TMA gather4 variant %d box rows %d          per_group %4d: %.3f ms -> %.2f TB/s, %.1f Gnnz/s  (bit-equal to ring: %s, close: %s)"
                                  % (v, box, pg, ms, nnz * 256 / ms / 1e9, nnz / ms / 1e6, same, close), flush=True)
                            if close:
                                break
                del out
        del adj, x, streams


if __name__ == "__main__":
    main()
