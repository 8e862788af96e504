"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?  Run under torchrun."""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    n = 1 << 26  # 256 MB of fp32
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "multicast", getattr(hdl, "multicast_ptr", None), flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pv = hdl.get_buffer(peer, (n,), torch.float32)
    # read the peer's buffer
    s = float(pv[:1024].sum())
    print(rank, "peer", peer, "read sum", s, "(want %g)" % (1024.0 * (peer + 1)), flush=True)
    hdl.barrier(channel=0)
    # P2P store bandwidth: copy my local tensor into the peer's buffer
    src = torch.full((n,), float(100 + rank), device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        pv.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pv.copy_(src)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, "peer store copy %.3f ms -> %.1f GB/s" % (ms, n * 4 / ms / 1e6), flush=True)
    hdl.barrier(channel=0)
    print(rank, "my buffer now", float(t[0]), "(want %g)" % (100 + (rank - 1) % world), flush=True)
    # NCCL all_gather of the same size for comparison
    out = torch.empty(world * n, device=dev)
    for _ in range(2):
        dist.all_gather_into_tensor(out, src)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        dist.all_gather_into_tensor(out, src)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, "nccl all_gather %.3f ms -> %.1f GB/s received per rank" % (ms, (world - 1) * n * 4 / ms / 1e6), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
