"""Two-GPU tests of the row-partitioned path with the FUSED all-gather (propagation epilogue storing into every
rank's gathered table through symmetric memory).  Needs >= 2 CUDA devices on one box; skipped otherwise
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu



def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    import torch.distributed as dist

    from hypergraph_diffusion_for_recommendation_b200 import dist_check

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # the same self-check `bench.py --gpus N` runs and reports as "parity" (the driver's test box has one GPU)
        res = dist_check.parity_check(rank, world, dev)
        assert res["ok"], [k for k, v in res["checks"].items() if not v]
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_fused_all_gather_matches_nccl_and_unsharded():
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
