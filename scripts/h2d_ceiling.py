"""Aggregate host -> device bandwidth of the box with N ranks copying at once
(VERDICT r1 item 5): run under torchrun with 1, 2, 4, 8 ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 scripts/h2d_ceiling.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                          # noqa: E402
import torch.distributed as dist                      # noqa: E402

import bench                                          # noqa: E402
from genetest_b200 import sharding                    # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bound = 0
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bound = sharding.bind_to_device_cpus(local)
gbs = bench.measure_h2d_ceiling(torch, dist, world, mib=512, reps=10)
cores = sorted(os.sched_getaffinity(0))
print("rank {}: bound to {} cores ({}..{})".format(local, bound or len(cores), cores[0], cores[-1]), flush=True)
if local == 0:
    print("{} rank(s): aggregate H2D {:.1f} GB/s ({:.1f} GB/s per GPU)".format(world, gbs, gbs / world), flush=True)
if world > 1:
    dist.destroy_process_group()
