"""Time the result-block gather in isolation (run under torchrun)."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x = torch.ones((76, 500000), dtype=torch.float64, device="cuda")
bucket = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
for name, fn in (("gather", lambda: dist.gather(x, bucket, dst=0)),
                 ("all_gather", None), ("sendrecv", None)):
    if name == "all_gather":
        ag = [torch.empty_like(x) for _ in range(world)]
        fn = lambda: dist.all_gather(ag, x)
    if name == "sendrecv":
        def fn():
            if rank == 0:
                for r in range(1, world):
                    dist.recv(bucket[r], src=r)
            else:
                dist.send(x, dst=0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    if rank == 0:
        print("{}: {:.2f} ms per call, {:.1f} GB/s into rank 0".format(
            name, dt * 1e3, x.numel() * 8 * (world - 1) / dt / 1e9), flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
