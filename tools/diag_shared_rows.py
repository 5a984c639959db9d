"""dev diagnostic (run under gpurun): is the /dev/shm mapping of shard.SharedHostRows really pinned, and how fast are copies into it?"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scaling_retriever_b200 import shard
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
host = shard.SharedHostRows(6980, 1000)
print("registered:", host._registered, "is_pinned:", host._raw.is_pinned(), "bytes:", host._raw.numel())
src_s = torch.randn(6980, 1000, device="cuda"); src_i = torch.randint(0, 1 << 40, (6980, 1000), device="cuda")
pin_s = torch.empty((6980, 1000), dtype=torch.float32, pin_memory=True); pin_i = torch.empty((6980, 1000), dtype=torch.int64, pin_memory=True)
for name, (ds, di) in {"shared": (host.scores[:6980], host.ids[:6980]), "pinned": (pin_s, pin_i)}.items():
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ds.copy_(src_s, non_blocking=True); di.copy_(src_i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        print(name, it, f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
t0 = time.perf_counter(); dist.barrier(); torch.cuda.synchronize(); print("barrier", f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
t0 = time.perf_counter(); dist.barrier(); torch.cuda.synchronize(); print("barrier", f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
dist.destroy_process_group()
