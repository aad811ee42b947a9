"""torchrun probe: c5-shaped repartition timings per stage (LFE_TRACE_SHARD=1)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from leanfe_b200 import leanfe
from leanfe_b200.engine import get_engine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = get_engine(local); eng.comm_init_from_torch()
n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
shard = bench.make_c5_rows(torch, dev, n5 * rank // world, n5 * (rank + 1) // world, n5)
kw = dict(formula=bench.C5_FORMULA, vcov="iid", strategy="alt_proj")
for i in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = leanfe(shard, profile=True, **kw)
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"fit": i, "ms": (time.perf_counter() - t0) * 1e3, "timings": {k: round(v, 3) for k, v in r.timings.items() if v}}), flush=True)
dist.destroy_process_group()
