"""torchrun --nproc-per-node N tools/microbench/exchange_ab.py : cost of the exchange step (flat gradient buffer of
config 2, 37 k floats; and 4.4 M floats = config 5) through the peer-memory kernel and through NCCL, inside CUDA graphs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "regt-gcn_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from regt_b200 import shard as S


def bench(fn, reps=20, inner=10):
    """fn enqueued `inner` times in one CUDA graph; median over reps of (graph time / inner), max over ranks"""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    ts = []
    tiny = torch.zeros(1, device=dev)
    for _ in range(reps):
        dist.all_reduce(tiny)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for n in (37_000, 4_420_000):
    region = S.PeerRegion(n, dev, rank, world)
    flat = torch.zeros(n, device=dev)
    res = {"peer kernel": bench(region.allreduce), "nccl all_reduce": bench(lambda: dist.all_reduce(flat))}
    lin = [torch.nn.Parameter(torch.zeros(n - 4, device=dev))]
    loss = torch.ones((), device=dev)
    for xp in ("peer", "nccl"):
        ex = S.GradExchange(lin, world, transport=xp)
        def seq():
            ex.zero(); ex.add_loss(loss); return ex.sync()
        res[f"zero+add_loss+sync ({ex.transport})"] = bench(seq)
    if rank == 0:
        print(f"n={n} floats, world={world}: " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()), flush=True)
dist.barrier()
os._exit(0)
