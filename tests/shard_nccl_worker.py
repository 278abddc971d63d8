"""one rank of the 2-GPU NCCL test of the region-sharded step (tests/test_gpu_shard.py).
usage: python shard_nccl_worker.py RANK WORLD PORT OUT.pt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist


def main(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from parity_util import W, build_cuda, oracle_step
        from regt_b200 import shard as S
        w = W.make_workload(3)
        state = oracle_step(w, 1)["state"]      # deterministic synthetic parameters
        m = build_cuda(w, state, device=dev)
        x, y = w.inputs(2)
        sm = S.RegionShardedModel(m, w.edge_index.to(dev), [e.to(dev) for e in w.reg_edge_index],
                                  [a.to(dev) for a in w.reg_edge_attr], rank, world)
        loss, out, hid = sm.fused_step(x.to(dev), y.to(dev), gather_outputs=True)
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({"loss": float(loss), "out": out.cpu(), "hid": hid.cpu(),
                        "grads": {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
