"""one rank of the multi-GPU test of the NVLink peer-memory all-reduce (csrc/peer.cu; tests/test_gpu_shard.py).
usage: python peer_allreduce_worker.py RANK WORLD PORT"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist


def main(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from regt_b200 import shard as S
        # one block, a sliver, many blocks with a grid-stride tail, and a two-stage size whose float4 count (65 552) is not
        # divisible by 3, 5, 6 or 7 ranks: the last rank's reduce-scatter slot then extends past n floats of scratch
        for n in (37_000, 4, 1_000_003 // 4 * 4 + 4, 262_208):
            region = S.PeerRegion(n, dev, rank, world)
            g = torch.Generator(device="cpu").manual_seed(1000 * n % 9973 + rank)
            every = torch.empty(world, n, device=dev)
            for it in range(6):                              # the barrier epoch advances from call to call
                mine = (torch.rand(n, generator=g) - 0.5).to(dev)
                region.data.copy_(mine)
                dist.all_gather_into_tensor(every.view(-1), mine)
                ref = every[0].clone()
                for r in range(1, world):                    # the kernel sums in rank order: bit-exact, identical on all ranks
                    ref += every[r]
                region.allreduce()
                torch.cuda.synchronize()
                assert torch.equal(region.data, ref), f"rank {rank} n={n} it={it}: max err {(region.data - ref).abs().max()}"
            # captured in a CUDA graph and replayed (the epoch lives in device memory)
            src = (torch.rand(n, generator=g) - 0.5).to(dev)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                region.data.copy_(src)
                region.allreduce()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                region.data.copy_(src)
                region.allreduce()
            dist.all_gather_into_tensor(every.view(-1), src)
            ref = every[0].clone()
            for r in range(1, world):
                ref += every[r]
            for _ in range(5):
                graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(region.data, ref), f"rank {rank} n={n}: graph replay differs"
            assert region.error() == 0
            dist.barrier()
            region.close()                                   # unmaps the peers' regions, frees the own one
            dist.barrier()
        # the flat gradient buffer of a model rides the same path
        lin = torch.nn.Linear(64, 64).to(dev)
        ex = S.GradExchange(list(lin.parameters()), world)
        assert ex.transport == "peer"
        ex.flat.fill_(float(rank + 1))
        ex.add_loss(torch.tensor(2.0, device=dev))
        loss = ex.sync()
        torch.cuda.synchronize()
        tot = world * (world + 1) / 2
        assert float(loss) == tot + 2.0 * world and float(lin.weight.grad.min()) == tot == float(lin.bias.grad.max())
        # something rebinds .grad to a tensor of its own: sync() must refuse to all-reduce a buffer nobody writes into
        lin.weight.grad = torch.zeros_like(lin.weight)
        try:
            ex.sync()
            raise AssertionError("sync() accepted a detached gradient")
        except RuntimeError as e:
            assert "exchange buffer" in str(e)
        ex.attach(lin.weight)
        ex.sync()
        ex.close()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
