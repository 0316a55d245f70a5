"""Data-parallel check on N GPUs (torchrun): the peer-memory exchange fused with the optimizer (parallel.PeerGradSync,
csrc/peer.cu) against the NCCL all-reduce path (parallel.GradSync) on the same per-rank batches.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_peer_ddp.py [size] [steps]

Asserts: identical parameters on every rank after the steps (both modes), peer and NCCL loss trajectories equal, no
time-out recorded by the flag protocol.  Prints ms/step of both modes (captured steps, max over ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ctunet_b200 as C
from ctunet_b200.parallel import GradSync, PeerGradSync
from ctunet_b200.synthetic import make_training_batch
from ctunet_b200.trainer import TrainStep


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    dist.init_process_group("nccl", device_id=dev)
    img, (sk_t, fl_t) = make_training_batch(2, 2, size, seed=100 + rank, device=dev)
    out = {}
    for mode in ("nccl", "peer"):
        torch.manual_seed(0)
        net = C.UNetSP().to(dev)
        sync = PeerGradSync(net) if mode == "peer" else GradSync(net, deferred=True)
        step = TrainStep(net, "double", 1.0, 1.0, lr=1e-3, scheduler=True, grad_sync=sync, graph=True)
        hist = []
        for it in range(steps):
            hist.append(step(img, (sk_t, fl_t)).tolist())
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for it in range(20):
            step(img, (sk_t, fl_t))
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        flat = torch.cat([p.detach().flatten() for p in net.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        err = sync.error() if mode == "peer" else 0
        out[mode] = (hist, float(ms), same, err)
        if mode == "peer":
            sync.close()
        dist.barrier()
    ok = out["nccl"][2] and out["peer"][2] and out["peer"][3] == 0
    worst = max(abs(a - b) for x, y in zip(out["nccl"][0], out["peer"][0]) for a, b in zip(x, y))
    ok = ok and worst < 2e-3
    if rank == 0:
        print("world %d size %d: nccl %.3f ms/step, peer %.3f ms/step; params identical across ranks: nccl %s peer %s; "
              "peer flag error %d; worst |loss_nccl - loss_peer| %.2e; first/last loss %.5f / %.5f  -> %s"
              % (world, size, out["nccl"][1], out["peer"][1], out["nccl"][2], out["peer"][2], out["peer"][3], worst,
                 out["peer"][0][0][-1], out["peer"][0][-1][-1], "OK" if ok else "FAILED"))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
