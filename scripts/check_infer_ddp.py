"""BASELINE config 4 (sliding-window inference of a 512 x 512 x 256 volume) sharded over the GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_infer_ddp.py [--batch 4]

Every rank holds the same volume and the same seed-0 UNetSP; rank r labels its share of the 32 patches
(parallel.shard_range, no collective in the forward passes) and one all-reduce of the disjoint label volumes is the stitch.
Checks (rank 0): the merged labels are BIT-IDENTICAL to a single-rank pass over all patches on the same GPU; prints the
per-volume time of both (device events, max over ranks) and the aggregate voxels/s.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    dist.init_process_group("nccl", device_id=dev)
    import ctunet_b200 as C
    from ctunet_b200 import preprocess as P

    C.set_compute_dtype("bf16")
    torch.manual_seed(0)
    net = C.UNetSP().to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    D, H, W = 256, 512, 512
    vol = torch.rand(2, D, H, W, device=dev, generator=g)
    vol[1] = (vol[1] > 0.5).float()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.reps):
            out = fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / a.reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    ms_shard, out_shard = timed(lambda: P.sliding_window_argmax(net, vol, 128, a.batch, rank=rank, world=world))
    ms_one, out_one = timed(lambda: P.sliding_window_argmax(net, vol, 128, a.batch))
    same = all(torch.equal(x, y) for x, y in zip(out_shard, out_one))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        vox = D * H * W
        print("sliding-window inference 512x512x256, 32 patches of 128^3, batch %d: %d GPUs %.2f ms/volume (%.2f Gvox/s) "
              "vs 1 GPU %.2f ms (%.2f Gvox/s); sharded labels bit-identical to the single-rank pass on every rank: %s"
              % (a.batch, world, ms_shard, vox / ms_shard / 1e6, ms_one, vox / ms_one / 1e6, bool(int(flag))), flush=True)
    dist.destroy_process_group()
    if not int(flag):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
