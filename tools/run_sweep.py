#!/usr/bin/env python
"""BASELINE config 5: an N-episode 5-way K-shot evaluation sweep sharded across the GPUs of one box.

    python tools/run_sweep.py --episodes 10000                                    # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_sweep.py --episodes 10000

Episode e goes to rank e mod world; every rank holds a full weight replica; the only collective is one all-reduce
of [n, sum acc, sum acc^2, sum loss] (clip_spm_b200/sweep.py).  Prints accuracy +- 95 % CI and loss as
run/main_run.py:286-289 does, plus episodes/s (device time, max over ranks)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=1000)
    ap.add_argument("--backbone", default="ViT-B/16", choices=["ViT-B/16", "RN50"])
    ap.add_argument("--way", type=int, default=5)
    ap.add_argument("--shot", type=int, default=5)
    ap.add_argument("--query-per-class", type=int, default=1)
    ap.add_argument("--seq-len", type=int, default=8)
    ap.add_argument("--text-classes", type=int, default=24)
    ap.add_argument("--episodes-per-call", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    __graft_entry__.build()
    from clip_spm_b200 import CNN, sweep
    from clip_spm_b200.config import make_cfg
    if world > 1:
        sweep.bind_to_gpu_cpus(local)   # keep each rank on its GPU's NUMA node
    D = 512 if args.backbone == "ViT-B/16" else 1024
    net = CNN(make_cfg(args.backbone, args.seq_len, False, args.way), max_episodes=args.episodes_per_call, device=dev)
    net.init_random_(seed=0)
    net.text_features_test = torch.randn(args.text_classes, D, generator=torch.Generator().manual_seed(0))
    sweep.run_sweep(net, min(args.episodes, 2 * world * args.episodes_per_call), args.way, args.shot,
                    args.query_per_class, args.text_classes, rank, world, args.episodes_per_call)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = sweep.run_sweep(net, args.episodes, args.way, args.shot, args.query_per_class, args.text_classes, rank,
                          world, args.episodes_per_call)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        res.update(world_size=world, seconds=float(dt), episodes_per_s=args.episodes / float(dt),
                   note="includes on-device synthetic episode generation between calls")
        print("%d episodes: accuracy %.2f +/- %.2f, loss %.4f" % (res["n"], res["accuracy"], res["confidence"], res["loss"]))
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
