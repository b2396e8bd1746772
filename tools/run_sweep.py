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
    ap.add_argument("--check-sharding", action="store_true",
                    help="gather every rank's predictions and compare them, on rank 0, with an unsharded run of the same "
                         "episodes (SURVEY.md 8e); exits non-zero on any difference")
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
    if args.check_sharding:
        res2, table, my_logits = sweep.run_sweep(net, args.episodes, args.way, args.shot, args.query_per_class,
                                                 args.text_classes, rank, world, args.episodes_per_call, True)
        ok = True
        if rank == 0:
            # the unsharded run: this ONE GPU evaluates every rank's shard in turn (same episodes per call as the sharded
            # run, no collective), so equality must be exact
            ref_table, ref_logits, st = None, {}, None
            for r in range(world):
                part, t, lg = sweep.run_sweep(net, args.episodes, args.way, args.shot, args.query_per_class,
                                              args.text_classes, r, world, args.episodes_per_call, True, local_only=True)
                ref_table = t if ref_table is None else torch.maximum(ref_table, t)
                ref_logits.update(lg)
                n_r = part["n"]
                v = torch.tensor([n_r, part["accuracy"] / 100 * n_r, 0.0, part["loss"] * n_r], dtype=torch.float64)
                st = v if st is None else st + v
            ref = dict(n=int(st[0]), accuracy=100.0 * float(st[1] / st[0]), loss=float(st[3] / st[0]))
            same_pred = bool(torch.equal(table, ref_table))
            same_logits = all(torch.equal(v, ref_logits[e]) for e, v in my_logits.items())
            same_stats = all(abs(res2[k] - ref[k]) < 1e-9 for k in ("accuracy", "loss")) and res2["n"] == ref["n"]
            ok = same_pred and same_logits and same_stats and int((table < 0).sum()) == 0
            print(json.dumps({"check_sharding": "ok" if ok else "MISMATCH", "world_size": world, "episodes": args.episodes,
                              "predictions_equal": same_pred, "rank0_logits_bit_equal": same_logits,
                              "reduced_stats_equal": same_stats, "sharded": res2, "unsharded": ref}))
        if world > 1:
            dist.barrier()
        if not ok:
            sys.exit(1)
    if rank == 0:
        res.update(world_size=world, seconds=float(dt), episodes_per_s=args.episodes / float(dt),
                   note="includes on-device synthetic episode generation between calls")
        print("%d episodes: accuracy %.2f +/- %.2f, loss %.4f" % (res["n"], res["accuracy"], res["confidence"], res["loss"]))
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
