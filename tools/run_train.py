"""Data-parallel training iterations on N GPUs of one node (one process per GPU, NCCL through torch.distributed):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_train.py
Each optimiser step covers TASKS_PER_BATCH synthetic tasks (run/main_run.py:203-209); the ranks split them
(train.shard_tasks), accumulate their gradients, exchange them with ONE bucketed all-reduce (train.allreduce_gradients) and
take the same Adam step.  --check: the exchanged gradients of one more task batch against the gradients rank 0 accumulates alone
over all of its tasks (dropout off for that: the masks are drawn per forward).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def build(dev, way, T, tower, tasks_per_batch):
    from clip_spm_b200 import CNN
    from clip_spm_b200.config import make_cfg
    net = CNN(make_cfg("ViT-B/16", T, False, way, tasks_per_batch=tasks_per_batch), max_episodes=1, device=dev)
    net.init_random_(seed=0)
    text = torch.randn(24, 512, generator=torch.Generator().manual_seed(0))
    net.text_features_test, net.text_features_train = text, text
    net.train_backbone = tower
    net.train()
    return net


def run_steps(net, opt, scaler, tasks, steps, tpb, rank, world, dev, train, sweep, way, shot, T, phase=None):
    for s in range(steps):
        t0 = time.perf_counter()
        for t in train.shard_tasks(tpb, rank, world):
            ep = sweep.synthetic_episode_batch([s * tpb + t], way, shot, 1, T, 24, dev)
            inp = {k: (v if k.endswith("images") else v[0]) for k, v in ep.items()}
            out = net(inp)
            scaler.scale(net.loss(out, inp["target_labels"])).backward()
            tasks[0] += 1
        if phase is not None:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        if world > 1:          # (the single-process reference of --check runs inside an initialised 2-rank job: no exchange)
            train.allreduce_gradients(net.trainable_parameters())
        if phase is not None:
            torch.cuda.synchronize()
            phase["tasks_ms"] = phase.get("tasks_ms", 0.0) + 1e3 * (t1 - t0)
            phase["allreduce_ms"] = phase.get("allreduce_ms", 0.0) + 1e3 * (time.perf_counter() - t1)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tasks-per-batch", type=int, default=4)
    ap.add_argument("--way", type=int, default=5)
    ap.add_argument("--shot", type=int, default=1)
    ap.add_argument("--seq-len", type=int, default=8)
    ap.add_argument("--frozen-tower", action="store_true")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from clip_spm_b200 import optim, sweep, train
    if world > 1:
        sweep.bind_to_gpu_cpus(local)      # each rank on the CPUs local to its GPU (and off the other rank's)
    net = build(dev, args.way, args.seq_len, not args.frozen_tower, args.tasks_per_batch)
    net.train_dropout = not args.check
    params = net.trainable_parameters()
    opt, scaler = optim.Adam(params, lr=1e-5, betas=(0.5, 0.999)), optim.GradScaler(dev, init_scale=256.0)
    tasks = [0]
    run_steps(net, opt, scaler, tasks, 1, args.tasks_per_batch, rank, world, dev, train, sweep, args.way, args.shot, args.seq_len)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    phase = {}
    run_steps(net, opt, scaler, tasks, args.steps, args.tasks_per_batch, rank, world, dev, train, sweep, args.way, args.shot,
              args.seq_len, phase)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    line = {"what": "data-parallel training iterations (ViT-B/16 %s + CLIP-SPM head, %d-way %d-shot, T=%d)" %
            ("frozen tower" if args.frozen_tower else "tower", args.way, args.shot, args.seq_len), "n_gpus": world,
            "tasks_per_batch": args.tasks_per_batch, "optimiser_steps": args.steps,
            "tasks_per_s": args.steps * args.tasks_per_batch / float(dt), "ms_per_optimiser_step": 1e3 * float(dt) / args.steps,
            "trainable_parameters": int(sum(p.numel() for p in params)),
            "rank0_ms_per_step": {k: v / args.steps for k, v in phase.items()}}
    if args.check:
        # one more task batch WITHOUT the optimiser step: the exchanged gradients of the sharded run against the gradients rank 0
        # accumulates alone over all tasks on the same weights (Adam's normalised update would turn a last-bit difference of a
        # near-zero gradient into a full +-lr step, so the gradients themselves are what is compared)
        def accumulate(r, w):
            opt.zero_grad()
            for t in train.shard_tasks(args.tasks_per_batch, r, w):
                ep = sweep.synthetic_episode_batch([777000 + t], args.way, args.shot, 1, args.seq_len, 24, dev)
                inp = {k: (v if k.endswith("images") else v[0]) for k, v in ep.items()}
                scaler.scale(net.loss(net(inp), inp["target_labels"])).backward()
        accumulate(rank, world)
        if world > 1:
            train.allreduce_gradients(params)
        sharded = [p.grad.clone() for p in params]
        if world > 1:
            dist.barrier()
        if rank == 0:
            accumulate(0, 1)
            worst = max(float((a - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-30)) for a, p in zip(sharded, params))
            line.update(check="ok" if worst < 1e-4 else "MISMATCH", worst_gradient_rel_diff=worst, gradients_compared=len(params))
        if world > 1:
            dist.barrier()       # nobody leaves the job while rank 0 is still computing
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
