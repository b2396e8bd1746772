"""Episode sweep (the caller side of the path: run/main_run.py:256-293 `Learner.test`): episodes are independent
units, so rank r of a world of n evaluates episodes r, r+n, r+2n, ... with a full weight replica and no
intra-episode exchange; the only collective is one all-reduce (SUM) of four fp64 sufficient statistics
[n, sum acc, sum acc^2, sum loss] at the end, from which accuracy, the 95 % confidence interval and the mean loss
of run/main_run.py:286-289 follow."""
import math

import torch


def shard_episodes(n_episodes, rank, world_size):
    """Global episode indices owned by `rank` (round-robin: e -> e mod world_size)."""
    return list(range(rank, n_episodes, world_size))


def bind_to_gpu_cpus(gpu_index):
    """One process per GPU: restrict this process to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe
    root), so the pinned staging buffers it allocates afterwards are first-touched next to the GPU and the host->device
    copies of 8 ranks do not cross the socket interconnect.  Returns the CPU set it bound to, or None when nothing was
    changed (no NVML, no usable CPUs inside the current cpuset).  Never raises."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber the GPUs: resolve the CUDA device through its PCI address
            pr = torch.cuda.get_device_properties(int(gpu_index))
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def make_stats(acc, loss):
    """Per-rank sufficient statistics from per-episode accuracy / loss tensors (any device)."""
    acc, loss = acc.double().flatten(), loss.double().flatten()
    return torch.stack([torch.tensor(float(acc.numel()), dtype=torch.float64, device=acc.device), acc.sum(),
                        (acc * acc).sum(), loss.sum()])


def reduce_stats(stats, group=None):
    """The path's single collective: all-reduce(SUM) of 32 bytes (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats):
    """accuracy = mean*100, confidence = 196*std/sqrt(n) (population std, np.std), loss = mean
    (run/main_run.py:286-289)."""
    n, s1, s2, sl = [float(v) for v in stats.tolist()]
    mean = s1 / n
    var = max(s2 / n - mean * mean, 0.0)
    return dict(n=int(n), accuracy=100.0 * mean, confidence=196.0 * math.sqrt(var) / math.sqrt(n), loss=sl / n)


def synthetic_episode_batch(episode_ids, way, shot, query_per_class, T, n_text_cls, device, pin=False):
    """Synthetic episodes of the benchmark shape, seeded by GLOBAL episode index so that any sharding evaluates the
    same episodes.  Pixels are iid U[0,1) float32 (the sampler's format: video_reader.py:65,271); labels follow
    video_reader.py:312-326 (shuffled float tensors).  Returns stacked tensors (dim 0 = episode-major)."""
    S, Q = way * shot, way * query_per_class
    E = len(episode_ids)
    dev = torch.device(device)
    su = torch.empty(E * S * T, 3, 224, 224, device=dev, pin_memory=pin and dev.type == "cpu")
    qu = torch.empty(E * Q * T, 3, 224, 224, device=dev, pin_memory=pin and dev.type == "cpu")
    lab, rs, rt, tl = [], [], [], []
    for i, e in enumerate(episode_ids):
        g = torch.Generator(device=dev).manual_seed(1000 + int(e))
        su[i * S * T:(i + 1) * S * T].uniform_(0, 1, generator=g)
        qu[i * Q * T:(i + 1) * Q * T].uniform_(0, 1, generator=g)
        gc = torch.Generator().manual_seed(1000 + int(e))
        sl = torch.arange(way).repeat_interleave(shot)
        sl = sl[torch.randperm(sl.numel(), generator=gc)]
        ql = torch.arange(way).repeat_interleave(query_per_class)
        ql = ql[torch.randperm(ql.numel(), generator=gc)]
        cmap = torch.randperm(n_text_cls, generator=gc)[:way]
        lab.append(sl.float()); rs.append(cmap[sl].float()); rt.append(cmap[ql].float()); tl.append(ql.long())
    mk = lambda xs: torch.stack(xs)
    out = dict(context_images=su, target_images=qu, context_labels=mk(lab), real_support_labels=mk(rs),
               real_target_labels=mk(rt), target_labels=mk(tl))
    if dev.type != "cpu":
        out = {k: v.to(dev) for k, v in out.items()}
    return out


def gather_predictions(pred, episode_ids, n_episodes, group=None):
    """Per-episode predictions of every rank, in global episode order (the optional gather used to check
    that a sharded sweep predicts exactly what the unsharded one does).  pred [n_local, Q] int; rows of episodes this
    rank does not own are -1 before the exchange, so one all-reduce(MAX) assembles the table on every rank."""
    import torch.distributed as dist
    Q = pred.shape[1] if pred.dim() == 2 else 0
    table = torch.full((n_episodes, Q), -1, dtype=torch.int32, device=pred.device)
    if len(episode_ids):
        table[torch.as_tensor(list(episode_ids), device=pred.device, dtype=torch.long)] = pred.to(torch.int32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(table, op=dist.ReduceOp.MAX, group=group)
    return table


def run_sweep(net, n_episodes, way, shot, query_per_class, n_text_cls, rank=0, world_size=1, episodes_per_call=1,
              return_predictions=False, local_only=False):
    """Evaluate this rank's shard on the CUDA path and return the REDUCED statistics (identical on every rank); with
    return_predictions also the gathered [n_episodes, Q] prediction table and this rank's logits (by episode id)."""
    mine = shard_episodes(n_episodes, rank, world_size)
    accs, losses, preds, logits = [], [], [], {}
    for i in range(0, len(mine), episodes_per_call):
        ids = mine[i:i + episodes_per_call]
        b = synthetic_episode_batch(ids, way, shot, query_per_class, net.seq_len, n_text_cls, net._dev)
        out = net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"],
                                   b["real_support_labels"], b["real_target_labels"], len(ids), b["target_labels"])
        accs.append(out["acc"]); losses.append(out["loss"])
        if return_predictions:
            preds.append(out["pred"])
            for j, e in enumerate(ids):
                logits[e] = out["logits"][j].clone()
    dev = net._dev
    acc = torch.cat(accs) if accs else torch.zeros(0, device=dev)
    loss = torch.cat(losses) if losses else torch.zeros(0, device=dev)
    stats = make_stats(acc, loss)
    res = summarize(stats if local_only else reduce_stats(stats))
    if not return_predictions:
        return res
    Q = way * query_per_class
    pred = torch.cat(preds) if preds else torch.zeros(0, Q, dtype=torch.int32, device=dev)
    if local_only:
        table = torch.full((n_episodes, Q), -1, dtype=torch.int32, device=dev)
        if mine:
            table[torch.as_tensor(mine, device=dev, dtype=torch.long)] = pred.to(torch.int32)
        return res, table, logits
    return res, gather_predictions(pred, mine, n_episodes), logits


def run_listing_sweep(net, split, load_frame, n_episodes, way, shot, n_queries, seed=0, episodes_per_call=8, rank=0,
                      world_size=1, jpeg=False):
    """The reference's test loop (run/main_run.py:256-293 `Learner.test` over `VideoDataset` episodes) on this library,
    from DECODED frames: episode e is sampled with `frames.sample_episode_plan` (rng seeded `seed + e`, so any sharding
    sees the same episodes), its frames come from `load_frame(handle) -> uint8 [H, W, 3]` (the handles stored in
    `split.videos`), the Resize/CenterCrop/ToTensor chain, the encoder, the head, loss and accuracy run on the GPU
    (`CNN.evaluate_host_u8`), and while one batch of episodes computes the first chunk of the next is already being
    copied (`next_images`).  All frames must share one size.  Returns the reduced statistics of `summarize`.
    jpeg=True: `load_frame(handle)` returns the JPEG FILE CONTENTS (bytes) instead of decoded pixels -- the files are
    decoded on the GPU (CNN.evaluate_jpeg -> spm_jpeg_decode, bit-identical to the PIL decode of video_reader.py:227-230),
    so the whole input path of the reference's loader from the file on runs in the library."""
    import random
    from . import frames as F
    T = net.seq_len
    mine = shard_episodes(n_episodes, rank, world_size)

    def build(ids):
        su, qu, lab, rs, rt, tl = [], [], [], [], [], []
        for e in ids:
            plan = F.sample_episode_plan(split, way, shot, n_queries, T, train=False, rng=random.Random(seed + e))
            su += [load_frame(split.videos[v][f]) for v, fr in plan["support"] for f in fr]
            qu += [load_frame(split.videos[v][f]) for v, fr in plan["target"] for f in fr]
            lab.append(plan["support_labels"]); rs.append(plan["real_support_labels"])
            rt.append(plan["real_target_labels"]); tl.append([int(x) for x in plan["target_labels"]])
        pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
        return dict(su=pin(torch.stack([torch.as_tensor(x) for x in su]).contiguous()),
                    qu=pin(torch.stack([torch.as_tensor(x) for x in qu]).contiguous()),
                    lab=torch.tensor(lab, dtype=torch.float32), rs=torch.tensor(rs, dtype=torch.float32),
                    rt=torch.tensor(rt, dtype=torch.float32), tl=torch.tensor(tl, dtype=torch.int64), n=len(ids))

    batches = [mine[i:i + episodes_per_call] for i in range(0, len(mine), episodes_per_call)]
    accs, losses = [], []
    if jpeg:
        for ids in batches:
            su, qu, lab, rs, rt, tl = [], [], [], [], [], []
            for e in ids:
                plan = F.sample_episode_plan(split, way, shot, n_queries, T, train=False, rng=random.Random(seed + e))
                su += [load_frame(split.videos[v][f]) for v, fr in plan["support"] for f in fr]
                qu += [load_frame(split.videos[v][f]) for v, fr in plan["target"] for f in fr]
                lab.append(plan["support_labels"]); rs.append(plan["real_support_labels"])
                rt.append(plan["real_target_labels"]); tl.append([int(x) for x in plan["target_labels"]])
            out = net.evaluate_jpeg(su, torch.tensor(lab, dtype=torch.float32), qu, torch.tensor(rs, dtype=torch.float32),
                                    torch.tensor(rt, dtype=torch.float32), torch.tensor(tl, dtype=torch.int64), len(ids))
            accs.append(out["acc"].cpu()); losses.append(out["loss"].cpu())
        batches = []
    cur = build(batches[0]) if batches else None
    for bi in range(len(batches)):
        nxt = build(batches[bi + 1]) if bi + 1 < len(batches) else None
        out = net.evaluate_host_u8(cur["su"], cur["lab"], cur["qu"], cur["rs"], cur["rt"], cur["tl"], cur["n"], way,
                                   next_images=None if nxt is None else (nxt["su"], nxt["qu"]))
        accs.append(out["acc"]); losses.append(out["loss"])
        cur = nxt
    acc = torch.cat(accs) if accs else torch.zeros(0)
    loss = torch.cat(losses) if losses else torch.zeros(0)
    stats = make_stats(acc, loss)
    if torch.cuda.is_available():
        stats = stats.to(net._dev)
    return summarize(reduce_stats(stats))
