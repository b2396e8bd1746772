#!/usr/bin/env python
"""Headline benchmark: 5-way 5-shot CLIP-SPM (ViT-B/16, 8 frames @224) episode evaluation, episodes/s and frames/s.

    python bench.py --gpus 1 --steps K --warmup W                  # this repo's CUDA path
    torchrun ... bench.py --gpus N ...                              # one rank per GPU, episodes sharded (weak scaling)
    python bench.py --impl reference ...                            # the reference algorithm on the host CPU cores

A "step" = one pass of the hot path (frame encoder + metric head + loss/accuracy) over a batch of
`--episodes-per-step` synthetic episodes per GPU.  `value` is measured with the inputs already resident in HBM
(CUDA events, barrier + synchronize on both sides, max over ranks); `e2e` is the same metric through the public
host-buffer call (CNN.evaluate_host -> spm_eval_host) with pinned host inputs, H2D/D2H copies inside the timed
region.  Inputs of one step (8 x 144.5 MB) are larger than the 126 MB L2, so no explicit L2 flush is needed."""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WAY, SHOT, QPC, T, N_TEXT = 5, 5, 1, 8, 24           # BASELINE.json configs[1]: Kinetics-shape 5-way 5-shot
S, Q = WAY * SHOT, WAY * QPC
FRAMES = (S + Q) * T                                  # 240 frames per episode
VIT_GFLOP_PER_FRAME = 35.127                          # BASELINE.md section 3 (full tower)
# executed: the last block's out-proj / MLP run on the class token only (exact pruning, SURVEY 8d asks to count
# executed FLOPs): 35.127 - (0.2324 + 0.9296 + 0.9296) * 196/197
VIT_GFLOP_PER_FRAME_EXECUTED = 33.046
# DRAM bytes per launch of the dominant kernel (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of
# the four encoder GEMM flavours at 512-frame chunks, M = 100864: qkv 571 MB, out-proj 721, fc 728, proj 1239;
# profiles/r01_ncu_full_gemm2_chunk512_summary.txt).  Algorithmic mean (operands + residual + output once): 856 MB.
GEMM_DRAM_BYTES_PER_LAUNCH_NCU = 815e6
WORKLOAD = "CLIP-SPM ViT-B/16 5-way 5-shot Kinetics-shape eval (S=25,Q=5,T=8: 240 frames@224 per episode), bf16"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm=d.get("hbm_gbs"), which="measured")
    return dict(bf16=1590.0, hbm=6650.0, which="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region through in-process NVML (a polling
    `nvidia-smi -lms` subprocess was measured to stall kernel submission intermittently, so it is only the fallback)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        import threading
        self.index, self.thread, self.stop_evt = index, None, threading.Event()
        # NVML queries were measured to perturb kernel submission when issued every 50 ms (episodes/s dips of 10-30 %
        # in some runs); a few samples per second are enough for a median clock and the throttle-reason set
        self.interval = min(5.0, float(os.environ.get("SPM_BENCH_CLOCK_INTERVAL", "0.4")))
        self.sm, self.mx, self.reasons = [], 0.0, set()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_evt.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_evt.wait(self.interval)  # interruptible: stop() never waits for a full interval

    def start(self):
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return None
        self.stop_evt.set()
        self.thread.join(timeout=10.0)
        if not self.sm:
            return None
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml"}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_leg(frames_sample, threads):
    """The reference algorithm (oracle port, pinned against the executed reference: oracle/pin_against_reference.py)
    on the host cores: `frames_sample` frames through the ViT tower + one full 5-way 5-shot head.  The tower is
    99.5 % of the reference's time, so an episode costs tower_time * 240/frames + head_time."""
    from oracle import clipspm_oracle as O
    torch.set_num_threads(threads)
    w = O.make_weights("ViT-B/16", seed=0, protocol="P0")
    text = O.make_text_features(N_TEXT, 512, seed=0)
    ep = O.make_episode(1000, WAY, SHOT, QPC, T, N_TEXT, images=False)
    su, qu = O.make_features(1000, S, Q, T, 512, ep["context_labels"], ep["target_labels"].float())
    imgs = torch.rand(frames_sample, 3, 224, 224, generator=torch.Generator().manual_seed(1))

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.vit_forward(w, imgs, chunk=frames_sample)
            t1 = time.perf_counter()
            st = O.head_forward(w, text, su, qu, ep["context_labels"], ep["real_support_labels"],
                                ep["real_target_labels"], O.DEFAULT_PARAMS)
            O.loss_and_acc(st["logits"], st["dists"], ep["target_labels"])
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fs = args.ref_frames
    step = cpu_reference_leg(fs, threads)
    for _ in range(args.warmup):
        step()
    tt = th = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = step()
        tt += a; th += b
    wall = time.perf_counter() - t0
    ep_s = 1.0 / ((tt / args.steps) * FRAMES / fs + th / args.steps)
    line = {"impl": "reference", "metric": "episodes_per_sec", "value": ep_s, "unit": "episodes/s",
            "frames_per_s": ep_s * FRAMES, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace(", bf16", ", fp32 CPU"), "way": WAY, "shot": SHOT, "seq_len": T},
            "cpu_baseline": {"value": ep_s, "unit": "episodes/s", "cores": threads, "kind": "port",
                             "sample": "per step: %d of the 240 frames of one episode through the ViT-B/16 tower + "
                                       "one full metric head; episode time = tower_time*240/%d + head_time" % (fs, fs)},
            "e2e": {"value": ep_s, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--episodes-per-step", type=int, default=16)
    ap.add_argument("--episodes-per-call", type=int, default=8)
    ap.add_argument("--ref-frames", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import __graft_entry__
    __graft_entry__.build()
    from clip_spm_b200 import CNN, _lib, sweep
    from clip_spm_b200.config import make_cfg
    # several ranks on one box: keep each rank (and the pinned buffers it allocates) on its GPU's NUMA node
    bound = sweep.bind_to_gpu_cpus(local) if world > 1 else None
    lib = _lib.load()
    EPS, EPC = args.episodes_per_step, args.episodes_per_call
    net = CNN(make_cfg("ViT-B/16", T, False, WAY), max_episodes=EPC, device=dev)
    net.init_random_(seed=0)
    net.text_features_test = torch.randn(N_TEXT, 512, generator=torch.Generator().manual_seed(0))

    # this rank's episodes of one step (global ids: weak scaling, every rank has EPS of its own), resident in HBM
    ids = [rank * EPS + i for i in range(EPS)]
    dev_batch = sweep.synthetic_episode_batch(ids, WAY, SHOT, QPC, T, N_TEXT, dev)

    def device_step():
        accs, losses = [], []
        for i in range(0, EPS, EPC):
            n = min(EPC, EPS - i)
            o = net.forward_episodes(dev_batch["context_images"][i * S * T:(i + n) * S * T],
                                     dev_batch["context_labels"][i:i + n],
                                     dev_batch["target_images"][i * Q * T:(i + n) * Q * T],
                                     dev_batch["real_support_labels"][i:i + n],
                                     dev_batch["real_target_labels"][i:i + n], n, dev_batch["target_labels"][i:i + n])
            accs.append(o["acc"]); losses.append(o["loss"])
        return torch.cat(accs), torch.cat(losses)

    # run on a dedicated non-default stream: kernel submission to the legacy NULL stream was seen to stall
    # sporadically (the library takes whatever stream is current: CNN passes torch.cuda.current_stream())
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    for _ in range(args.warmup):
        a, l = device_step()
        # the sweep's final reduce (torch ops + one all-reduce) is part of the timed region: run it in the warm-up
        # too, so that its first-use lazy kernel loading is not billed to the timed steps
        sweep.summarize(sweep.reduce_stats(sweep.make_stats(a, l)))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.spm_launch_count()
    # per-GEMM CUDA events (roofline.achieved): armed for a window of timed steps in the middle of the region (not
    # the first steps, where the launch queue is still shallow and event intervals include submission latency).
    # The event pool is created here, outside the timed region; arming it later only resets two counters.
    prof_first = min(2, args.steps - 1)
    # While armed the library runs those steps on ONE stream (an event pair must bracket exactly one kernel); the
    # other steps overlap frame chunks / heads on three streams, so the window is kept short (2 of the timed steps).
    prof_steps = max(1, min(2, args.steps - prof_first))
    prof_records = prof_steps * EPS * 140 + 64  # upper bound on GEMM launches of the window
    flops4 = (ctypes.c_double * 4)(); ms4 = (ctypes.c_double * 4)(); cnt4 = (ctypes.c_int * 4)()
    _lib.check(lib.spm_profile_begin(prof_records))
    _lib.check(lib.spm_profile_end(flops4, ms4, cnt4))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    acc_all, loss_all = [], []
    step_ev = [ev0]
    for k in range(args.steps):
        if k == prof_first:
            _lib.check(lib.spm_profile_begin(prof_records))
        a, l = device_step()
        if k == prof_first + prof_steps - 1:
            lib.spm_profile_disarm()  # exactly steps [prof_first, prof_first + prof_steps) carry GEMM events
        acc_all.append(a); loss_all.append(l)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        step_ev.append(e)
    stats = sweep.reduce_stats(sweep.make_stats(torch.cat(acc_all), torch.cat(loss_all)))  # the path's one collective
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    step_ms = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
    launches = lib.spm_launch_count() - launches0
    _lib.check(lib.spm_profile_end(flops4, ms4, cnt4))
    clocks = sampler.stop()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    total_eps = world * EPS * args.steps
    value = total_eps / (ms / 1e3)
    summary = sweep.summarize(stats)

    # ---- e2e: host buffers through the public host call, copies inside the timed region
    e2e = e2e_u8 = None
    if not args.no_e2e:
        hb = sweep.synthetic_episode_batch(ids, WAY, SHOT, QPC, T, N_TEXT, "cpu", pin=True)
        call = lambda: net.evaluate_host(hb["context_images"], hb["context_labels"].contiguous(), hb["target_images"],
                                         hb["real_support_labels"].contiguous(), hb["real_target_labels"].contiguous(),
                                         hb["target_labels"].contiguous(), EPS, WAY,
                                         next_images=(hb["context_images"], hb["target_images"]))
        call(); call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = call()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = EPS * (FRAMES * 3 * 224 * 224 * 4 + (2 * S + Q) * 4 + Q * 8)
        d2h = EPS * (Q * WAY * 4 + 3 * 4 + Q * 4)
        e2e = {"value": total_eps / float(dt), "unit": "episodes/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "pipelining": "the host call copies the first episode of step k+1 while step k's last chunk computes "
                             "(spm_eval_host_set_next); every step's inputs are copied host->device inside the timed region"}
        # ---- same call on DECODED frames (uint8 340x256, the size Kinetics frames are extracted at): the data
        # loader's Resize/CenterCrop/ToTensor runs on the GPU (bit-exact, csrc/frame_transform.cu), H2D shrinks 2.3x
        hl = {k: hb[k] for k in ("context_labels", "real_support_labels", "real_target_labels", "target_labels")}
        del hb, call
        FH, FW = 256, 340
        g = torch.Generator().manual_seed(1234 + rank)
        su8 = torch.randint(0, 256, (EPS * S * T, FH, FW, 3), dtype=torch.uint8, generator=g).pin_memory()
        qu8 = torch.randint(0, 256, (EPS * Q * T, FH, FW, 3), dtype=torch.uint8, generator=g).pin_memory()
        call8 = lambda: net.evaluate_host_u8(su8, hl["context_labels"].contiguous(), qu8,
                                             hl["real_support_labels"].contiguous(),
                                             hl["real_target_labels"].contiguous(), hl["target_labels"].contiguous(),
                                             EPS, WAY, next_images=(su8, qu8))
        call8(); call8()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = call8()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_u8 = {"value": total_eps / float(dt), "unit": "episodes/s", "input": "uint8 frames %dx%d" % (FW, FH),
                  "h2d_bytes_per_step": EPS * (FRAMES * FH * FW * 3 + (2 * S + Q) * 4 + Q * 8),
                  "d2h_bytes_per_step": d2h}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    # dominant kernel = the bf16 tcgen05 GEMM with 256-wide tiles (tag 1: gemm2_tcgen05_kernel, cta_group::2, a
    # 256x256 tile per CTA pair -- plus the few 128x256 one-CTA launches); its algorithmic FLOPs / its launch time
    dom = 1
    achieved = (flops4[dom] / 1e12) / (ms4[dom] / 1e3) if ms4[dom] > 0 else None
    gemm_ms_all = sum(ms4)
    roofline = {"bound": "tensor", "kernel": "gemm2_tcgen05_kernel (bf16, cta_group::2, 256x256 per CTA pair; frame-encoder linears)",
                "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["bf16"]) if achieved else None, "peak_source": peaks["which"] +
                " bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks["which"] == "measured" else "fallback 1590",
                "traffic": GEMM_DRAM_BYTES_PER_LAUNCH_NCU, "traffic_unit": "bytes/launch (ncu, mean of the 4 encoder GEMMs)",
                "launches_timed": int(cnt4[dom]),
                "gemm_share_of_step": gemm_ms_all / max(sum(step_ms[prof_first:prof_first + prof_steps]), 1e-9),
                "profiled_steps": [prof_first, prof_first + prof_steps],
                "profiled_steps_schedule": "single stream (kernels serialised so that events bracket one launch)",
                "step_ms_profiled": step_ms[prof_first:prof_first + prof_steps],
                "whole_step_tflops_executed": value / world * FRAMES * VIT_GFLOP_PER_FRAME_EXECUTED / 1e3,
                "whole_step_tflops_nominal": value / world * FRAMES * VIT_GFLOP_PER_FRAME / 1e3}
    line = {"metric": "episodes_per_sec", "value": value, "unit": "episodes/s", "frames_per_s": value * FRAMES,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "way": WAY, "shot": SHOT, "seq_len": T,
                       "episodes_per_step_per_gpu": EPS, "episodes_per_call": EPC,
                       "rank0_cpu_binding": ("%d CPUs local to its GPU (NVML)" % len(bound)) if bound else "none",
                       "l2": "step inputs (%.0f MB) exceed the 126 MB L2; no explicit flush" % (EPS * FRAMES * 0.602112),
                       "weights": "random-init (no checkpoints offline)"},
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "e2e_u8": e2e_u8, "clocks": clocks,
            "sweep_stats": summary}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_reference_leg(args.ref_frames, threads)
        step()
        a, b = step()
        ep_s = 1.0 / (a * FRAMES / args.ref_frames + b)
        line["cpu_baseline"] = {"value": ep_s, "unit": "episodes/s", "cores": threads, "kind": "port",
                                "sample": "%d of 240 frames of one episode through the oracle ViT-B/16 tower + one "
                                          "full metric head, scaled to one episode" % args.ref_frames}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
