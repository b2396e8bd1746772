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
RN50_GFLOP_PER_FRAME = 11.59
# DRAM bytes per launch of the dominant kernel come from the committed ncu --set full summary of THIS build's kernels
# (tools/ncu_summary.py --json; dram__bytes_read.sum + dram__bytes_write.sum averaged over the profiled launches of the
# kernel); null when the file is absent -- never a constant typed into this script.
NCU_TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
DOMINANT_KERNEL = "gemm2_tcgen05_kernel"
WORKLOAD = ("CLIP-SPM ViT-B/16 5-way 5-shot Kinetics-shape eval (S=25,Q=5,T=8: 240 frames@224 per episode); "
            "GPU arm bf16 tensor cores, CPU arm fp32")


def bench_config(eps, epc):
    """`config` of the JSON line -- the same dict for the CUDA arm and the reference arm (same workload)."""
    return {"workload": WORKLOAD, "way": WAY, "shot": SHOT, "seq_len": T, "frames_per_episode": FRAMES,
            "episodes_per_step_per_gpu": eps, "episodes_per_call": epc,
            "l2": "step inputs (%.0f MB) exceed the 126 MB L2; no explicit flush" % (eps * FRAMES * 0.602112),
            "weights": "random-init (no checkpoints offline)"}


def ncu_traffic(kernel):
    try:
        d = json.load(open(NCU_TRAFFIC_JSON))
        k = d["kernels"][kernel]
        return float(k["dram_bytes_per_launch"]), "%s (%s, %d launches)" % (os.path.relpath(NCU_TRAFFIC_JSON, ROOT),
                                                                          d.get("source", "ncu --set full"), k["launches"])
    except Exception:
        return None, "no committed ncu --set full summary for this build"


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
class CpuReference:
    """The reference algorithm (oracle port, pinned against the executed reference: oracle/pin_against_reference.py)
    on the host cores.  tower(n) pushes n frames through the ViT-B/16 tower, head() runs one full 5-way 5-shot metric
    head + loss/accuracy; the tower is 99.5 % of the reference's time."""

    def __init__(self, threads):
        from oracle import clipspm_oracle as O
        self.O = O
        torch.set_num_threads(threads)
        self.w = O.make_weights("ViT-B/16", seed=0, protocol="P0")
        self.text = O.make_text_features(N_TEXT, 512, seed=0)
        self.ep = O.make_episode(1000, WAY, SHOT, QPC, T, N_TEXT, images=False)
        self.su, self.qu = O.make_features(1000, S, Q, T, 512, self.ep["context_labels"],
                                           self.ep["target_labels"].float())
        self.imgs = torch.rand(FRAMES, 3, 224, 224, generator=torch.Generator().manual_seed(1))

    def tower(self, n, chunk=None):
        t0 = time.perf_counter()
        with torch.no_grad():
            self.O.vit_forward(self.w, self.imgs[:n], chunk=chunk or n)
        return time.perf_counter() - t0

    def head(self):
        O, ep = self.O, self.ep
        t0 = time.perf_counter()
        with torch.no_grad():
            st = O.head_forward(self.w, self.text, self.su, self.qu, ep["context_labels"], ep["real_support_labels"],
                                ep["real_target_labels"], O.DEFAULT_PARAMS)
            O.loss_and_acc(st["logits"], st["dists"], ep["target_labels"])
        return time.perf_counter() - t0


def run_reference(args):
    """`--impl reference`: each step = `--ref-frames` frames (default 64) through the tower + one full head, scaled to
    an episode; before the steps ONE full 240-frame episode is timed as well, so the line shows how linear the
    per-frame cost is (`full_episode`) instead of resting on the extrapolation alone."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fs = max(1, min(FRAMES, args.ref_frames))
    ref = CpuReference(threads)
    for _ in range(args.warmup):
        ref.tower(fs); ref.head()
    full = None
    if not args.no_full_episode:
        t_full = ref.tower(FRAMES, chunk=fs) + ref.head()
        full = {"seconds": t_full, "episodes_per_s": 1.0 / t_full, "frames": FRAMES}
    tt = th = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tt += ref.tower(fs); th += ref.head()
    wall = time.perf_counter() - t0
    ep_s = 1.0 / ((tt / args.steps) * FRAMES / fs + th / args.steps)
    if full is not None:
        full["sampled_over_full"] = ep_s / full["episodes_per_s"]   # 1.0 = the per-frame cost is linear in the sample
    line = {"impl": "reference", "metric": "episodes_per_sec", "value": ep_s, "unit": "episodes/s",
            "frames_per_s": ep_s * FRAMES, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args.episodes_per_step, args.episodes_per_call),
            "cpu_baseline": {"value": ep_s, "unit": "episodes/s", "cores": threads, "kind": "port",
                             "sample": "per step: %d of the 240 frames of one episode through the ViT-B/16 tower + "
                                       "one full metric head; episode time = tower_time*240/%d + head_time; one "
                                       "full 240-frame episode timed besides (full_episode)" % (fs, fs),
                             "full_episode": full},
            "e2e": {"value": ep_s, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- CUDA arm
def _gemm_profile(lib, _lib, fn, max_records=8192):
    """Run fn() with every GEMM launch event-timed; returns per-flavour {tflops, ms, launches} (flavour = kind*2 +
    (tile N == 256); kind 0 = bf16, 1 = tf32)."""
    flops4 = (ctypes.c_double * 4)(); ms4 = (ctypes.c_double * 4)(); cnt4 = (ctypes.c_int * 4)()
    _lib.check(lib.spm_profile_begin(max_records))
    fn()
    _lib.check(lib.spm_profile_end(flops4, ms4, cnt4))
    names = ["bf16_n128", "bf16_n256", "tf32_n128", "tf32_n256"]
    return {names[i]: {"tflops": (flops4[i] / 1e12) / (ms4[i] / 1e3), "ms": ms4[i], "launches": int(cnt4[i])}
            for i in range(4) if ms4[i] > 0}


def config_leg(lib, _lib, sweep, CNN, make_cfg, dev, peaks, backbone, way, shot, qpc, T_, n_text, eps_per_call, calls,
               gflop_per_frame, label):
    """episodes/s of another BASELINE config (device-resident inputs, CUDA events), with its own roofline sub-record:
    the tower's algorithmic FLOPs per second against the measured sustained bf16 peak, and the GEMM flavours' own
    launch-time throughput from one event-timed call."""
    D = 512 if backbone == "ViT-B/16" else 1024
    net = CNN(make_cfg(backbone, T_, False, way), max_episodes=eps_per_call, device=dev)
    net.init_random_(seed=0)
    net.text_features_test = torch.randn(n_text, D, generator=torch.Generator().manual_seed(0))
    S_, Q_ = way * shot, way * qpc
    frames = (S_ + Q_) * T_
    batches = [sweep.synthetic_episode_batch([10000 + c * eps_per_call + i for i in range(eps_per_call)], way, shot, qpc,
                                             T_, n_text, dev) for c in range(2)]

    def call(b):
        return net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"],
                                    b["real_support_labels"], b["real_target_labels"], eps_per_call, b["target_labels"])
    for c in range(3):
        call(batches[c % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.spm_launch_count()
    e0.record()
    for c in range(calls):
        out = call(batches[c % 2])
    e1.record()
    torch.cuda.synchronize()
    launches = lib.spm_launch_count() - l0
    ms = e0.elapsed_time(e1)
    ep_s = calls * eps_per_call / (ms / 1e3)
    prof = _gemm_profile(lib, _lib, lambda: call(batches[0]))
    tfl = ep_s * frames * gflop_per_frame / 1e3
    rec = {"workload": label, "value": ep_s, "unit": "episodes/s", "frames_per_s": ep_s * frames,
           "episodes_per_call": eps_per_call, "calls_timed": calls, "ms_per_call": ms / calls, "gpu_launches": int(launches),
           "finite": bool(torch.isfinite(out["logits"]).all()),
           "roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16"], "unit": "TFLOP/s",
                        "frac": tfl / peaks["bf16"], "what": "frame-encoder algorithmic FLOPs (%.3f GFLOP/frame) per second "
                        "of the whole step, against the %s sustained bf16 peak" % (gflop_per_frame, peaks["which"]),
                        "gemm_launch_tflops": prof}}
    del net, batches
    torch.cuda.empty_cache()
    return rec


def single_episode_leg(CNN, make_cfg, sweep, dev, way, shot, qpc, n_episodes, label, host_dict=False):
    """The reference caller's own loop (run/main_run.py:266-279 Learner.test): per episode, move the task's tensors to
    the device (prepare_task), ONE model(inputs) call on the dict, loss / accuracy, two .item() host syncs."""
    net = CNN(make_cfg("ViT-B/16", T, False, way), max_episodes=1, device=dev)
    net.init_random_(seed=0)
    net.text_features_test = torch.randn(N_TEXT, 512, generator=torch.Generator().manual_seed(0))
    S_, Q_ = way * shot, way * qpc
    hosts = [sweep.synthetic_episode_batch([20000 + i], way, shot, qpc, T, N_TEXT, "cpu", pin=True) for i in range(4)]

    def one_host(h, nxt=None):
        # the same caller without prepare_task's `.to(device)` of the images: the model takes the pinned host tensors and
        # overlaps their copy with the encoder (CNN.forward -> spm_eval_host); loss / accuracy arrive with the outputs
        out = net({"context_images": h["context_images"], "target_images": h["target_images"],
                   "context_labels": h["context_labels"][0], "real_support_labels": h["real_support_labels"][0],
                   "real_target_labels": h["real_target_labels"][0], "target_labels": h["target_labels"][0],
                   "next_images": None if nxt is None else (nxt["context_images"], nxt["target_images"])})
        return out["loss"].item(), out["acc"].item()

    def one(h, nxt=None):
        if host_dict:
            return one_host(h, nxt)
        inp = {"context_images": h["context_images"].to(dev, non_blocking=True),
               "target_images": h["target_images"].to(dev, non_blocking=True),
               "context_labels": h["context_labels"][0].to(dev), "real_support_labels": h["real_support_labels"][0].to(dev),
               "real_target_labels": h["real_target_labels"][0].to(dev), "target_labels": h["target_labels"][0].to(dev)}
        out = net(inp)                                   # CNN.forward(dict) -> {"logits", "dists"}
        loss, acc = _loss_acc(out, inp["target_labels"], net.tasks_per_batch)
        return loss.item(), acc.item()                   # run/main_run.py:278-279
    for i in range(3):
        one(hosts[i % 4], hosts[(i + 1) % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_episodes):
        one(hosts[i % 4], hosts[(i + 1) % 4])    # look-ahead of one episode, as a prefetching DataLoader provides
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    frames = (S_ + Q_) * T
    rec = {"workload": label, "value": n_episodes / dt, "unit": "episodes/s", "ms_per_episode": 1e3 * dt / n_episodes,
           "frames_per_s": n_episodes * frames / dt, "episodes_timed": n_episodes,
           "call": ("CNN.forward(dict of PINNED HOST tensors + the next episode's images as a look-ahead) per episode: the next "
                    "episode's copy runs behind this one's compute; loss/accuracy returned with the outputs + 2 x .item()") if host_dict else
                   "CNN.forward(dict) per episode from pinned host tensors + loss/accuracy + 2 x .item()",
           "h2d_bytes_per_episode": frames * 3 * 224 * 224 * 4}
    del net, hosts
    torch.cuda.empty_cache()
    return rec


def train_step_leg(CNN, make_cfg, sweep, optim, _lib, dev, way, shot, qpc, tower, n_steps, label):
    """The reference's training iteration (run/main_run.py:245-254 train_task + :207-209 optimiser): model(input) in train mode
    (dropout active, prompt rows of text_features_train), loss, scaler.scale(loss).backward(), scaler.step, scaler.update,
    zero_grad -- through clip_spm_b200's train mode, the library's Adam(betas=(0.5, 0.999)) and GradScaler.  tower=False: the
    frame encoder runs frozen on the bf16 evaluation kernels and only the head is differentiated; tower=True (ViT-B/16): the
    differentiable tf32 tower as well, Adam over every parameter.  Inputs resident on the device, CUDA-event timing."""
    # a Python-driven training step launches ~1500 kernels: keep the process on the CPUs next to its GPU while it runs
    # (unbound, the same step was measured anywhere between 54 and 176 ms); the affinity is restored afterwards
    saved_affinity = os.sched_getaffinity(0)
    bound = sweep.bind_to_gpu_cpus(torch.device(dev).index or 0)
    net = CNN(make_cfg("ViT-B/16", T, False, way), max_episodes=1, device=dev)
    net.init_random_(seed=0)
    text = torch.randn(N_TEXT, 512, generator=torch.Generator().manual_seed(0))
    net.text_features_test, net.text_features_train = text, text
    net.train_backbone = bool(tower)
    net.train()
    ep = sweep.synthetic_episode_batch([30000], way, shot, qpc, T, N_TEXT, dev)
    inp = {k: (v if k.endswith("images") else v[0]) for k, v in ep.items()}
    opt = optim.Adam(net.trainable_parameters(), lr=1e-6, betas=(0.5, 0.999))
    scaler = optim.GradScaler(dev)

    def step():
        out = net(inp)
        loss = net.loss(out, inp["target_labels"])
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = _lib.load().spm_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
    evs[0].record()
    for i in range(n_steps):
        loss = step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]
    # the step is ~600-1500 Python-driven launches and sits close to the launch-bound regime: host jitter shows up as single
    # slow steps, so the median is the figure and min / max are reported beside it
    ms = statistics.median(per_step)
    frames = (way * shot + way * qpc) * T
    rec = {"workload": label, "value": 1e3 / ms, "unit": "episodes/s (training steps/s)", "ms_per_step": ms, "ms_per_step_min_max": [min(per_step), max(per_step)],
           "steps_timed": n_steps,
           "frames_per_step": frames, "trainable_parameters": int(sum(p.numel() for p in net.trainable_parameters())),
           "library_launches_per_step": (_lib.load().spm_launch_count() - l0) / n_steps, "last_loss": float(loss.detach()),
           "arithmetic": ("tf32 tensor-core products (tcgen05 GEMMs, mma.sync attention), fp32 accumulation and state" if tower else
                          "frozen tower: bf16 tcgen05 evaluation kernels; head: tf32 products"),
           "call": "CNN.forward(dict) in train mode + loss + optim.GradScaler.scale(loss).backward() + step + update + zero_grad"}
    rec["cpu_binding"] = ("%d CPUs local to the GPU (NVML)" % len(bound)) if bound else "none"
    os.sched_setaffinity(0, saved_affinity)
    del net, opt, ep, inp
    torch.cuda.empty_cache()
    return rec


def train_step_graph_leg(CNN, make_cfg, sweep, optim, train, dev, way, shot, qpc, tower, n_steps, label):
    """The same training iteration replayed as ONE CUDA graph (train.GraphedStep: forward, loss, backward, Adam through the
    GradScaler; dropout seeds from a device counter incremented inside the graph).  tower=True: from the frames, tower + head;
    tower=False: the head on frame features computed once by the frozen tower (the features are the step's input)."""
    net = CNN(make_cfg("ViT-B/16", T, False, way), max_episodes=1, device=dev)
    net.init_random_(seed=0)
    text = torch.randn(N_TEXT, 512, generator=torch.Generator().manual_seed(0))
    net.text_features_test, net.text_features_train = text, text
    net.train_backbone = bool(tower)
    ep = sweep.synthetic_episode_batch([30001], way, shot, qpc, T, N_TEXT, dev)
    inp = {k: (v if k.endswith("images") else v[0]) for k, v in ep.items()}
    fwd = None
    if not tower:
        with torch.no_grad():
            inp["su"] = net.encode_frames(inp.pop("context_images")).view(1, -1, T, 512)
            inp["qu"] = net.encode_frames(inp.pop("target_images")).view(1, -1, T, 512)
        fwd = lambda n, x: n.head(x["su"], x["qu"], x["context_labels"], x["real_support_labels"], x["real_target_labels"])  # noqa: E731
    net.train()
    opt = optim.Adam(net.trainable_parameters(), lr=1e-6, betas=(0.5, 0.999))
    step = train.GraphedStep(net, opt, optim.GradScaler(dev), inp, forward=fwd, warmup=3)
    for _ in range(3):
        step(inp)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
    evs[0].record()
    for i in range(n_steps):
        loss = step(inp)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]
    ms = statistics.median(per_step)
    rec = {"workload": label, "value": 1e3 / ms, "unit": "episodes/s (training steps/s)", "ms_per_step": ms,
           "ms_per_step_min_max": [min(per_step), max(per_step)], "steps_timed": n_steps,
           "trainable_parameters": int(sum(p.numel() for p in net.trainable_parameters())), "last_loss": float(loss),
           "call": "train.GraphedStep(...)(inputs): one cudaGraphLaunch per training iteration"}
    step.close()
    del net, opt, step, ep, inp
    torch.cuda.empty_cache()
    return rec


def _loss_acc(out, target_labels, tasks_per_batch):
    """run/main_run.py:390-392 + utils/utils.py:174-186,259-264 on the forward's outputs (plain torch, as the caller's
    own _loss_and_acc is; the hot path's fused version is spm_eval)."""
    lg = out["logits"][0]
    loss = torch.nn.functional.cross_entropy(lg, target_labels.long(), reduction="sum") / tasks_per_batch \
        + 0.001 * out["dists"]
    acc = (lg.argmax(-1) == target_labels.long()).float().mean()
    return loss, acc


def eager_gpu_leg(net, dev, episodes, peaks):
    """PyTorch-eager library kernels on the same GPU (tools/eager_baseline.py): the reference's tower modules
    (nn.MultiheadAttention / nn.Linear / nn.Conv2d) with the product model's weights, fp32 and autocast(bf16),
    batched 2 episodes (480 frames) per forward.  Tower only: the metric head is excluded, which favours the baseline
    (its eager head adds ~9000 small kernels per episode)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("spm_eager_baseline", os.path.join(ROOT, "tools", "eager_baseline.py"))
    eb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eb)
    tower = eb.build_from(net.state_dict(), dev)
    imgs = torch.rand(episodes * FRAMES, 3, 224, 224, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    mine = net.encode_frames(imgs[:FRAMES])
    rec = {"what": "reference ViT-B/16 tower under PyTorch eager (ATen MultiheadAttention, cuBLASLt), cudnn.benchmark, "
                   "same weights, %d frames per forward; metric head excluded (upper bound for the baseline)" % (2 * FRAMES),
           "torch": torch.__version__}
    for name, ac, steps in (("autocast_bf16", True, 3), ("fp32", False, 1)):
        fps, feats = eb.tower_frames_per_s(tower, imgs, ac, 2 * FRAMES, steps)
        rec[name] = {"frames_per_s": fps, "episodes_per_s_tower_only": fps / FRAMES,
                     "tflops": fps * VIT_GFLOP_PER_FRAME / 1e3, "frac_of_sustained_bf16_peak": fps * VIT_GFLOP_PER_FRAME / 1e3 / peaks["bf16"],
                     "max_rel_diff_vs_product_features": float((feats[:FRAMES] - mine).abs().max() / feats[:FRAMES].abs().max())}
    del tower, imgs
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--episodes-per-step", type=int, default=16)
    ap.add_argument("--episodes-per-call", type=int, default=8)
    ap.add_argument("--distinct-steps", type=int, default=6, help="distinct step batches resident in HBM (cycled)")
    ap.add_argument("--ref-frames", type=int, default=64)
    ap.add_argument("--no-full-episode", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip eager_gpu / single_episode / config3 / config4 (N=1 only legs)")
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
    from clip_spm_b200 import CNN, _lib, optim, sweep, train
    from clip_spm_b200.config import make_cfg
    # several ranks on one box: keep each rank (and the pinned buffers it allocates) on its GPU's NUMA node
    bound = sweep.bind_to_gpu_cpus(local) if world > 1 else None
    lib = _lib.load()
    EPS, EPC = args.episodes_per_step, args.episodes_per_call
    net = CNN(make_cfg("ViT-B/16", T, False, WAY), max_episodes=EPC, device=dev)
    net.init_random_(seed=0)
    net.text_features_test = torch.randn(N_TEXT, 512, generator=torch.Generator().manual_seed(0))

    # this rank's episodes (global ids: weak scaling, every rank has its own), resident in HBM: `distinct` different
    # step batches are cycled, so consecutive steps never see the same pixels or labels
    distinct = max(1, min(args.distinct_steps, args.steps))
    dev_batches = [sweep.synthetic_episode_batch([(k * world + rank) * EPS + i for i in range(EPS)], WAY, SHOT, QPC, T,
                                                 N_TEXT, dev) for k in range(distinct)]

    def device_step(k):
        b = dev_batches[k % distinct]
        accs, losses = [], []
        for i in range(0, EPS, EPC):
            n = min(EPC, EPS - i)
            o = net.forward_episodes(b["context_images"][i * S * T:(i + n) * S * T], b["context_labels"][i:i + n],
                                     b["target_images"][i * Q * T:(i + n) * Q * T], b["real_support_labels"][i:i + n],
                                     b["real_target_labels"][i:i + n], n, b["target_labels"][i:i + n])
            accs.append(o["acc"]); losses.append(o["loss"])
        return torch.cat(accs), torch.cat(losses)

    # run on a dedicated non-default stream: kernel submission to the legacy NULL stream was seen to stall
    # sporadically (the library takes whatever stream is current: CNN passes torch.cuda.current_stream())
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    for k in range(args.warmup):
        a, l = device_step(k)
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
    # window is kept short (2 of the timed steps) because the event pairs cost ~3 % on the steps that carry them.
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
        a, l = device_step(k)
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
    summary["distinct_episodes"] = world * EPS * distinct
    del dev_batches[1:]

    # ---- e2e: host buffers through the public host call, copies inside the timed region (two distinct pinned step
    # batches alternate, so a step never re-reads the buffers the previous step copied)
    e2e = e2e_u8 = None
    if not args.no_e2e:
        hbs = [sweep.synthetic_episode_batch([(100 + k) * world * EPS + rank * EPS + i for i in range(EPS)], WAY, SHOT,
                                             QPC, T, N_TEXT, "cpu", pin=True) for k in range(2)]
        for hb in hbs:
            for key in ("context_labels", "real_support_labels", "real_target_labels", "target_labels"):
                hb[key] = hb[key].contiguous()

        def call(k):
            hb, nx = hbs[k % 2], hbs[(k + 1) % 2]
            return net.evaluate_host(hb["context_images"], hb["context_labels"], hb["target_images"],
                                     hb["real_support_labels"], hb["real_target_labels"], hb["target_labels"], EPS, WAY,
                                     next_images=(nx["context_images"], nx["target_images"]))
        call(0); call(1)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            r = call(k)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = EPS * (FRAMES * 3 * 224 * 224 * 4 + (2 * S + Q) * 4 + Q * 8)
        d2h = EPS * (Q * WAY * 4 + 3 * 4 + Q * 4)
        e2e = {"value": total_eps / float(dt), "unit": "episodes/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "pipelining": "the host call copies the first chunk of step k+1 while step k's last chunk computes "
                             "(spm_eval_host_set_next); every step's inputs are copied host->device inside the timed region"}
        # ---- same call on DECODED frames (uint8 340x256, the size Kinetics frames are extracted at): the data
        # loader's Resize/CenterCrop/ToTensor runs on the GPU (bit-exact, csrc/frame_transform.cu), H2D shrinks 2.3x
        hl = {k: hbs[0][k] for k in ("context_labels", "real_support_labels", "real_target_labels", "target_labels")}
        del hbs, call
        FH, FW = 256, 340
        g = torch.Generator().manual_seed(1234 + rank)
        su8 = torch.randint(0, 256, (EPS * S * T, FH, FW, 3), dtype=torch.uint8, generator=g).pin_memory()
        qu8 = torch.randint(0, 256, (EPS * Q * T, FH, FW, 3), dtype=torch.uint8, generator=g).pin_memory()
        call8 = lambda: net.evaluate_host_u8(su8, hl["context_labels"], qu8, hl["real_support_labels"],
                                             hl["real_target_labels"], hl["target_labels"], EPS, WAY,
                                             next_images=(su8, qu8))
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
        del su8, qu8

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
    traffic, traffic_src = ncu_traffic(DOMINANT_KERNEL)
    roofline = {"bound": "tensor", "kernel": "gemm2_tcgen05_kernel (bf16, cta_group::2, 256x256 per CTA pair; frame-encoder linears)",
                "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["bf16"]) if achieved else None, "peak_source": peaks["which"] +
                " bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks["which"] == "measured" else "fallback 1590",
                "traffic": traffic, "traffic_unit": "DRAM bytes/launch, mean over the profiled launches of the kernel",
                "traffic_source": traffic_src,
                "launches_timed": int(cnt4[dom]),
                "gemm_share_of_step": gemm_ms_all / max(sum(step_ms[prof_first:prof_first + prof_steps]), 1e-9),
                "profiled_steps": [prof_first, prof_first + prof_steps],
                "profiled_steps_schedule": "single stream (kernels serialised so that events bracket one launch)",
                "step_ms_profiled": step_ms[prof_first:prof_first + prof_steps],
                "whole_step_tflops_executed": value / world * FRAMES * VIT_GFLOP_PER_FRAME_EXECUTED / 1e3,
                "whole_step_tflops_nominal": value / world * FRAMES * VIT_GFLOP_PER_FRAME / 1e3,
                "whole_step_frac_executed": value / world * FRAMES * VIT_GFLOP_PER_FRAME_EXECUTED / 1e3 / peaks["bf16"]}
    line = {"metric": "episodes_per_sec", "value": value, "unit": "episodes/s", "frames_per_s": value * FRAMES,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(EPS, EPC),
            "rank0_cpu_binding": ("%d CPUs local to its GPU (NVML)" % len(bound)) if bound else "none",
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "e2e_u8": e2e_u8, "clocks": clocks,
            "sweep_stats": summary}
    if world == 1 and not args.no_extra_legs:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        legs = {}
        try:
            legs["eager_gpu"] = eager_gpu_leg(net, dev, 4, peaks)
            bf = legs["eager_gpu"]["autocast_bf16"]["episodes_per_s_tower_only"]
            legs["eager_gpu"]["speedup_value_over_eager_bf16"] = value / bf
            if e2e:
                legs["eager_gpu"]["speedup_e2e_over_eager_bf16"] = e2e["value"] / bf
        except Exception as ex:   # a baseline leg must never take the headline down with it
            legs["eager_gpu"] = {"error": repr(ex)[:300]}
        del net
        torch.cuda.empty_cache()
        for key, fn in (
            ("single_episode_config1", lambda: single_episode_leg(CNN, make_cfg, sweep, dev, 5, 1, 1, 40,
                "BASELINE config 1 shape: ViT-B/16 5-way 1-shot, 80 frames, one episode per CNN.forward(dict) call")),
            ("single_episode_config2", lambda: single_episode_leg(CNN, make_cfg, sweep, dev, 5, 5, 1, 24,
                "BASELINE config 2 shape: ViT-B/16 5-way 5-shot, 240 frames, one episode per CNN.forward(dict) call")),
            ("single_episode_config2_host_dict", lambda: single_episode_leg(CNN, make_cfg, sweep, dev, 5, 5, 1, 24,
                "BASELINE config 2 shape, one episode per CNN.forward(dict) call, images handed over as pinned host tensors", True)),
            ("single_episode_config1_host_dict", lambda: single_episode_leg(CNN, make_cfg, sweep, dev, 5, 1, 1, 40,
                "BASELINE config 1 shape, one episode per CNN.forward(dict) call, images handed over as pinned host tensors", True)),
            ("train_step_head_config2", lambda: train_step_leg(CNN, make_cfg, sweep, optim, _lib, dev, 5, 5, 1, False, 12,
                "training iteration at BASELINE config 2 shape (240 frames): frozen ViT-B/16 tower, differentiable CLIP-SPM head")),
            ("train_step_full_config1", lambda: train_step_leg(CNN, make_cfg, sweep, optim, _lib, dev, 5, 1, 1, True, 8,
                "training iteration at BASELINE config 1 shape (80 frames): ViT-B/16 tower AND head differentiated, Adam over all parameters")),
            ("config3", lambda: config_leg(lib, _lib, sweep, CNN, make_cfg, dev, peaks, "ViT-B/16", 5, 1, 1, 16, 24, 8, 6,
                VIT_GFLOP_PER_FRAME_EXECUTED, "BASELINE config 3: ViT-B/16 SSv2-Full shape 5-way 1-shot, T=16 (160 frames), "
                "bidirectional OTAM 16x18")),
            ("config4", lambda: config_leg(lib, _lib, sweep, CNN, make_cfg, dev, peaks, "RN50", 5, 3, 1, 8, 10, 8, 6,
                RN50_GFLOP_PER_FRAME, "BASELINE config 4: RN50 HMDB51 shape 5-way 3-shot (160 frames), D=1024")),
            # the CUDA-graph legs come last: a capture that fails must not be able to disturb the legs above
            ("train_step_full_config1_graph", lambda: train_step_graph_leg(CNN, make_cfg, sweep, optim, train, dev, 5, 1, 1, True, 8,
                "training iteration at BASELINE config 1 shape (80 frames), tower + head, replayed as one CUDA graph")),
            ("train_step_head_config2_graph", lambda: train_step_graph_leg(CNN, make_cfg, sweep, optim, train, dev, 5, 5, 1, False, 12,
                "training iteration of the CLIP-SPM head on the features of a config 2 episode (25 + 5 videos), one CUDA graph"))):
            try:
                legs[key] = fn()
            except Exception as ex:
                legs[key] = {"error": repr(ex)[:300]}
        line.update(legs)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref = CpuReference(threads)
        ref.tower(16); ref.head()                       # warm-up (thread pool, allocator)
        t_full = ref.tower(FRAMES, chunk=60) + ref.head()
        line["cpu_baseline"] = {"value": 1.0 / t_full, "unit": "episodes/s", "cores": threads, "kind": "port",
                                "sample": "one FULL episode (all 240 frames through the oracle ViT-B/16 tower in chunks "
                                          "of 60 + the metric head + loss/accuracy), timed once after a 16-frame warm-up: "
                                          "%.1f s" % t_full}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
