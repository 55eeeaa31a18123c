"""Benchmark of the VINCE hot path on B200:  python bench.py --gpus N --steps K --warmup W [--config C] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input (vince_solver.py:397-428,497-499):
    key-encoder forward (no grad, train-mode BN)  ->  query-encoder forward  ->  fused InfoNCE + metrics against
    [keys || queue]  ->  (N>1: NCCL all-gather of keys)  ->  fused EMA + ring-buffer enqueue.
FORWARD-ONLY: the query-encoder backward + SGD step (vince_solver.py:463-469) is not part of the
timed step on either arm; the JSON line says so in `config.step`.

Workloads = BASELINE.json configs (--config):
    1  ResNet18, 4 views/clip, batch=256 frames, K=65536, D=128, T=0.07
    2  ResNet50, 4 views/clip, batch=256 frames, K=65536, D=128, T=0.2  (MoCoV2 config)   <- default at N=1: the
       largest single-GPU configuration
    3  config 2 per GPU on N GPUs with the key all-gather                                  <- default at N>1
    4  ResNet50 + jigsaw branch (coin flip per step: one encoder sees the 9 patches of every frame), batch=128,
       K=131072

metric  "frames/sec (224^2 multi-view)": encoder-forward frames per second through the full step; both encoders
        count (2*B frames per step per GPU).  `value` has inputs resident in HBM; `e2e` is the same
        through the public API with HOST (pinned) inputs, H2D copies and a D2H read of the loss inside the timed region.
        Input wire format: uint8 HWC frames (what the reference's dataset workers hold before ToTensor + Normalize,
        utils/transforms.py:89-101) with the normalisation fused into the stem packing; `e2e_fp32_wire` is the same
        step fed with the reference's fp32 NCHW tensors (4x the PCIe bytes).
roofline  for the dominant kernel (conv_gemm, tensor-core bound): algorithmic conv FLOPs per launch (2*M*N*K, no credit
        for the 3 fp16 passes) / CUDA-event duration of each launch.
cpu_baseline / --impl reference  the reference's own VinceModel / VinceQueueModel / StorageQueue (oracle/_ref, staged by
        oracle/build_ref.py; kind "reference") on the host cores, per-phase timings; falls back to the oracle port
        (kind "port") if the staged copy is absent.

Multi-GPU (launched by torchrun): each rank owns B frames (weak scaling), one all-gather of keys per step.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(backbone="ResNet18", nf=4, B=256, K=65536, D=128, T=0.07, m=0.999, H=224, jigsaw=False),
    2: dict(backbone="ResNet50", nf=4, B=256, K=65536, D=128, T=0.2, m=0.999, H=224, jigsaw=False),
    3: dict(backbone="ResNet50", nf=4, B=256, K=65536, D=128, T=0.2, m=0.999, H=224, jigsaw=False),
    4: dict(backbone="ResNet50", nf=4, B=128, K=131072, D=128, T=0.2, m=0.999, H=224, jigsaw=True),
}
CONFIG_TEXT = {
    1: "BASELINE.json configs[1]: ResNet18, 4 views/clip, batch=256 frames/GPU, queue K=65536, dim=128, 224x224",
    2: "BASELINE.json configs[2]: ResNet50, 4 views/clip, batch=256 frames/GPU, queue K=65536, dim=128, 224x224 (MoCoV2 config)",
    3: "BASELINE.json configs[3]: ResNet50, 4 views/clip, per-GPU batch=256 frames, queue K=65536, dim=128, 224x224, key all-gather",
    4: "BASELINE.json configs[4]: ResNet50 + jigsaw branch, 4 views + 9 patches (75x75 of the 225-padded frame), batch=128 frames/GPU, queue K=131072, dim=128",
}
METRIC = "frames/sec (224^2 multi-view) through encoder+InfoNCE+EMA/enqueue step"
STEP_TEXT = ("forward-only scoring step: key-encoder fwd + query-encoder fwd + fused InfoNCE/metrics + EMA + enqueue "
             "(no query-encoder backward / SGD on either arm)")


def pick_config(a, world):
    cfg = a.config if a.config else (2 if world == 1 else 3)
    wl = dict(CONFIGS[cfg])
    wl["cfg"] = cfg
    return wl


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_conv_traffic(cfg):
    """Average DRAM bytes (read + write) per conv_gemm launch from the committed ncu capture of one timed step of this
    command and config (profiles/r02_step_metrics_cfg<N>.csv; dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    import csv
    for name in ("r02_step_metrics_cfg%d.csv" % cfg, "r01_step_metrics_all_launches.csv" if cfg == 1 else ""):
        path = os.path.join(ROOT, "profiles", name)
        if not name or not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(open(path, newline="")))
            hdr = rows[0]
            k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ur, uw = units[rows[1][r]], units[rows[1][w]]             # (ncu picks a unit per column)
            vals = [float(x[r].replace(",", "")) * ur + float(x[w].replace(",", "")) * uw for x in rows[2:] if "conv_gemm" in x[k]]
            if vals:
                return round(sum(vals) / len(vals)), len(vals), name
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """The timed region starts now: earlier samples are dropped (the last one is kept as the region's first)."""
        self.lines = self.lines[-1:]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gpu_numa_cpus(gpu_index):
    """CPUs of the NUMA node the GPU hangs off (best effort): pinned host buffers allocated by a thread bound there
    land in that node's memory (first touch), so the H2D copies of 8 ranks do not all cross one socket."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------------------------------
def make_args(dev, wl):
    import vince_b200
    return types.SimpleNamespace(
        backbone=getattr(vince_b200, wl["backbone"]), num_frames=wl["nf"], use_attention=False,
        feature_extractor_gpu_ids=[dev], pytorch_gpu_ids=[dev], vince_embedding_size=wl["D"],
        vince_queue_size=wl["K"], vince_temperature=wl["T"], vince_self_temperature=0.03, vince_momentum=wl["m"],
        jigsaw=wl["jigsaw"], inter_batch_comparison=True, self_batch_comparison=False, batch_size=wl["B"],
        use_imagenet=False)


class HotPath:
    """The step, written against the reference-facing API exactly as VinceSolver.run_train_iteration calls it."""

    def __init__(self, dev, wl, rank, world, gather, local_rank=0):
        import torch

        import vince_b200
        self.torch, self.wl, self.dev, self.world, self.gather = torch, wl, dev, world, gather
        self.args = make_args(dev, wl)
        torch.manual_seed(0)
        self.model = vince_b200.VinceModel(self.args)
        self.model.to(dev)
        self.model.train()
        self.qm = vince_b200.VinceQueueModel(self.args, self.model)
        self.qm.to(dev)
        self.qm.train()
        self.queue = vince_b200.StorageQueue(wl["K"], wl["D"], device=dev)
        g = torch.Generator().manual_seed(1234 + rank)
        # pinned host buffers are first touched by this thread: bind it to the GPU's NUMA node while they are created
        self.numa_node, cpus = gpu_numa_cpus(local_rank)
        old_aff = None
        if cpus:
            try:
                old_aff = os.sched_getaffinity(0)
                os.sched_setaffinity(0, cpus & old_aff or old_aff)
            except OSError:
                old_aff = None
        B, H = wl["B"], wl["H"]
        self.host = {
            "uint8": (torch.randint(0, 256, (B, H, H, 3), generator=g, dtype=torch.uint8).pin_memory(),
                      torch.randint(0, 256, (B, H, H, 3), generator=g, dtype=torch.uint8).pin_memory()),
        }
        if wl.get("fp32_leg", True):
            self.host["fp32"] = (torch.randn((B, 3, H, H), generator=g).pin_memory(),
                                 torch.randn((B, 3, H, H), generator=g).pin_memory())
        if old_aff is not None:
            os.sched_setaffinity(0, old_aff)
        self.devbuf = {k: (v[0].to(dev), v[1].to(dev)) for k, v in self.host.items()}
        self.coin = random.Random(2020 + rank)
        self.opt = None
        self.launches = 0
        self.prefetch = None

    def step(self, data, queue_data, train=False):
        """train=False: the forward-only scoring step (under no_grad, like the CPU arm).  train=True: the whole
        run_train_iteration - the same forward with the backward tape, loss.backward() through the sm_100a backward
        kernels, the fused SGD step, then enqueue + EMA (vince_solver.py:397-499)."""
        if not train:
            with self.torch.no_grad():
                return self._step(data, queue_data, False)
        return self._step(data, queue_data, True)

    def _step(self, data, queue_data, train):
        wl = self.wl
        batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [wl["B"]],
                 "data_source": "synthetic", "num_frames": wl["nf"]}
        if wl["jigsaw"]:                                                              # vince_solver.py:397-403
            key_gets_patches = self.coin.random() < 0.5
            queue_batches = self.qm(batch, jigsaw=key_gets_patches, shuffle=True)
            launches = self.qm.launches
            outputs = self.model.get_embeddings(batch, jigsaw=not key_gets_patches, shuffle=True)
        else:
            queue_batches = self.qm(batch, shuffle=True)                              # vince_solver.py:405
            launches = self.qm.launches
            outputs = self.model.get_embeddings(batch, shuffle=True)                  # :406
        launches += self.model.launches
        output = outputs[0]
        output.update(self.queue.dequeue())                                           # :420
        output.update({"data_source": "synthetic", "num_frames": wl["nf"]})
        output.update(queue_batches[0])
        self.model.launches = 0
        output.update(self.model(output))                                             # :424
        loss_pair = self.model.loss(output)["nce_loss"]                               # :425
        loss = loss_pair[1]
        self.model.get_metrics(output)                                                # :426
        launches += self.model.launches
        if train:                                                                     # :463-469
            if self.opt is None:
                from vince_b200.optim import FusedSGD
                self.opt = FusedSGD(self.model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4)
            self.opt.zero_grad()
            (loss_pair[0] * loss).backward()
            self.opt.step()
            launches += getattr(self.model, "backward_launches", 0) + 1
        keys = output["queue_embeddings"]
        # :497-499 (N>1: the key all-gather, the ring-buffer scatter and the EMA share one call / one kernel)
        self.qm.vince_update(self.model, enqueue=(self.queue, keys, [None] * wl["B"], "synthetic"), gather=self.gather)
        launches += self.qm.launches
        self.launches = launches
        return loss.detach()

    def run_e2e(self, steps, kind):
        """Host (pinned) buffers in, loss on the host out, every step.  As in the reference solver, whose prefetch
        thread copies batch i+1 while batch i trains (vince_solver.py:340-370), the H2D copy of the next step's
        inputs runs on a copy stream underneath the current step; each step's loss is copied back to pinned memory
        and read by the host one step later (so the host never idles the GPU), the last one after the loop."""
        torch = self.torch
        from vince_b200.prefetch import BatchPrefetcher
        if self.prefetch is None:
            self.prefetch = BatchPrefetcher(self.dev, depth=2)
            self.loss_host = torch.empty((1024,), dtype=torch.float32).pin_memory()
        pf = self.prefetch
        host_batch = {"data": self.host[kind][0], "queue_data": self.host[kind][1]}
        events, losses = [], []
        pf.submit(host_batch)
        for i in range(steps):
            if i + 1 < steps:
                pf.submit(host_batch)
            batch = pf.next()
            loss = self.step(batch["data"], batch["queue_data"])
            pf.release(batch)
            self.loss_host[i % 1024].copy_(loss, non_blocking=True)      # D2H read of the step's result
            ev = torch.cuda.Event()
            ev.record()
            events.append(ev)
            if i > 0:
                events[i - 1].synchronize()
                losses.append(float(self.loss_host[(i - 1) % 1024]))
        events[-1].synchronize()
        losses.append(float(self.loss_host[(steps - 1) % 1024]))
        self.h2d_bytes_per_step = pf.h2d_bytes
        return losses


def run_ours(a):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the vince_b200 hot path has no CPU fallback")
    dev = "cuda:%d" % local_rank
    torch.cuda.set_device(local_rank)
    dist = None
    gather = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
        from vince_b200.distributed import KeyGather
        gather = KeyGather(dev)
    from vince_b200 import ops
    wl = pick_config(a, world)
    kind = a.input
    other = "fp32" if kind == "uint8" else "uint8"
    hp = HotPath(dev, wl, rank, world, gather, local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    data, queue_data = hp.devbuf[kind]
    if a.profile_train:              # under ncu: capture exactly one full training step
        for _ in range(3):
            hp.step(data, queue_data, train=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        hp.step(data, queue_data, train=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            emit({"profile_train": True, "config": wl["cfg"]})
        return
    warm = max(a.warmup, 3)
    for _ in range(warm):
        hp.step(data, queue_data)
    torch.cuda.synchronize()
    # the W warm-up steps build the plans and capture the graphs; short steps (ResNet-18: 7 ms) are then through them
    # before the clocks / power management have settled, and the timed region read 5-8 % slow against every later leg:
    # keep stepping (untimed, counted in "warmup") for about another 0.5 s of steady work.  The number of extra steps is
    # the same on every rank (the step contains a collective at N > 1): it is derived from the slowest rank's step time.
    # the clock sampler (an nvidia-smi process polling every 100 ms) is started here, so that its start-up (process launch,
    # NVML initialisation: 100-300 ms of driver calls) falls into the untimed steps, not into a 0.14 s timed region; only the
    # samples taken from the start of the timed region on are kept
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if not a.profile_only:
        t0 = time.perf_counter()
        hp.step(data, queue_data)
        torch.cuda.synchronize()
        t_step = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
        extra = int(min(200, max(0, round(0.5 / max(float(t_step), 1e-4)))))
        for _ in range(extra):
            hp.step(data, queue_data)
        warm += 1 + extra
    # ---- device-resident leg (value) with per-launch events on the tensor-core kernel ----
    sampler.mark()
    torch.cuda.profiler.start()      # ncu --profile-from-start off captures exactly the timed steps (no-op otherwise)
    ms = timed(lambda: hp.step(data, queue_data), a.steps)
    torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    frames_per_step = 2 * wl["B"] * world
    value = frames_per_step * a.steps / (ms / 1e3)
    if a.profile_only:               # under ncu: the launch list / --set full capture of the timed steps is all we want
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            emit({"profile_only": True, "ms_per_step_under_profiler": round(ms / a.steps, 3), "config": wl["cfg"]})
        return
    # ---- the same steps once more with CUDA events around every tensor-core launch (the launch lists then run eagerly
    #      instead of as CUDA graphs, which is why this is not the timed region itself) ----
    ops.PROFILE = []
    ms_prof = timed(lambda: hp.step(data, queue_data), a.steps)
    prof, ops.PROFILE = ops.PROFILE, None
    conv_ms = sum(e0.elapsed_time(e1) for k_, _, _, e0, e1 in prof if k_ == "gemm")
    conv_flops = sum(f for _, f, _, _, _ in prof)
    n_conv = len([1 for k_, _, _, _, _ in prof if k_ == "gemm"])
    # ---- same kernel timed with the two encoders serialised (no co-running streaming kernels of the other encoder) ----
    os.environ["VINCE_B200_OVERLAP"] = "0"
    iso_steps = max(2, min(a.steps, 10))
    hp.step(data, queue_data)
    ops.PROFILE = []
    iso_ms = timed(lambda: hp.step(data, queue_data), iso_steps)
    prof_iso, ops.PROFILE = ops.PROFILE, None
    os.environ["VINCE_B200_OVERLAP"] = "1"
    # roles of the conv_gemm launches: "gemm" = raw fp32 output (+ fused BatchNorm statistics): the tensor-bound role the
    # top-level roofline describes; "apply" = recompute pass with the BatchNorm + residual + ReLU epilogue writing fp16
    # planes, and "stats" = transposed statistics pass: streaming roles, bound by HBM, reported against the copy peak
    def role(kind):
        sel = [(f, b, e0.elapsed_time(e1)) for k_, f, b, e0, e1 in prof_iso if k_ == kind]
        return sum(x[0] for x in sel), sum(x[1] for x in sel), sum(x[2] for x in sel), len(sel)
    iso_conv_flops, _, iso_conv_ms, iso_n = role("gemm")
    iso_all_ms = sum(e0.elapsed_time(e1) for _, _, _, e0, e1 in prof_iso)
    # ---- InfoNCE step (similarity+CE+metrics + [all-gather] + EMA + enqueue) timed alone, device resident ----
    keys = torch.nn.functional.normalize(torch.randn((wl["B"], wl["D"]), device=dev), dim=1)
    qv = torch.nn.functional.normalize(torch.randn((wl["B"], wl["D"]), device=dev), dim=1)

    def nce_step(backward=False):
        out = {"embeddings": qv, "extracted_features": qv, "queue_embeddings": keys, "data_source": "synthetic",
               "num_frames": wl["nf"]}
        out.update(hp.queue.dequeue())
        out.update(hp.model(out))
        hp.model.loss(out)
        hp.model.get_metrics(out)
        if backward:
            hp.model.embedding_gradients(out)          # d loss / d embeddings (what vince_solver.py:465 needs first)
        hp.qm.vince_update(hp.model, enqueue=(hp.queue, keys, [None] * wl["B"], "synthetic"), gather=gather)
    for _ in range(3):
        nce_step()
        nce_step(True)
    nce_ms = timed(nce_step, 20) / 20
    nce_bwd_ms = timed(lambda: nce_step(True), 20) / 20
    # ---- full training step (forward with tape + backward + fused SGD + enqueue + EMA), device resident ----
    train = None
    if not wl["jigsaw"] and not a.no_train:
        try:
            t_steps = max(3, min(a.steps, 10))
            for _ in range(4):          # (tape plan + backward list are built, run eagerly once, then captured as graphs)
                hp.step(data, queue_data, train=True)
            t_ms = timed(lambda: hp.step(data, queue_data, train=True), t_steps)
            train = {"ms_per_step": round(t_ms / t_steps, 4), "value": round(frames_per_step * t_steps / (t_ms / 1e3), 1),
                     "unit": "frames/s", "steps": t_steps, "gpu_launches_per_step": hp.launches,
                     "what": "whole run_train_iteration (vince_solver.py:397-499): both encoder forwards, fused InfoNCE, "
                             "loss.backward() through the sm_100a dgrad / wgrad / BatchNorm-backward kernels, fused "
                             "momentum-SGD, enqueue + EMA; device-resident inputs; gradients are NOT all-reduced in "
                             "this leg"}
            hp.opt.zero_grad()
            hp.model.feature_extractor.module.runner._plans.clear()       # drop the taped plan's activations
            hp.step(data, queue_data)
        except Exception as e:  # noqa: BLE001  (the forward-only metric must survive a failing extra leg)
            train = {"error": str(e)[:300]}
    # ---- end-to-end leg: host (pinned) inputs, H2D inside the timed region, loss read back every step ----
    hp.run_e2e(3, kind)
    e2e_steps = a.steps
    e2e_losses = []
    e2e_ms = timed(lambda: e2e_losses.extend(hp.run_e2e(e2e_steps, kind)), 1)
    e2e_value = frames_per_step * e2e_steps / (e2e_ms / 1e3)
    h2d_main = hp.h2d_bytes_per_step
    assert len(e2e_losses) == e2e_steps and all(l == l for l in e2e_losses)
    # second wire format, fewer steps: an extra number for the record
    e2e_other = None
    if other in hp.host:
        o_steps = max(3, min(a.steps, 10))
        hp.run_e2e(2, other)
        o_ms = timed(lambda: hp.run_e2e(o_steps, other), 1)
        e2e_other = {"value": round(frames_per_step * o_steps / (o_ms / 1e3), 1), "unit": "frames/s",
                     "ms_per_step": round(o_ms / o_steps, 4), "h2d_bytes_per_step": hp.h2d_bytes_per_step,
                     "steps": o_steps}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    traffic = ncu_conv_traffic(wl["cfg"] if wl["cfg"] != 3 else 2)
    achieved_tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak_tf = peaks["tf_sustained"]
    iso_tf = iso_conv_flops / (iso_conv_ms / 1e3) / 1e12 if iso_conv_ms > 0 else 0.0
    step_flops = conv_flops / max(a.steps, 1)
    streaming = {}
    for rkind, what in (("apply", "recompute pass of the two-pass BatchNorm route: GEMM + BatchNorm scale/shift + residual + "
                                 "ReLU -> fp16 planes in the epilogue (Bottleneck 1x1 expansions with K <= 128)"),
                       ("stats", "transposed statistics pass of the same route (nothing stored)")):
        f_, b_, t_, n_ = role(rkind)
        if n_:
            streaming[rkind] = {"bound": "hbm", "launches_timed": n_, "avg_launch_us": round(t_ * 1e3 / n_, 2),
                               "algorithmic_mbytes_per_launch": round(b_ / n_ / 1e6, 1),
                               "achieved": round(b_ / (t_ / 1e3) / 1e9, 1), "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": round(b_ / (t_ / 1e3) / 1e9 / peaks["hbm"], 4),
                               "share_of_step": round(t_ / iso_ms, 4), "what": what}
    roofline = {
        "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, fp16x3), launches in the GEMM role (raw fp32 output + fused "
                  "BatchNorm statistics)", "bound": "tensor",
        # the kernel's own duration: CUDA events around every conv_gemm launch, on its launching stream, with the key /
        # query encoders serialised on one stream (VINCE_B200_OVERLAP=0) in steps run right after the timed region
        "achieved": round(iso_tf, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(iso_tf / peak_tf, 4),
        "traffic": traffic[0] if traffic else None,
        "traffic_source": ("bytes per launch: mean dram__bytes_read.sum + dram__bytes_write.sum over the %d conv_gemm "
                           "launches of one timed step of this workload in profiles/%s (ncu capture of this command; a "
                           "committed measurement, not a counter of this run)" % (traffic[1], traffic[2])) if traffic else None,
        "peak_source": "%s bf16 cuBLAS sustained (same tensor rate as fp16; kernel timed inside a long step)" % peaks["source"],
        "launches_timed": iso_n, "avg_launch_us": round(iso_conv_ms * 1e3 / max(iso_n, 1), 2),
        "algorithmic_gflop_per_launch": round(iso_conv_flops / max(iso_n, 1) / 1e9, 3),
        "share_of_step": round(iso_conv_ms / iso_ms, 4), "ms_per_step": round(iso_ms / iso_steps, 4),
        "all_conv_gemm_launches_share_of_step": round(iso_all_ms / iso_ms, 4),
        "streaming_roles": streaming,
        "whole_step": {"algorithmic_conv_gflop_per_step": round(step_flops / 1e9, 1),
                       "achieved": round(step_flops / (ms / a.steps / 1e3) / 1e12, 2),
                       "frac": round(step_flops / (ms / a.steps / 1e3) / 1e12 / peak_tf, 4),
                       "what": "all conv/linear FLOPs of one step / the device-timed step (two-stream schedule)"},
        "how": "per-launch CUDA events on the launching stream over %d steps with the two encoders serialised on one "
               "stream (VINCE_B200_OVERLAP=0), run right after the timed region: this is the kernel's own duration, "
               "and its share of that step is what the ncu launch list in profiles/ must agree with" % iso_steps,
        "timed_region": {"achieved": round(achieved_tf, 2), "frac": round(achieved_tf / peak_tf, 4),
                         "launches_timed": n_conv, "avg_launch_us": round(conv_ms * 1e3 / max(n_conv, 1), 2),
                         "share_of_step": round(conv_ms / ms_prof, 4),
                         "ms_per_step_with_events": round(ms_prof / a.steps, 4),
                         "what": "the same per-launch events with the two encoders on two streams, as in the timed region "
                                 "(re-run with events, launch lists eager): a conv launch there shares the SMs and HBM with the other encoder's "
                                 "kernels, so its event duration (and share_of_step, which can exceed 1) includes that "
                                 "co-running work"},
        "note": "algorithmic FLOPs = 2*M*N*K of the fp32 conv; the kernel issues 3 fp16 MMAs per k-step to reach "
                "fp32-grade accuracy, so frac <= 1/3 by construction",
    }
    # the CPU leg is a property of the box, not of N: measured on rank 0 of the N=1 run only
    cpu = cpu_baseline(wl, seconds=20.0) if world == 1 else {
        "value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
        "sample": "not measured at N>1 (see the N=1 line of the same run, or --impl reference)"}
    # context leg, opt-in (--ref-gpu): it runs the reference's PyTorch-eager + cuDNN path in this process, and the default
    # bench process should launch nothing but this repo's kernels; the committed bench lines under profiles/ carry it
    ref_gpu = reference_same_gpu(wl, dev) if (world == 1 and a.ref_gpu and not a.no_ref_gpu) else None
    wire = {"uint8": "uint8 HWC raw frames (the dataset workers' format), ToTensor+Normalize fused into the stem packing",
            "fp32": "fp32 NCHW normalised frames (the reference's post-transform wire format)"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": warm, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (fp16x3 split MMA, fp32 accumulate; TF32 for InfoNCE negatives)",
        "data": "synthetic",
        "config": {"workload": CONFIG_TEXT[wl["cfg"]], "step": STEP_TEXT, "per_gpu_batch": wl["B"],
                   "frames_per_step": frames_per_step,
                   "parallelism": "dp%d (replicated weights+queue, NCCL all-gather of keys)" % world if world > 1 else "single GPU",
                   "input": wire[kind],
                   "streams": "key encoder on the caller's stream, query encoder on a side stream (joined before "
                              "get_embeddings returns); VINCE_B200_OVERLAP=0 serialises them",
                   "l2_policy": "inputs larger than L2: >1 GB of activations per encoder forward; queue 32-64 MiB + "
                                "EMA 360-870 MB streamed per step",
                   "host_numa_node": hp.numa_node},
        "infonce_step_ms": round(nce_ms, 4),
        "infonce_step_with_dq_backward_ms": round(nce_bwd_ms, 4),
        "infonce_step_what": "similarity+CE+metrics + %sEMA + enqueue, device resident, B=%d K=%d D=%d" % (
            "key all-gather + " if world > 1 else "", wl["B"], wl["K"], wl["D"]),
        "train_step": train,
        "reference_same_gpu": ref_gpu,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "frames/s", "ms_per_step": round(e2e_ms / e2e_steps, 4),
                "h2d_bytes_per_step": h2d_main, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "h2d_gbs_per_rank": round(h2d_main / (e2e_ms / e2e_steps / 1e3) / 1e9, 2),
                "pipeline": "H2D of step i+1 on a copy stream under step i (vince_b200.prefetch.BatchPrefetcher, "
                            "mirrors the solver's prefetch thread); loss of step i read by the host during step i+1"},
        ("e2e_fp32_wire" if other == "fp32" else "e2e_uint8_wire"): e2e_other,
        "gpu_launches": hp.launches * a.steps,
        "gpu_launches_per_step": hp.launches,
    }
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own modules (oracle/_ref) or, failing that, the oracle port - the only places bench.py
# touches oracle/
# ----------------------------------------------------------------------------------------------------------
class _Phases:
    def __init__(self):
        self.t = {}

    def add(self, name, dt):
        self.t[name] = self.t.get(name, 0.0) + dt


def reference_step_runner(wl, B):
    """The reference's own VinceModel / VinceQueueModel / StorageQueue on CPU, driven exactly as
    vince_solver.py:397-428,497-499 drives them (forward-only, like the GPU arm).  Returns (step, phases, kind)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    torch.set_num_threads(os.cpu_count())
    ph = _Phases()
    g = torch.Generator().manual_seed(1234)
    H = wl["H"]
    data = torch.randn((B, 3, H, H), generator=g)
    queue_data = torch.randn((B, 3, H, H), generator=g)
    coin = random.Random(2020)
    if ref_loader.reference_available():
        import warnings
        warnings.simplefilter("ignore")
        ref = ref_loader.load_reference()
        args = ref_loader.make_args(backbone=wl["backbone"], num_frames=wl["nf"], batch_size=B, queue_size=wl["K"],
                                    embedding_size=wl["D"], temperature=wl["T"], momentum=wl["m"], jigsaw=wl["jigsaw"])
        torch.manual_seed(0)
        model = ref.VinceModel(args)
        model.train()
        qm = ref.VinceQueueModel(args, model)
        qm.train()
        queue = ref.StorageQueue(wl["K"], wl["D"], device="cpu")

        def step():
            batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [B],
                     "data_source": "synthetic", "num_frames": wl["nf"]}
            with torch.no_grad():       # forward-only on both arms (the solver itself keeps the query graph for backward)
                t0 = time.perf_counter()
                kj = wl["jigsaw"] and coin.random() < 0.5
                queue_batches = qm(batch, jigsaw=kj, shuffle=True)
                t1 = time.perf_counter()
                outputs = model.get_embeddings(batch, jigsaw=wl["jigsaw"] and not kj, shuffle=True)
                t2 = time.perf_counter()
                output = outputs[0]
                output.update(queue.dequeue())
                output.update({"data_source": "synthetic", "num_frames": wl["nf"]})
                output.update(queue_batches[0])
                output.update(model(output))
                loss = model.loss(output)["nce_loss"][1]
                model.get_metrics(output)
                t3 = time.perf_counter()
                queue.enqueue(output["queue_embeddings"], [None] * B, "synthetic")
                qm.vince_update(model)
                t4 = time.perf_counter()
            ph.add("key_encoder_fwd", t1 - t0), ph.add("query_encoder_fwd", t2 - t1)
            ph.add("similarity_ce_metrics", t3 - t2), ph.add("enqueue_ema", t4 - t3)
            return float(loss)
        return step, ph, "reference"
    import vince_oracle as vo
    sd_q = vo.make_state_dict(wl["backbone"], wl["D"], jigsaw=wl["jigsaw"], seed=0)
    sd_k = vo.clone_state_dict(sd_q)
    queue = vo.StorageQueue(wl["K"], wl["D"])
    names = vo.vince_parameter_names(sd_q, wl["jigsaw"])

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            perm_k, perm_q = torch.randperm(B), torch.randperm(B)
            out = vo.train_step(data, queue_data, sd_q, sd_k, queue, wl["backbone"], wl["nf"], wl["T"], wl["m"],
                                shuffle_q=perm_q, shuffle_k=perm_k, ema_names=names)
        ph.add("whole_step", time.perf_counter() - t0)
        return float(out["losses"]["nce_loss"])
    return step, ph, "port"


def reference_same_gpu(wl, dev, steps=5):
    """For context (not the contract's reference arm): the UNMODIFIED reference (oracle/_ref: PyTorch eager + cuDNN,
    torch's default math modes - TF32 allowed for convolutions, fp32 matmuls) running the same forward-only step on this
    very GPU, full batch, CUDA-event timed.  Answers "what would the reference's own GPU path do on a B200"."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_loader
        if not ref_loader.reference_available():
            return None
        import warnings
        warnings.simplefilter("ignore")
        ref = ref_loader.load_reference()
        B = wl["B"]
        args = ref_loader.make_args(backbone=wl["backbone"], num_frames=wl["nf"], batch_size=B, queue_size=wl["K"],
                                    embedding_size=wl["D"], temperature=wl["T"], momentum=wl["m"], jigsaw=wl["jigsaw"])
        args.feature_extractor_gpu_ids = [dev]
        args.pytorch_gpu_ids = [dev]
        torch.manual_seed(0)
        model = ref.VinceModel(args)
        model.to(dev)
        model.train()
        qm = ref.VinceQueueModel(args, model)
        qm.to(dev)
        qm.train()
        queue = ref.StorageQueue(wl["K"], wl["D"], device=dev)
        g = torch.Generator().manual_seed(1234)
        data = torch.randn((B, 3, wl["H"], wl["H"]), generator=g).to(dev)
        queue_data = torch.randn((B, 3, wl["H"], wl["H"]), generator=g).to(dev)
        coin = random.Random(2020)

        def step():
            batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [B],
                     "data_source": "synthetic", "num_frames": wl["nf"]}
            with torch.no_grad():
                kj = wl["jigsaw"] and coin.random() < 0.5
                queue_batches = qm(batch, jigsaw=kj, shuffle=True)
                outputs = model.get_embeddings(batch, jigsaw=wl["jigsaw"] and not kj, shuffle=True)
                output = outputs[0]
                output.update(queue.dequeue())
                output.update({"data_source": "synthetic", "num_frames": wl["nf"]})
                output.update(queue_batches[0])
                output.update(model(output))
                model.loss(output)
                model.get_metrics(output)
                queue.enqueue(output["queue_embeddings"], [None] * B, "synthetic")
                qm.vince_update(model)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del model, qm, queue
        torch.cuda.empty_cache()
        return {"value": round(2 * B / (ms / 1e3), 1), "unit": "frames/s", "ms_per_step": round(ms, 3), "steps": steps,
                "what": "the unmodified reference (oracle/_ref; PyTorch %s eager + cuDNN, default math modes: TF32 "
                        "convolutions allowed) running the same forward-only step on this GPU at batch=%d - context "
                        "only; the contract's reference arm is the CPU run" % (torch.__version__, B)}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def cpu_sample_batch(wl):
    """Largest batch (<= the workload's) whose forward-only step fits comfortably in host RAM and a bounded time."""
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
    except (ValueError, OSError):
        avail_gb = 16.0
    B = wl["B"]
    per_frame_gb = 0.12 if wl["backbone"] == "ResNet50" else 0.04      # peak live activations of a no_grad forward
    if wl["jigsaw"]:
        per_frame_gb *= 1.3
    while B > 16 and B * per_frame_gb > 0.4 * avail_gb:
        B //= 2
    return B


def run_cpu_steps(wl, B, warmup, max_steps, seconds):
    step, ph, kind = reference_step_runner(wl, B)
    for _ in range(warmup):
        step()
    ph.t.clear()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        step()
        n += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    phases = {k + "_ms": round(v / n * 1e3, 1) for k, v in ph.t.items()}
    return n, dt, phases, kind


def cpu_sample_text(wl, B, n, dt, kind):
    what = ("the reference's own VinceModel/VinceQueueModel/StorageQueue/loss_util (oracle/_ref, unmodified)" if kind == "reference"
            else "oracle port of the reference step (oracle/vince_oracle.py)")
    return ("%s on torch CPU fp32 with %d threads, forward-only step (no_grad) at batch=%d frames%s, K=%d, D=%d, %dx%d; "
            "%d steps in %.1f s after warm-up" % (what, os.cpu_count(), B,
                                                 "" if B == wl["B"] else " (bounded sample of the batch=%d workload)" % wl["B"],
                                                 wl["K"], wl["D"], wl["H"], wl["H"], n, dt))


def cpu_baseline(wl, seconds):
    B = cpu_sample_batch(wl)
    if wl["backbone"] == "ResNet50":
        B = min(B, 64)               # ~1.4 s per step: keeps the default bench run within minutes
    n, dt, phases, kind = run_cpu_steps(wl, B, 1, 50, seconds)
    return {"value": round(2 * B * n / dt, 2), "unit": "frames/s", "cores": os.cpu_count(), "kind": kind,
            "sample": cpu_sample_text(wl, B, n, dt, kind), "phases": phases}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on this box's host cores, same metric and
    config; each step is the workload's full batch when host RAM allows, else a bounded sample; at most ~150 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = pick_config(a, world)
    B = cpu_sample_batch(wl)
    warm = max(1, min(a.warmup, 1))
    n, dt, phases, kind = run_cpu_steps(wl, B, warm, max(1, a.steps), 150.0)
    v = round(2 * B * n / dt, 2)
    sample = cpu_sample_text(wl, B, n, dt, kind)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": world,
            "steps": n, "warmup": warm, "ms_per_step": round(dt / n * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_TEXT[wl["cfg"]] + "; CPU batch=%d frames" % B, "step": STEP_TEXT},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind, "sample": sample,
                             "phases": phases},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_JSON_FD = None


def emit(obj):
    """The ONE JSON line goes to the process's original stdout; everything else printed during the run (NCCL's version
    banner, library chatter) was re-routed to stderr by main()."""
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)              # C-level and Python-level stdout -> stderr for the rest of the run
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs index; 0 (default) = 2 on one GPU (largest single-GPU config), 3 under torchrun")
    ap.add_argument("--input", default="uint8", choices=["fp32", "uint8"],
                    help="uint8 = raw HWC frames, normalisation fused into the stem packing (default: the format the "
                         "reference's workers hold); fp32 = the reference's normalised NCHW tensors")
    ap.add_argument("--no-train", action="store_true", help="skip the extra full-training-step leg")
    ap.add_argument("--ref-gpu", action="store_true",
                    help="also time the unmodified reference (PyTorch eager + cuDNN) on this GPU, for context")
    ap.add_argument("--no-ref-gpu", action="store_true", help="(default; kept for older scripts)")
    ap.add_argument("--profile-train", action="store_true", help="one full training step between profiler start/stop (ncu)")
    ap.add_argument("--profile-only", action="store_true",
                    help="stop after the device-resident timed steps (for runs under ncu; prints no bench value)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
