"""Benchmark of the VINCE hot path on B200:  python bench.py --gpus N --steps K --warmup W  [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input (vince_solver.py:405-428,497-499):
    key-encoder forward (no grad, train-mode BN)  ->  query-encoder forward  ->  fused InfoNCE + metrics against
    [keys || queue]  ->  (N>1: NCCL all-gather of keys)  ->  fused EMA + ring-buffer enqueue.
Workload at N=1 = BASELINE.json configs[1]: ResNet18, 4 views/clip, batch=256 frames, queue K=65536, dim=128.

metric  "frames/sec (224^2 multi-view)": encoder-forward frames per second through the full step; both encoders
        count (2*B frames per step per GPU).  `value` has inputs resident in HBM; `e2e` is the same
        through the public API with HOST (pinned) inputs, H2D copies and a D2H read of the loss inside the timed region.
roofline  for the dominant kernel (conv_gemm, tensor-core bound): algorithmic conv FLOPs per launch (2*M*N*K, no credit
        for the 3 fp16 passes) / CUDA-event duration of each launch, measured during the timed region.
cpu_baseline  the oracle port (oracle/vince_oracle.py = the reference's algorithm on torch CPU) on a bounded sample.

Multi-GPU (launched by torchrun): each rank owns B frames (weak scaling), one all-gather of keys per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(backbone="ResNet18", nf=4, B=256, K=65536, D=128, T=0.07, m=0.999, H=224)
METRIC = "frames/sec (224^2 multi-view) through encoder+InfoNCE+EMA/enqueue step"


def conv_flops_per_frame(backbone):
    # SURVEY.md 8d (probed with hooks on the reference): 2*MACs of the convs through layer4
    return {"ResNet18": 3.627e9, "ResNet50": 8.174e9}[backbone]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_conv_traffic():
    """Average DRAM bytes (read + write) per conv_gemm launch from the committed ncu capture of this same command
    (profiles/r01_step_metrics_all_launches.csv, dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01_step_metrics_all_launches.csv")
    try:
        rows = list(csv.reader(open(path, newline="")))
        hdr = rows[0]
        k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[rows[1][r]]
        vals = [(float(x[r].replace(",", "")) + float(x[w].replace(",", ""))) * unit for x in rows[2:] if "conv_gemm" in x[k]]
        return (round(sum(vals) / len(vals)), len(vals)) if vals else None
    except Exception:
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


# ----------------------------------------------------------------------------------------------------------
def make_args(dev, wl):
    import vince_b200
    return types.SimpleNamespace(
        backbone=getattr(vince_b200, wl["backbone"]), num_frames=wl["nf"], use_attention=False,
        feature_extractor_gpu_ids=[dev], pytorch_gpu_ids=[dev], vince_embedding_size=wl["D"],
        vince_queue_size=wl["K"], vince_temperature=wl["T"], vince_self_temperature=0.03, vince_momentum=wl["m"],
        jigsaw=False, inter_batch_comparison=True, self_batch_comparison=False, batch_size=wl["B"], use_imagenet=False)


class HotPath:
    """The step, written against the reference-facing API exactly as VinceSolver.run_train_iteration calls it."""

    def __init__(self, dev, wl, rank, world, gather):
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
        if wl.get("input", "fp32") == "uint8":
            shape = (wl["B"], wl["H"], wl["H"], 3)
            self.host_data = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).pin_memory()
            self.host_queue_data = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).pin_memory()
        else:
            shape = (wl["B"], 3, wl["H"], wl["H"])
            self.host_data = torch.randn(shape, generator=g).pin_memory()
            self.host_queue_data = torch.randn(shape, generator=g).pin_memory()
        self.dev_data = self.host_data.to(dev)
        self.dev_queue_data = self.host_queue_data.to(dev)
        self.launches = 0
        self.prefetch = None

    def step(self, data, queue_data):
        wl = self.wl
        batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [wl["B"]],
                 "data_source": "synthetic", "num_frames": wl["nf"]}
        queue_batches = self.qm(batch, shuffle=True)                                  # vince_solver.py:405
        launches = self.qm.launches
        outputs = self.model.get_embeddings(batch, shuffle=True)                      # :406
        launches += self.model.launches
        output = outputs[0]
        output.update(self.queue.dequeue())                                           # :420
        output.update({"data_source": "synthetic", "num_frames": wl["nf"]})
        output.update(queue_batches[0])
        self.model.launches = 0
        output.update(self.model(output))                                             # :424
        loss = self.model.loss(output)["nce_loss"][1]                                 # :425
        self.model.get_metrics(output)                                                # :426
        launches += self.model.launches
        keys = output["queue_embeddings"]
        if self.gather is not None:                                                   # :497 with the key all-gather
            self.gather.enqueue(self.queue, keys, None, "synthetic")
            self.qm.vince_update(self.model)
            launches += 2
        else:
            self.qm.vince_update(self.model, enqueue=(self.queue, keys, [None] * wl["B"], "synthetic"))   # :497-499
        launches += 1
        self.launches = launches
        return loss

    def run_e2e(self, steps):
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
        host_batch = {"data": self.host_data, "queue_data": self.host_queue_data}
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
    wl = dict(WORKLOAD)
    if a.backbone == "ResNet50":
        wl.update(backbone="ResNet50", T=0.2)
    wl["input"] = a.input
    hp = HotPath(dev, wl, rank, world, gather)

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

    for _ in range(max(a.warmup, 3)):
        hp.step(hp.dev_data, hp.dev_queue_data)
    # ---- device-resident leg (value) with per-launch events on the tensor-core kernel ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.PROFILE = []
    torch.cuda.profiler.start()      # ncu --profile-from-start off captures exactly the timed steps (no-op otherwise)
    ms = timed(lambda: hp.step(hp.dev_data, hp.dev_queue_data), a.steps)
    torch.cuda.profiler.stop()
    prof, ops.PROFILE = ops.PROFILE, None
    clocks = sampler.stop() if rank == 0 else None
    frames_per_step = 2 * wl["B"] * world
    value = frames_per_step * a.steps / (ms / 1e3)
    conv_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in prof)
    conv_flops = sum(f for _, f, _, _ in prof)
    n_conv = len(prof)
    # ---- same kernel timed with the two encoders serialised (no co-running streaming kernels of the other encoder) ----
    os.environ["VINCE_B200_OVERLAP"] = "0"
    iso_steps = max(2, min(a.steps, 10))
    hp.step(hp.dev_data, hp.dev_queue_data)
    ops.PROFILE = []
    iso_ms = timed(lambda: hp.step(hp.dev_data, hp.dev_queue_data), iso_steps)
    prof_iso, ops.PROFILE = ops.PROFILE, None
    os.environ["VINCE_B200_OVERLAP"] = "1"
    iso_conv_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in prof_iso)
    iso_conv_flops = sum(f for _, f, _, _ in prof_iso)
    if a.profile_only:               # under ncu: the launch list / --set full capture of the timed steps is all we want
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            emit({"profile_only": True, "ms_per_step_under_profiler": round(ms / a.steps, 3)})
        return
    # ---- InfoNCE step (similarity+CE+metrics + EMA + enqueue) timed alone, device resident ----
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
        hp.qm.vince_update(hp.model, enqueue=(hp.queue, keys, [None] * wl["B"], "synthetic"))
    for _ in range(3):
        nce_step()
        nce_step(True)
    nce_ms = timed(nce_step, 20) / 20
    nce_bwd_ms = timed(lambda: nce_step(True), 20) / 20
    # ---- end-to-end leg: host (pinned) inputs, H2D inside the timed region, loss read back every step ----
    hp.run_e2e(3)
    e2e_steps = a.steps
    e2e_losses = []
    e2e_ms = timed(lambda: e2e_losses.extend(hp.run_e2e(e2e_steps)), 1)
    e2e_value = frames_per_step * e2e_steps / (e2e_ms / 1e3)
    assert len(e2e_losses) == e2e_steps and all(l == l for l in e2e_losses)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    traffic = ncu_conv_traffic() if wl["backbone"] == "ResNet18" else None
    achieved_tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak_tf = peaks["tf_sustained"]
    iso_tf = iso_conv_flops / (iso_conv_ms / 1e3) / 1e12 if iso_conv_ms > 0 else 0.0
    roofline = {
        "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, fp16x3)", "bound": "tensor",
        # the kernel's own duration: CUDA events around every conv_gemm launch, on its launching stream, with the key /
        # query encoders serialised on one stream (VINCE_B200_OVERLAP=0) in %d steps run right after the timed region
        "achieved": round(iso_tf, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(iso_tf / peak_tf, 4),
        "traffic": traffic[0] if traffic else None,
        "traffic_source": ("bytes per launch: mean dram__bytes_read.sum + dram__bytes_write.sum over the %d conv_gemm "
                           "launches of one step of the ResNet-18 workload in profiles/r01_step_metrics_all_launches.csv "
                           "(ncu capture of this command)" % traffic[1]) if traffic else None,
        "peak_source": "%s bf16 cuBLAS sustained (same tensor rate as fp16; kernel timed inside a long step)" % peaks["source"],
        "launches_timed": len(prof_iso), "avg_launch_us": round(iso_conv_ms * 1e3 / max(len(prof_iso), 1), 2),
        "algorithmic_gflop_per_launch": round(iso_conv_flops / max(len(prof_iso), 1) / 1e9, 3),
        "share_of_step": round(iso_conv_ms / iso_ms, 4), "ms_per_step": round(iso_ms / iso_steps, 4),
        "how": "per-launch CUDA events on the launching stream over %d steps with the two encoders serialised on one "
               "stream (VINCE_B200_OVERLAP=0), run right after the timed region: this is the kernel's own duration, "
               "and its share of that step is what the ncu launch list in profiles/ must agree with" % iso_steps,
        "timed_region": {"achieved": round(achieved_tf, 2), "frac": round(achieved_tf / peak_tf, 4),
                         "launches_timed": n_conv, "avg_launch_us": round(conv_ms * 1e3 / max(n_conv, 1), 2),
                         "share_of_step": round(conv_ms / ms, 4),
                         "what": "the same per-launch events inside the timed region itself, where the two encoders run "
                                 "on two streams: a conv launch there shares the SMs and HBM with the other encoder's "
                                 "kernels, so its event duration (and share_of_step, which exceeds 1) includes that "
                                 "co-running work"},
        "note": "algorithmic FLOPs = 2*M*N*K of the fp32 conv; the kernel issues 3 fp16 MMAs per k-step to reach "
                "fp32-grade accuracy, so frac <= 1/3 by construction",
    }
    # the CPU leg is a property of the box, not of N: measured on rank 0 of the N=1 run only
    cpu = cpu_baseline(wl, seconds=15.0) if world == 1 else {
        "value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
        "sample": "not measured at N>1 (see the N=1 line of the same run, or --impl reference)"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (fp16x3 split MMA, fp32 accumulate; TF32 for InfoNCE negatives)",
        "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[%d]: %s, 4 views/clip, batch=256 frames/GPU, queue K=65536, "
                               "dim=128, 224x224" % (1 if wl["backbone"] == "ResNet18" else 2, wl["backbone"]), "per_gpu_batch": wl["B"], "frames_per_step": frames_per_step,
                   "parallelism": "dp%d (replicated weights+queue, NCCL all-gather of keys)" % world if world > 1 else "single GPU",
                   "input": "fp32 NCHW normalised frames (the reference's wire format)" if a.input == "fp32" else
                            "uint8 HWC raw frames, ToTensor+Normalize fused into the stem packing",
                   "streams": "key encoder on the caller's stream, query encoder on a side stream (joined before "
                              "get_embeddings returns); VINCE_B200_OVERLAP=0 serialises them",
                   "l2_policy": "inputs larger than L2: 2 x 154 MB of fp32 frames + >1 GB of activations per step"},
        "infonce_step_ms": round(nce_ms, 4),
        "infonce_step_with_dq_backward_ms": round(nce_bwd_ms, 4),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "frames/s", "ms_per_step": round(e2e_ms / e2e_steps, 4),
                "h2d_bytes_per_step": hp.h2d_bytes_per_step, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "pipeline": "H2D of step i+1 on a copy stream under step i (vince_b200.prefetch.BatchPrefetcher, "
                            "mirrors the solver's prefetch thread); loss of step i read by the host during step i+1"},
        "gpu_launches": hp.launches * a.steps,
        "gpu_launches_per_step": hp.launches,
    }
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port (reference algorithm on torch CPU) - the only places bench.py touches oracle/
# ----------------------------------------------------------------------------------------------------------
def cpu_step_runner(wl, B):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vince_oracle as vo
    torch.set_num_threads(os.cpu_count())
    sd_q = vo.make_state_dict(wl["backbone"], wl["D"], seed=0)
    sd_k = vo.clone_state_dict(sd_q)
    g = torch.Generator().manual_seed(1234)
    data = torch.randn((B, 3, wl["H"], wl["H"]), generator=g)
    queue_data = torch.randn((B, 3, wl["H"], wl["H"]), generator=g)
    queue = vo.StorageQueue(wl["K"], wl["D"])
    names = vo.vince_parameter_names(sd_q)

    def step():
        with torch.no_grad():
            perm_k, perm_q = torch.randperm(B), torch.randperm(B)
            out = vo.train_step(data, queue_data, sd_q, sd_k, queue, wl["backbone"], wl["nf"], wl["T"], wl["m"],
                                shuffle_q=perm_q, shuffle_k=perm_k, ema_names=names)
        return float(out["losses"]["nce_loss"])
    return step


def cpu_baseline(wl, seconds):
    B = 32
    step = cpu_step_runner(wl, B)
    step()
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > seconds or n >= 50:
            break
    dt = time.perf_counter() - t0
    return {"value": round(2 * B * n / dt, 2), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle port (torch CPU fp32, %d threads) of the same step at batch=%d frames (8 clips x 4 views), "
                      "K=%d, D=%d, 224x224; %d steps in %.1f s after 1 warm-up" % (os.cpu_count(), B, wl["K"], wl["D"], n, dt)}


def run_reference(a):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Python reference itself cannot travel to
    the GPU box) on this box's host cores, same metric/config, each step a bounded sample (batch 32 frames)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = dict(WORKLOAD)
    if a.backbone == "ResNet50":
        wl.update(backbone="ResNet50", T=0.2)
    B = 32
    step = cpu_step_runner(wl, B)
    for _ in range(max(1, min(a.warmup, 2))):
        step()
    steps = max(1, min(a.steps, 12))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = round(2 * B * steps / dt, 2)
    sample = ("oracle port of the reference step (torch CPU fp32, %d threads), batch=%d frames per step "
              "(bounded sample of the batch=256 workload), K=%d, D=%d, 224x224" % (os.cpu_count(), B, wl["K"], wl["D"]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": steps, "warmup": max(1, min(a.warmup, 2)), "ms_per_step": round(dt / steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[%d]: %s, 4 views/clip, queue K=65536, dim=128, 224x224; "
                                   "CPU sample batch=32 frames" % (1 if wl["backbone"] == "ResNet18" else 2, wl["backbone"])},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backbone", default="ResNet18", choices=["ResNet18", "ResNet50"],
                    help="ResNet18 = BASELINE.json configs[1] (the bench contract's workload, default); ResNet50 = configs[2] "
                         "(MoCoV2 config, T=0.2), an extra line for the record")
    ap.add_argument("--input", default="fp32", choices=["fp32", "uint8"],
                    help="fp32 = the reference's wire format (normalised NCHW frames, default); uint8 = raw HWC frames with "
                         "the normalisation fused into the stem packing (SURVEY.md 8f rank 3), an extra line for the record")
    ap.add_argument("--profile-only", action="store_true",
                    help="stop after the device-resident timed steps (for runs under ncu; prints no bench value)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
