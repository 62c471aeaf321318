#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): frames/sec of one PMF-ResNet34 training step (forward + loss + backward +
optimizer step) on the SemanticKITTI-shaped synthetic workload — batch 8 per GPU on the 480x640 camera grid
(a frame = one 64x2048 = 131 072-point sweep perspective-projected onto a 480x640 RGB image).

    python bench.py --gpus 1 --steps K --warmup W                      our arm (libpmf_b200.so through pmf_b200.PMFNet)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   frame-parallel DDP, weak scaling
    python bench.py --impl reference ...                                the reference's CPU PyTorch path (oracle port)

Prints ONE JSON line on rank 0.  `value` = frames/s with inputs resident in HBM; `e2e` = the same step driven from
pinned HOST buffers through the public module call (H2D of the (B,8,H,W) frame tensor + labels and D2H of the loss
inside the timed region).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  The per-step working set (tens of GB of activations) is far larger than the 126 MB L2, so no explicit
L2 flush is needed between iterations (stated in config.l2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FWD_KFLOP_PER_PX = 1806.6   # PMF-ResNet34 forward, 2*MAC, convs only
STEP_KFLOP_PER_PX = 5400.7  # forward + dgrad + wgrad (minus the two input dgrads)
METRIC = "frames/sec PMF-ResNet34 fwd+bwd (480x640 camera grid, batch 8/GPU)"
HALO_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "halo_traffic.json")  # ncu dram bytes of the dominant kernel (see kernel_roofline)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu-eager"])
    ap.add_argument("--loss", default="fused", choices=["fused", "torch"],
                    help="our arm: pmf_b200.loss.TrainerLoss implementation (fused = libpmf_b200.so; torch = the block as the "
                         "unchanged trainer.py runs it)")
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--backbone", default="resnet34", help="camera encoder (BASELINE config 4: resnet50)")
    ap.add_argument("--nclasses", type=int, default=20, help="20 SemanticKITTI / 17 nuScenes")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--cpu-sample-frames", type=int, default=1)
    ap.add_argument("--ddp", default="flat", choices=["flat", "torch"],
                    help="N > 1: pmf_b200.dist.FrameParallel (per-segment all-reduce of the flat gradient buffer) or stock torch DDP")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--kernel-timing", type=int, default=1, help="extra instrumented step for the roofline object")
    ap.add_argument("--extras", type=int, default=1, help="also time the eval forward + KNN tail (rank 0)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ shared pieces
def make_frames(B, H, W, seed):
    """(B,8,H,W) normalised frame tensor + (B,H,W) labels, as tasks/pmf/trainer.py:291-299 hands them to the model."""
    from tests import synth
    feat, _mask, label = synth.frame_tensor(B, H, W, seed=seed, density=0.1)
    if NCLASSES < 20:
        label = label % NCLASSES  # nuScenes-shaped: 16 classes + ignored 0
    return feat, label


NCLASSES, LAMBDA, GAMMA, TAU = 20, 1.0, 0.5, 0.7  # tasks/pmf/config_server_kitti.yaml:29-31
BACKBONE = "resnet34"


def load_reference():
    """The UNMODIFIED reference sources staged under baseline/_ref (tools/stage_reference.py; git-ignored, travels with the
    gpurun snapshot), imported under an alias with a stub parent package (its own __init__ pulls tensorboardX and the
    nuScenes devkit, which are not installed).  Returns a namespace with .models and .loss, or None when nothing is
    staged (the comparator arms then fall back to the oracle port)."""
    import importlib
    import types
    base = os.path.join(ROOT, "baseline", "_ref", "pc_processor")
    if not os.path.isdir(os.path.join(base, "models")):
        return None
    alias = "ref_pc_processor_bench"
    if alias not in sys.modules:
        pkg = types.ModuleType(alias)
        pkg.__path__ = [base]
        pkg.__package__ = alias
        sys.modules[alias] = pkg
    pkg = sys.modules[alias]
    for sub in ("models", "loss"):
        setattr(pkg, sub, importlib.import_module(alias + "." + sub))
    return pkg


class ReferenceLossBlock:
    """tasks/pmf/trainer.py:188-252, 305-332 composed from the reference's OWN loss classes (comparator arms only)."""

    def __init__(self, ref, device):
        import numpy as np
        alpha = np.ones(NCLASSES)
        alpha[0] = 0  # trainer.py:200-201
        self.lovasz = ref.loss.Lovasz_softmax(ignore=0).to(device)
        self.kl = torch.nn.KLDivLoss(reduction="none").to(device)
        self.focal = ref.loss.FocalSoftmaxLoss(NCLASSES, gamma=2, alpha=alpha, softmax=False).to(device)

    def __call__(self, lidar_pred, camera_pred, label):
        import math
        label_mask = label.gt(0)
        lidar_log = torch.log(lidar_pred.clamp(min=1e-8))
        pcd_entropy = -(lidar_pred * lidar_log).sum(1) / math.log(NCLASSES)
        loss_foc, loss_lov = self.focal(lidar_pred, label, mask=label_mask), self.lovasz(lidar_pred, label)
        camera_log = torch.log(camera_pred.clamp(min=1e-8))
        img_entropy = -(camera_pred * camera_log).sum(1) / math.log(NCLASSES)
        loss_foc_cam, loss_lov_cam = self.focal(camera_pred, label, mask=label_mask), self.lovasz(camera_pred, label)
        pcd_conf, img_conf = 1 - pcd_entropy, 1 - img_entropy
        imp = pcd_conf - img_conf
        pcd_w = imp.gt(0).float() * imp.abs() * pcd_conf.ge(TAU).float()
        img_w = imp.lt(0).float() * imp.abs() * img_conf.ge(TAU).float()
        loss_per = (self.kl(lidar_log, camera_pred) * img_w.unsqueeze(1)).mean() + \
                   (self.kl(camera_log, lidar_pred) * pcd_w.unsqueeze(1)).mean()
        return loss_foc + loss_lov * LAMBDA + loss_foc_cam + loss_lov_cam * LAMBDA + loss_per * GAMMA


def oracle_loss_block(lidar_pred, camera_pred, label):
    """Fallback of the comparator arms when baseline/_ref is not staged: the oracle restatement of the same block."""
    from oracle import loss_oracle as lo
    alpha = torch.ones(NCLASSES, device=lidar_pred.device)
    alpha[0] = 0
    return lo.total_loss(lidar_pred, camera_pred, label, alpha, NCLASSES, LAMBDA, GAMMA, TAU)


def make_optimizers(lidar_params, camera_params):
    """tasks/pmf/trainer.py:80-98: AdamW(lr 1e-3) on the LiDAR stream, nesterov SGD on the camera encoder+decoder."""
    return (torch.optim.AdamW(lidar_params, lr=1e-3), torch.optim.SGD(camera_params, lr=1e-3, momentum=0.9, nesterov=True,
                                                                        weight_decay=1e-5))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1354.7), d.get("hbm_gbs", 6546.2), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ comparator arms
class _RefStepper:
    """One training step of the reference path: forward + the trainer's loss block + backward + both optimisers
    (tasks/pmf/trainer.py:289-341), on `device`.  Model and loss are the reference's OWN classes when baseline/_ref is
    staged (kind "reference"), else the oracle port (kind "port")."""

    def __init__(self, device, channels_last=False, autocast_bf16=False):
        self.dev = torch.device(device)
        self.ref = load_reference()
        self.autocast = autocast_bf16
        torch.manual_seed(1)
        if self.ref is not None:
            import warnings
            self.kind = "reference"
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.model = self.ref.models.PMFNet(pcd_channels=5, img_channels=3, nclasses=NCLASSES, base_channels=32,
                                                    image_backbone=BACKBONE, imagenet_pretrained=False).to(self.dev)
            if channels_last:
                self.model = self.model.to(memory_format=torch.channels_last)
            self.model.train()
            self.loss = ReferenceLossBlock(self.ref, self.dev)
            lidar = list(self.model.lidar_stream.parameters())
            camera = list(self.model.camera_stream_encoder.parameters()) + list(self.model.camera_stream_decoder.parameters())
        else:
            from oracle import pmf_oracle as po
            self.kind = "port"
            self.po = po
            sd = po.synth_state_dict(po.pmf_param_shapes(NCLASSES, 32, BACKBONE), seed=1)
            self.params = {k: (v.clone().to(self.dev).requires_grad_(True) if v.dtype.is_floating_point and "running" not in k
                               else v.clone().to(self.dev)) for k, v in sd.items()}
            self.loss = oracle_loss_block
            lidar = [v for k, v in self.params.items() if k.startswith("lidar_stream") and v.requires_grad]
            camera = [v for k, v in self.params.items() if not k.startswith("lidar_stream") and v.requires_grad]
        self.opt_a, self.opt_b = make_optimizers(lidar, camera)
        self.channels_last = channels_last

    def step(self, feat, label):
        pcd, img = feat[:, 0:5], feat[:, 5:8]  # channel-slice views like trainer.py:296-297
        if self.channels_last:
            pcd, img = pcd.contiguous(memory_format=torch.channels_last), img.contiguous(memory_format=torch.channels_last)
        with torch.autocast(self.dev.type, dtype=torch.bfloat16, enabled=self.autocast):
            if self.kind == "reference":
                lid, cam = self.model(pcd, img)
            else:
                lid, cam, ctx = self.po.pmf_forward(self.params, pcd, img, BACKBONE, train=True, return_ctx=True)
        loss = self.loss(lid.float(), cam.float(), label)
        self.opt_a.zero_grad(set_to_none=True)
        self.opt_b.zero_grad(set_to_none=True)
        loss.backward()
        self.opt_a.step()
        self.opt_b.step()
        if self.kind == "port":
            for k, v in ctx.new_stats.items():
                self.params[k] = v
        return loss.detach()


def cpu_reference_step_rate(B, H, W, steps, warmup, threads=None):
    """The reference's CPU PyTorch path for the same step (fp32, batch-stat BN, the trainer's loss block and optimisers)
    on the host cores.  Returns (frames/s, seconds per step, kind)."""
    if threads:
        torch.set_num_threads(threads)
    st = _RefStepper("cpu")
    feat, label = make_frames(B, H, W, seed=1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        float(st.step(feat, label))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return B / sec, sec, st.kind


def workload_config(B, H, W, world):
    px = H * W
    return {"workload": "PMF-%s %s-shaped synthetic, batch %d/GPU, %dx%d camera grid, fwd + trainer loss block "
                        "(focal + Lovasz on both heads + perception-aware KL) + bwd + AdamW/SGD step, train-mode BN + Dropout2d"
                        % (BACKBONE.replace("resnet", "ResNet"), "SemanticKITTI" if NCLASSES == 20 else "nuScenes", B, H, W),
            "backbone": BACKBONE, "nclasses": NCLASSES,
            "frames_per_gpu": B, "height": H, "width": W, "parallelism": "dp%d (frame-parallel, gradient all-reduce over NCCL)" % world,
            "l2": "no explicit flush: per-step working set (>20 GB of activations) >> 126 MB L2",
            "step_gflop_per_frame": STEP_KFLOP_PER_PX * px / 1e6}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the host cores (all threads), each step a
    bounded sample of the workload (cpu_sample_frames frames of the same shape); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = max(1, args.cpu_sample_frames)
    fps, sec, kind = cpu_reference_step_rate(B, args.height, args.width, args.steps, max(args.warmup, 1), threads=cores)
    sample = "%d frame(s) per step at %dx%d (of the batch-%d workload), fwd + trainer loss block + bwd + optimizer step, %d warm-up + %d timed steps" % (
        B, args.height, args.width, args.batch, max(args.warmup, 1), args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.batch, args.height, args.width, max(args.gpus, 1)),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu_eager(args):
    """--impl gpu-eager (run as a subprocess of our arm): the reference's eager PyTorch path on ONE B200 — the comparator
    the north-star 5x target is stated against: cudnn.benchmark=True (tasks/pmf/main.py:23), same step
    (forward + trainer loss block + backward + AdamW/SGD), same batch and shape, CUDA-event timed after warm-up.
    Variants: torch defaults (TF32 convolutions), true fp32 (allow_tf32=False), channels_last + bf16 autocast."""
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    B, H, W = args.batch, args.height, args.width
    feat, label = make_frames(B, H, W, seed=1000)
    feat, label = feat.to(dev), label.to(dev)
    out = {"batch": B, "height": H, "width": W, "steps": args.steps, "warmup": args.warmup}
    for name, tf32, cl, bf16 in (("tf32_default", True, False, False), ("fp32_no_tf32", False, False, False),
                                 ("channels_last_bf16_autocast", True, True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            st = _RefStepper(dev, channels_last=cl, autocast_bf16=bf16)
            for _ in range(max(args.warmup, 2)):
                st.step(feat, label)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = st.step(feat, label)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out[name] = {"frames_per_s": B / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss), "kind": st.kind,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except Exception as exc:  # noqa: BLE001  (an OOM of one variant must not lose the others)
            out[name] = {"error": repr(exc)[:300]}
        st = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    print(json.dumps({"impl": "gpu-eager", "gpu_eager_baseline": out}), flush=True)


def gpu_eager_subprocess(B, H, W, steps=4, warmup=2, timeout=600):
    """Runs run_gpu_eager in a fresh process (its ~100 GB of eager activations must not share the allocator with our CUDA
    graphs) and returns the parsed object; on failure {"error": ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "gpu-eager", "--batch", str(B), "--height", str(H), "--width", str(W),
           "--steps", str(steps), "--warmup", str(warmup), "--backbone", BACKBONE, "--nclasses", str(NCLASSES)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["gpu_eager_baseline"]
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as exc:  # noqa: BLE001
        return {"error": repr(exc)[:300]}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    import pmf_b200
    from pmf_b200 import _lib as L
    from pmf_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200; there is no CPU fallback"
    if world > 1:  # a rank that stops making progress must fail loudly with its stack, never hang the node
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("PMFB_BENCH_WATCHDOG", "420")), exit=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on STDOUT while the communicator is created; the contract is ONE JSON line on
        # stdout, so file descriptor 1 points at stderr until the communicator exists.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B, H, W = args.batch, args.height, args.width

    torch.manual_seed(1)
    model = pmf_b200.PMFNet(5, 3, NCLASSES, 32, False, BACKBONE).to(dev)
    model.train()
    net = model
    if world > 1:
        if args.ddp == "torch":
            # tasks/pmf/trainer.py:38-39 (what the unchanged trainer builds), plus gradient_as_bucket_view
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
        else:
            # pmf_b200.dist.FrameParallel: same semantics (rank-0 broadcast of parameters / BN buffers, averaged gradients),
            # but the all-reduce runs per backward segment on the module's flat gradient buffer and overlaps the backward
            net = pdist.wrap_ddp(model, local)
    opt_a, opt_b = make_optimizers(list(model.lidar_stream.parameters()),
                                   list(model.camera_stream_encoder.parameters()) + list(model.camera_stream_decoder.parameters()))
    # frame-parallel sharding: every rank owns its own B frames (weak scaling), no data-path collective
    feat, label = make_frames(B, H, W, seed=pdist.shard_seed(1, rank))
    h_feat, h_label = feat.pin_memory(), label.pin_memory()
    d_feat, d_label = h_feat.to(dev, non_blocking=True), h_label.to(dev, non_blocking=True)
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()

    from pmf_b200.loss import TrainerLoss
    crit = {"fused": TrainerLoss(NCLASSES, None, LAMBDA, GAMMA, TAU, impl="auto").to(dev),
            "torch": TrainerLoss(NCLASSES, None, LAMBDA, GAMMA, TAU, impl="torch").to(dev)}
    loss_sel = {"impl": args.loss}

    def step(x, y):
        lid, cam = net(x[:, 0:5], x[:, 5:8])  # channel-slice views like trainer.py:296-297
        loss = crit[loss_sel["impl"]](lid, cam, y)  # the trainer's loss block (trainer.py:305-332)
        opt_a.zero_grad(set_to_none=True)
        opt_b.zero_grad(set_to_none=True)
        loss.backward()
        opt_a.step()
        opt_b.step()
        return loss

    # e2e input path: every step's frames + labels are copied from pinned host memory into one of two device buffers on
    # a copy stream, one step ahead of the compute stream (what a DataLoader with pin_memory + non_blocking gives the
    # trainer); the H2D of step i+1 overlaps the kernels of step i.  Every step still pays its own H2D and D2H.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(d_feat), torch.empty_like(d_label)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]   # H2D into buffer k finished
    consumed = [torch.cuda.Event(), torch.cuda.Event()]  # the step reading buffer k finished
    state = {"k": 0, "primed": False}

    def prefetch(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            bufs[k][0].copy_(h_feat, non_blocking=True)
            bufs[k][1].copy_(h_label, non_blocking=True)
            ready[k].record(copy_stream)

    def step_e2e():
        cur = torch.cuda.current_stream()
        if not state["primed"]:
            for k in range(2):
                consumed[k].record(cur)
            prefetch(0)
            state["primed"] = True
        k = state["k"]
        prefetch(1 - k)                 # next step's inputs, overlapping this step's compute
        cur.wait_event(ready[k])
        loss = step(bufs[k][0], bufs[k][1])
        consumed[k].record(cur)
        state["k"] = 1 - k
        h_loss.copy_(loss.detach(), non_blocking=True)
        cur.synchronize()  # the user reads the loss every iteration (trainer.py:384-407)
        return float(h_loss)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t)
        return ms

    for _ in range(max(args.warmup, 3)):
        step(d_feat, d_label)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.launches
    ms = timed(lambda: step(d_feat, d_label), args.steps)
    launches = L.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # the same step with the loss block as the UNCHANGED trainer.py runs it (plain PyTorch ops on our outputs)
    ms_torch_loss = None
    if args.loss != "torch":
        loss_sel["impl"] = "torch"
        step(d_feat, d_label)
        ms_torch_loss = timed(lambda: step(d_feat, d_label), max(3, args.steps // 4)) / max(3, args.steps // 4)
        loss_sel["impl"] = args.loss

    # exposed (non-overlapped) gradient all-reduce: the same step with DDP's synchronisation switched off
    ddp_info = None
    if world > 1:
        def step_nosync():
            with net.no_sync():
                step(d_feat, d_label)
        step_nosync()
        k_ns = max(3, args.steps // 2)
        ms_ns = timed(step_nosync, k_ns) / k_ns
        ddp_info = {"ms_per_step_no_sync": ms_ns, "exposed_allreduce_ms": ms / args.steps - ms_ns,
                    "grad_bytes": sum(p.numel() for p in model.parameters()) * 4,
                    "backward_segments": int(os.environ.get("PMFB_BWD_SEGMENTS", "4")), "wrapper": args.ddp}

    frames = B * world * args.steps
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: the tcgen05 implicit-GEMM conv (forward + dgrad launches), timed live with
    # CUDA events around every launch of one extra step on the launching stream.  EVERY rank runs the instrumented
    # steps (they go through DDP's gradient all-reduce); only rank 0 reports.
    peak_tf, peak_bw, peak_src = peaks()
    roof = None
    if args.kernel_timing:
        roof = kernel_roofline(lambda: step(d_feat, d_label), dev, ms / args.steps, peak_tf, peak_src)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    extras = None
    if args.extras and world == 1:
        extras = inference_extras(model, d_feat, dev, B, H, W)
        extras["postproc_rooflines"] = postproc_rooflines(dev, peak_bw)
    cpu = eager = grid2 = None
    if world == 1:
        if args.extras:
            # secondary shape S2: the 64x2048 common grid of the metric string, same step, same batch
            del d_feat, d_label, bufs
            model._graphs.clear()
            torch.cuda.empty_cache()
            g_feat, g_label = make_frames(B, 64, 2048, seed=2)
            g_feat, g_label = g_feat.to(dev), g_label.to(dev)
            for _ in range(3):
                step(g_feat, g_label)
            k2 = max(5, args.steps // 2)
            ms2 = timed(lambda: step(g_feat, g_label), k2) / k2
            grid2 = {"workload": "same step on the 64x2048 common grid, batch %d" % B, "frames_per_s": B / (ms2 * 1e-3),
                     "ms_per_step": ms2, "step_tflops": STEP_KFLOP_PER_PX * 1e3 * 64 * 2048 * B / (ms2 * 1e-3) / 1e12}
            del g_feat, g_label
        model._graphs.clear()
        del net, model, opt_a, opt_b
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        if not args.no_gpu_eager:
            eager = gpu_eager_subprocess(B, H, W)
            base = (eager.get("tf32_default") or {}).get("frames_per_s")
            if base:
                eager["ours_over_tf32_default"] = value / base
                eager["ours_e2e_over_tf32_default"] = e2e / base
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fps, sec, kind = cpu_reference_step_rate(args.cpu_sample_frames, H, W, steps=5, warmup=1, threads=cores)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                   "sample": "%d frame(s)/step at %dx%d (the reference's PMFNet + loss classes, fp32, all host threads), fwd + trainer "
                             "loss block + bwd + optimizer step, 1 warm-up + 5 timed steps" % (args.cpu_sample_frames, H, W)}
    cfg = workload_config(B, H, W, world)
    from pmf_b200 import _lib as _PL
    cfg["precision"] = {"f16": "kind::f16 UMMAs on 16-bit operand shadows (fp16 forward, bf16 dgrad/wgrad) for the >=64-channel "
                               "stride-1 convolutions, kind::tf32 elsewhere; fp32 accumulate / storage / BatchNorm",
                        "tf32": "kind::tf32 operands, fp32 accumulate/storage",
                        "3xtf32": "3xTF32 hi/lo split operands (precise mode)"}[_PL.get_precision()]
    cfg["precision_mode"] = _PL.get_precision()
    cfg["loss_impl"] = args.loss
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"f16": "f16/bf16+tf32 operands, f32 accumulate", "tf32": "tf32", "3xtf32": "3xtf32"}[_PL.get_precision()],
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h_feat.numel() * 4 + h_label.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "step_tflops": STEP_KFLOP_PER_PX * 1e3 * H * W * B / (ms / args.steps * 1e-3) / 1e12,
            "ms_per_step_with_torch_loss_block": ms_torch_loss, "ddp": ddp_info,
            "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "common_grid_64x2048": grid2, "extras": extras}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _timed_launches(fn, reps, flush_l2=True):
    """Average device time of `fn` over `reps` launches, each bracketed by its own CUDA events; between launches a 256 MB
    buffer is overwritten so that no launch finds its inputs in the 126 MB L2 (timing rule for small working sets)."""
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda") if flush_l2 else None
    for _ in range(3):
        fn()
    evs = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps


def postproc_rooflines(dev, peak_bw):
    """HBM rooflines of the two scatter/gather kernels: KNN back-projection, algorithmic bytes
    12*H*W + 28*P per frame (range f32 + argmax i64 images once; px, py i64 + range f32 in, label i64 out per point), and
    the perspective projection, 16*N in + 40*H*W out.  BASELINE frame: 480x640 grid, one 64x2048 = 131072-point sweep."""
    import contextlib
    import io

    import pmf_b200
    from pmf_b200 import postproc
    from tests import synth
    H, W = 480, 640
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        knn = pmf_b200.KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), 20)
    for P in (32768, 131072):
        case = dict(name="bench", H=H, W=W, P=P, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.9, seed=5, kind="rand")
        inp = synth.knn_inputs(case)
        pr, ur, am, px, py = (torch.from_numpy(inp[k]).to(dev) for k in ("proj_range", "unproj_range", "proj_argmax", "px", "py"))
        for frames in (1, 8, 32):
            if hasattr(postproc, "knn_batched"):
                prb, amb = pr.unsqueeze(0).expand(frames, H, W).contiguous(), am.unsqueeze(0).expand(frames, H, W).contiguous()
                urb, pxb, pyb = ur.repeat(frames), px.repeat(frames), py.repeat(frames)
                offs = torch.arange(frames + 1, device=dev, dtype=torch.int64) * P
                fn = lambda: postproc.knn_batched(knn, prb, urb, amb, pxb, pyb, offs)  # noqa: E731
            elif frames == 1:
                fn = lambda: knn(pr, ur, am, px, py)  # noqa: E731
            else:
                continue
            ms = _timed_launches(fn, 10)
            nbytes = frames * (12.0 * H * W + 28.0 * P)
            gbs = nbytes / (ms * 1e-3) / 1e9
            out["knn_P%d_frames%d" % (P, frames)] = {
                "bound": "hbm", "kernel": "knn_vote_kernel", "achieved": gbs, "peak": peak_bw, "unit": "GB/s", "frac": gbs / peak_bw,
                "traffic": None, "alg_bytes_per_launch": nbytes, "ms_per_launch": ms, "points_per_s": frames * P / (ms * 1e-3),
                "frames_per_launch": frames, "l2": "flushed between launches"}
    pts, lab = synth.lidar_sweep(64, 2048, seed=1)
    dp, dl = torch.from_numpy(pts).to(dev), torch.from_numpy(lab).to(torch.int32).to(dev)
    M = synth.camera_matrix(H, W)
    ms = _timed_launches(lambda: pmf_b200.project_scatter(dp, dl, M, H, W), 10)
    nbytes = 16.0 * pts.shape[0] + 40.0 * H * W
    gbs = nbytes / (ms * 1e-3) / 1e9
    out["project_scatter_N131072"] = {
        "bound": "hbm", "kernel": "project_points_kernel + project_gather_kernel", "achieved": gbs, "peak": peak_bw, "unit": "GB/s",
        "frac": gbs / peak_bw, "traffic": None, "alg_bytes_per_launch": nbytes, "ms_per_launch": ms,
        "frames_per_s": 1.0 / (ms * 1e-3), "l2": "flushed between launches",
        "note": "ms includes the 7 torch.empty output allocations of the host wrapper"}
    return out


def inference_extras(model, d_feat, dev, B, H, W):
    """BASELINE config 5 shape of work on the PMF model: eval-mode forward (CUDA-graph replay) -> argmax -> KNN
    back-projection (pc_processor/postproc/knn.py call signature, un-batched per frame, k=5, S=5) of P = 32768 points per
    frame.  Reported next to the headline; rank 0 only, no collectives."""
    import contextlib
    import io

    import pmf_b200
    from tests import synth
    was_training = model.training
    model.eval()
    with contextlib.redirect_stdout(io.StringIO()):  # the reference's ctor banner (knn.py:40-53)
        knn = pmf_b200.KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), 20)
    case = dict(name="bench", H=H, W=W, P=32768, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.9, seed=5, kind="rand")
    inp = synth.knn_inputs(case)
    pr, ur, px, py = (torch.from_numpy(inp[k]).to(dev) for k in ("proj_range", "unproj_range", "px", "py"))

    def fwd():
        with torch.no_grad():
            return model(d_feat[:, 0:5], d_feat[:, 5:8])

    offs = torch.arange(B + 1, device=dev, dtype=torch.int64) * ur.numel()
    urb, pxb, pyb, prb = ur.repeat(B), px.repeat(B), py.repeat(B), pr.unsqueeze(0).expand(B, H, W).contiguous()

    def tail(lid):
        # on-device tail: crop-less argmax kernel + ONE batched KNN launch for the B frames (pmf_b200.postproc)
        return pmf_b200.knn_batched(knn, prb, urb, pmf_b200.argmax_nchw(lid), pxb, pyb, offs)

    def ev_time(fn, reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    for _ in range(3):
        lid, _cam = fwd()
        tail(lid)
    # median of three 5-replay measurements: a single measurement of this short loop occasionally came out 1.7x slow
    # (one run in five on the shared pool) while the replayed kernels themselves are unchanged
    runs = sorted((ev_time(fwd, 5) for _ in range(3)), key=lambda r: r[0])
    ms_fwd, (lid, _cam) = runs[1]
    ms_tail = sorted(ev_time(lambda: tail(lid), 5)[0] for _ in range(3))[1]
    model.train(was_training)
    epmf = epmf_sweep(dev, H, W, knn, (pr, ur, px, py), ev_time)
    return {"epmf_config5": epmf, "workload": "PMF-ResNet34 eval forward (batch %d, %dx%d) + argmax kernel + ONE batched KNN(k=5,S=5) launch, 32768 points/frame" % (B, H, W),
            "infer_fwd_frames_per_s": B / (ms_fwd * 1e-3), "infer_fwd_ms": ms_fwd,
            "knn_tail_ms_per_frame": ms_tail / B, "knn_points_per_s": B * 32768 / (ms_tail * 1e-3),
            "infer_plus_knn_frames_per_s": B / ((ms_fwd + ms_tail) * 1e-3),
            "fwd_tflops": FWD_KFLOP_PER_PX * 1e3 * H * W * B / (ms_fwd * 1e-3) / 1e12}


def epmf_sweep(dev, H, W, knn, knn_in, ev_time, batches=(1, 8, 32)):
    """BASELINE config 5: EPMF-ResNet34 inference-only (CUDA-graph replay) + argmax + KNN back-projection, batch sweep.
    1097.2 kFLOP/pixel algorithmic forward."""
    import pmf_b200
    from tests import synth
    torch.manual_seed(1)
    m = pmf_b200.EPMFNet(5, 3, 20, 32, False, "resnet34").to(dev).eval()
    pr, ur, px, py = knn_in
    out = []
    for B in batches:
        feat, _, _ = synth.frame_tensor(B, H, W, seed=7, density=0.1)
        x = feat.to(dev)

        def fwd():
            with torch.no_grad():
                return m(x[:, 0:5], x[:, 5:8])

        offs = torch.arange(B + 1, device=dev, dtype=torch.int64) * ur.numel()
        urb, pxb, pyb, prb = ur.repeat(B), px.repeat(B), py.repeat(B), pr.unsqueeze(0).expand(B, H, W).contiguous()

        def tail(lid):
            return pmf_b200.knn_batched(knn, prb, urb, pmf_b200.argmax_nchw(lid), pxb, pyb, offs)

        for _ in range(3):
            lid, _cam = fwd()
            tail(lid)
        ms_fwd, (lid, _cam) = ev_time(fwd, 5)
        ms_tail, _ = ev_time(lambda: tail(lid), 3)
        out.append({"batch": B, "fwd_ms": ms_fwd, "fwd_frames_per_s": B / (ms_fwd * 1e-3),
                    "fwd_plus_knn_frames_per_s": B / ((ms_fwd + ms_tail) * 1e-3),
                    "fwd_tflops": 1097.2e3 * H * W * B / (ms_fwd * 1e-3) / 1e12})
        del x
        m._graphs.clear()
        torch.cuda.empty_cache()
    return {"workload": "EPMF-ResNet34 eval forward %dx%d + argmax kernel + ONE batched KNN(k=5,S=5) launch, 32768 points/frame" % (H, W),
            "sweep": out}


def kernel_roofline(step_fn, dev, ms_per_step, peak_tf, peak_src):
    """One extra, eagerly launched step (CUDA graphs off) with CUDA events around EVERY C-ABI call on the launching
    stream: per-entry-point device time, and for the tcgen05 conv kernels the algorithmic FLOPs per launch
    (2 * pixels * c_out * c_in * taps)."""
    from pmf_b200 import _lib as L
    recs = []
    orig = L.call

    def vbytes(v, n, h, w, c):
        """Algorithmic bytes of one NHWC view operand (0 for NULL / broadcast views)."""
        v = getattr(v, "_obj", v)
        if v is None or not getattr(v, "ptr", None) or v.sx == 0:
            return 0.0
        return 4.0 * n * h * w * c

    def call_bytes(name, a):
        """ALGORITHMIC HBM bytes of one C-ABI call: every full-size operand read once, every result written once
        (DESIGN.md §3 per-kernel figures)."""
        if name == "pmfb_pointwise":
            n, h, w, c = a[5:9]
            e = a[9]._obj
            return vbytes(a[0], n, h, w, c) + 4.0 * n * h * w * c + sum(vbytes(v, n, h, w, c) for v in (e.r1, e.mul, e.r2))
        if name == "pmfb_bn_stats":
            return vbytes(a[0], *a[1:5])
        if name == "pmfb_bn_bwd_reduce":
            n, h, w, c = a[9:13]
            return sum(vbytes(v, n, h, w, c) for v in (a[0], a[1], a[2], a[4]))
        if name == "pmfb_bn_bwd_apply":
            n, h, w, c = a[12:16]
            b = sum(vbytes(v, n, h, w, c) for v in (a[0], a[1], a[2], a[4]))
            if a[16]:
                b += 4.0 * n * h * w * c
            if a[24]:
                b += 4.0 * n * h * w * c * (2 if a[28] else 1)
            return b
        if name == "pmfb_conv_fwd":
            d = a[0]._obj
            px = d.n_batch * d.out_h * d.out_w
            in_elems = 1.0
            for k in range(5):
                in_elems *= d.x.dims[k]
            return 4.0 * (in_elems + px * d.c_out * (1 + sum(1 for v in (d.epi.r1, d.epi.mul, d.epi.r2) if v.ptr and v.sx)))
        if name == "pmfb_conv_wgrad":
            d = a[0]._obj
            in_elems = 1.0
            for k in range(5):
                in_elems *= d.x.dims[k]
            return 4.0 * (in_elems + d.n_batch * d.out_h * d.out_w * d.c_out)
        if name in ("pmfb_pool3s2", "pmfb_pool3s2_bwd"):
            n, h, w, c = a[2:6]
            return 4.0 * n * h * w * c * 1.25 + (n * h * w * c / 4.0 if a[11] else 0.0)
        if name in ("pmfb_softmax_nchw", "pmfb_softmax_nchw_bwd"):
            n, h, w, c = (a[1:5] if name == "pmfb_softmax_nchw" else a[2:6])
            return 4.0 * n * h * w * c * (2 if name == "pmfb_softmax_nchw" else 3)
        return 0.0

    def call(name, *a):
        flops = 0.0
        if name in ("pmfb_conv_fwd", "pmfb_conv_wgrad"):
            d = a[0]._obj
            flops = 2.0 * d.n_batch * d.out_h * d.out_w * d.c_out * d.c_in * d.n_taps
            if d.n_taps == 7 and d.c_in == 32:
                flops *= 21.0 / 32.0  # the 7x7x3 stem runs as 7 taps over 21 real (+11 zero) unrolled channels
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        recs.append((name, flops, e0, e1, call_bytes(name, a)))
        if flops:
            shapes.append("%s n%d %dx%d cin%d cout%d taps%d" % (name[5:], d.n_batch, d.out_h, d.out_w, d.c_in, d.c_out, d.n_taps))
        else:
            shapes.append(None)

    shapes = []
    prev = os.environ.get("PMFB_CUDA_GRAPH")
    prev_side = os.environ.get("PMFB_WGRAD_STREAM")
    os.environ["PMFB_CUDA_GRAPH"] = "0"
    os.environ["PMFB_WGRAD_STREAM"] = "0"  # per-launch event timing needs every kernel on the launching stream ...
    from pmf_b200 import engine as _eng
    prev_branch = _eng.FWD_BRANCH
    _eng.FWD_BRANCH = False                # ... and alone on the GPU: no camera-stream branch next to the LiDAR stream
    try:
        step_fn()  # eager warm-up (allocator)
        torch.cuda.synchronize()
        L.call = call
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        step_fn()
        e_b.record()
        torch.cuda.synchronize()
    finally:
        L.call = orig
        _eng.FWD_BRANCH = prev_branch
        if prev is None:
            os.environ.pop("PMFB_CUDA_GRAPH", None)
        else:
            os.environ["PMFB_CUDA_GRAPH"] = prev
        if prev_side is None:
            os.environ.pop("PMFB_WGRAD_STREAM", None)
        else:
            os.environ["PMFB_WGRAD_STREAM"] = prev_side
    by = {}
    for (n, f, e0, e1, nb) in recs:
        r = by.setdefault(n, [0, 0.0, 0.0, 0.0])
        r[0] += 1
        r[1] += e0.elapsed_time(e1)
        r[2] += f
        r[3] += nb
    breakdown = {n.replace("pmfb_", ""): {"launches": r[0], "ms": round(r[1], 3), "alg_gb": round(r[3] / 1e9, 3),
                                          "alg_gbs": round(r[3] / 1e9 / max(r[1] * 1e-3, 1e-12), 1)}
                 for n, r in sorted(by.items(), key=lambda kv: -kv[1][1])}
    in_kernels = sum(r[1] for r in by.values())
    convs = sorted(((e0.elapsed_time(e1), f, sh) for (n, f, e0, e1, _nb), sh in zip(recs, shapes) if sh), reverse=True)
    top = [{"ms": round(t, 3), "tflops": round(f / (t * 1e-3) / 1e12, 1), "launch": sh} for t, f, sh in convs[:24]]
    if os.environ.get("PMFB_BENCH_DUMP"):
        with open(os.environ["PMFB_BENCH_DUMP"], "w") as fdump:
            for t, f, sh in convs:
                fdump.write("%.3f ms  %7.1f TFLOP/s  %s\n" % (t, f / (t * 1e-3) / 1e12, sh))
    out = {}
    for kind in ("pmfb_conv_fwd", "pmfb_conv_wgrad"):
        if kind in by:
            c, t, fl, _nb = by[kind]
            out[kind] = {"launches": c, "gflop": fl / 1e9, "ms": t, "tflops": fl / (t * 1e-3) / 1e12}
    dom = out.get("pmfb_conv_fwd")
    if not dom:
        return None
    traffic = None
    if os.path.exists(HALO_TRAFFIC_FILE):
        import hashlib
        raw = open(HALO_TRAFFIC_FILE, "rb").read()
        traffic = json.loads(raw)
        traffic["sha256_16"] = hashlib.sha256(raw).hexdigest()[:16]
    return {"bound": "tensor", "kernel": "conv_fwd_halo_kernel / conv_fwd_tc_kernel (tcgen05 kind::tf32 implicit GEMM: forward + dgrad launches)",
            "achieved": dom["tflops"], "peak": peak_tf, "unit": "TFLOP/s", "frac": dom["tflops"] / peak_tf,
            "peak_source": peak_src + "; kind::f16 launches (>= 64-channel layers in the default f16 mode) can reach it, kind::tf32 "
                                      "launches (thin layers, tf32 mode) issue at half that rate",
            # DRAM bytes per launch of this kernel: ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the
            # conv_fwd_halo_kernel launches of one step, read from the committed capture summary (profiles/halo_traffic.json,
            # written by tools/ncu_traffic.py from the ncu launch list); `alg_bytes_per_launch` is the algorithmic figure.
            "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_source": traffic,
            "alg_bytes_per_launch": by["pmfb_conv_fwd"][3] / max(dom["launches"], 1),
            "launches_per_step": dom["launches"], "ms_in_kernel_per_step": dom["ms"],
            "share_of_step": dom["ms"] / max(in_kernels, 1e-9),
            "wgrad": out.get("pmfb_conv_wgrad"), "cabi_ms_per_step": in_kernels,
            "eager_step_ms": e_a.elapsed_time(e_b), "breakdown": breakdown, "top_conv_launches": top}


def main():
    global NCLASSES, BACKBONE, METRIC, FWD_KFLOP_PER_PX, STEP_KFLOP_PER_PX
    args = parse()
    NCLASSES, BACKBONE = args.nclasses, args.backbone
    if args.backbone == "resnet50":  # PMF-ResNet50 forward 2156.8 kFLOP/px, forward + backward ~6450
        FWD_KFLOP_PER_PX, STEP_KFLOP_PER_PX = 2156.8, 6450.0
    if (args.backbone, args.batch, args.height, args.width) != ("resnet34", 8, 480, 640):
        METRIC = "frames/sec PMF-%s fwd+bwd (%dx%d camera grid, batch %d/GPU)" % (
            args.backbone.replace("resnet", "ResNet"), args.height, args.width, args.batch)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "gpu-eager":
        run_gpu_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
