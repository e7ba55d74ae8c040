#!/usr/bin/env python
"""Benchmark of the per-click VPUFormer forward (BASELINE.json metric: click-forwards/s, ViT-B/448).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

One step = one pass of the hot path (forward(image, points) -> instances + instances_aux) over one batch
of synthetic input: BASELINE.json configs[1], ViT-B/448 batch 64 per GPU (weak scaling: every rank runs
its own 64 click-forwards; no collective in the forward).  Prints ONE JSON line on rank 0.

  value        click-forwards/s, inputs resident in HBM, CUDA events over exactly K steps, max over ranks
  e2e          same metric through the public module call with HOST (pinned) inputs: H2D of the image and
               click tensors and D2H of `instances` inside the timed region, every step
  roofline     tcgen05 GEMM kernel class: algorithmic FLOPs / CUDA-event time of its launches (events
               recorded by the C ABI around every launch on the launching stream), vs MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (port of the reference forward) on this host's cores, bounded sample
  noc_loop     BASELINE.json configs[3]: ViT-H, 20-click NoC evaluation loop over a FIXED total of 1024 synthetic images (strong
               scaling), flip TTA, micro-batch 32, device-resident click sessions, images rank-sharded, the final NCCL
               all_gather of the IoU table inside the timed region; click-forwards/s, CUDA events, max over ranks
  gpu_eager_bar  (N=1) the same forward as plain torch eager ops (cuBLAS / cuDNN) on this B200 in fp32, TF32 and bf16 autocast:
               the GPU bar SURVEY.md 8(d) names; a comparator like cpu_baseline, never the product path
  latency_ms_b2  one NoBRS click with flip TTA (batch 2, BASELINE.json configs[0] shape) through the module call
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "click-forwards/sec"
GFLOP_PER_CLICK_FORWARD = {"vit_base": 170.7, "vit_large": 538.7, "vit_huge": 1422.8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="vit_base", choices=["vit_base", "vit_large", "vit_huge"])
    ap.add_argument("--batch", type=int, default=64, help="click-forwards per step per GPU")
    ap.add_argument("--no-aux", action="store_true", help="skip the 48-channel aux output (NoBRS only reads instances)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--ref-batch", type=int, default=8, help="click-forwards per CPU step (--impl reference)")
    ap.add_argument("--no-noc", action="store_true", help="skip the config-4 NoC loop")
    ap.add_argument("--noc-arch", default="vit_huge", choices=["vit_base", "vit_large", "vit_huge"])
    ap.add_argument("--noc-images", type=int, default=1024, help="TOTAL images of the NoC loop (fixed as N grows: strong scaling)")
    ap.add_argument("--noc-clicks", type=int, default=20)
    ap.add_argument("--noc-micro-batch", type=int, default=32)
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU bar")
    ap.add_argument("--eager-steps", type=int, default=5)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def workload(arch, batch, seed):
    from pvpuformer_b200 import synthetic
    image4 = synthetic.images(batch, seed=seed)
    points = synthetic.random_clicks(batch, seed=seed + 1, dtype=torch_mod().float64)
    return image4, points


def torch_mod():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm for the path, restated in oracle/vpu_oracle.py (the Python
# reference tree itself does not travel to the GPU box).  Times whole forwards on the host cores.
# ---------------------------------------------------------------------------------------------
def cpu_forward_rate(arch, batch, steps, warmup, budget_s=None):
    import torch
    from oracle import vpu_oracle as vo
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.weights import synthetic_state_dict
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    cfg = make_config(arch)
    sd = synthetic_state_dict(cfg, 0)
    image4, points = workload(arch, batch, seed=11)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vo.forward(sd, cfg, image4, points)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if budget_s is not None and i >= warmup and sum(times) > budget_s:
                break
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": "%d steps x %d click-forwards of the same synthetic %s workload, fp32, %d torch threads"
                      % (len(times), batch, arch, cores)}, total / len(times), len(times)


def run_reference(args, rank):
    if rank != 0:
        return
    cb, step_s, nsteps = cpu_forward_rate(args.arch, args.ref_batch, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": METRIC, "n_gpus": args.gpus,
            "steps": nsteps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "VPUFormer %s 448 per-click forward; CPU step = %d click-forwards" % (args.arch, args.ref_batch),
                       "arch": args.arch, "batch_per_step": args.ref_batch},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up; only samples stamped inside the timed window count."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, pw, mx, reasons, total = [], [], None, set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            total += 1
            if self.t0 is not None and not (self.t0 <= ts <= self.t1):
                continue
            sm.append(clk)
            mx = mxc
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "samples_total": total,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


def profile_classes(model, L, steps, run_step):
    import torch
    lib = L.load()
    L.check(lib.vpu_profile_begin(model._handle))
    for _ in range(steps):
        run_step()
    torch.cuda.synchronize()
    arr = (L.VpuProfileEntry * 64)()
    n = ctypes.c_int(0)
    L.check(lib.vpu_profile_end(model._handle, arr, 64, ctypes.byref(n)))
    out = []
    for i in range(n.value):
        e = arr[i]
        out.append({"name": e.name.decode(), "ms": e.ms / steps, "flops": e.flops / steps, "bytes": e.bytes / steps,
                    "launches": e.launches // steps})
    return out


def run_noc_loop(args, rank, world, dev, pk):
    """BASELINE.json configs[3] / SURVEY.md 8(d) config 4 (reference isegm/inference/vpu_evaluation.py:18-98 driven by
    scripts/evaluate_vpumodel.py:87-88,187-192): the 20-click NoC loop, NoBRS predictor with flip TTA and 448-px zoom-in, over a
    FIXED total of --noc-images synthetic images sharded contiguously over the ranks (strong scaling).  Every click of a
    micro-batch of 32 sessions is clicker -> prepare -> forward(batch 64) -> finish on device-resident state; the loop never
    stops early (IoU threshold 1.01), so the job is exactly 2 * images * clicks click-forwards.  Timed with CUDA events on the
    launching stream from the first clicker launch to the end of the NCCL all_gather of the IoU table, max over ranks; the
    synthetic images are generated before the clock starts (they stand for decoded images in host memory)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pvpuformer_b200 import lib as L
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.inference import compute_noc_metric
    from pvpuformer_b200.inference.datasets import MaterialisedShard, SyntheticEllipseDataset
    from pvpuformer_b200.inference.evaluation import evaluate_lockstep, evaluate_sharded, shard_range
    from pvpuformer_b200.model import build_model
    from pvpuformer_b200.weights import synthetic_state_dict
    arch, images, clicks, mb = args.noc_arch, args.noc_images, args.noc_clicks, args.noc_micro_batch
    cfg = make_config(arch)
    model = build_model(arch, state_dict=synthetic_state_dict(cfg, 0), device=dev)
    model.want_aux = False                         # NoBRS reads only ['instances'] (reference predictors/base.py:177)
    ds = MaterialisedShard(SyntheticEllipseDataset(images), rank, world)
    a, b = shard_range(images, rank, world)
    # warm-up: 3 clicks of one full micro-batch (weights packed, batch-64 workspace allocated, every kernel loaded)
    wds = SyntheticEllipseDataset(mb, seed0=10_000_000)
    wsamples = [(wds.get_sample(i).image, wds.get_sample(i).gt_mask(1)) for i in range(mb)]
    evaluate_lockstep(wsamples, model, dev, 1.01, max_clicks=3, micro_batch=mb, device_session=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    lib = L.load()
    l0 = lib.vpu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    table, local_s, stats = evaluate_sharded(ds, model, dev, rank, world, 1.01, max_clicks=clicks, micro_batch=mb,
                                             gather_device=dev if world > 1 else None, device_clicker=True, device_session=True)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = lib.vpu_launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = t[0].item(), t[1].item() * 1e-3
    del model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    fwd = 2 * images * clicks
    noc, _, _ = compute_noc_metric([r[np.isfinite(r)] for r in table], [0.8, 0.85, 0.9], max_clicks=clicks)
    rate = fwd / (ms * 1e-3)
    return {"metric": "click-forwards/sec (NoC loop: clicker, zoom-in, flip TTA, forward, paste, IoU; all_gather included)",
            "value": rate, "unit": METRIC, "scaling": "strong", "n_gpus": world, "arch": arch, "images_total": images,
            "images_per_rank": [shard_range(images, r, world)[1] - shard_range(images, r, world)[0] for r in range(world)],
            "clicks": clicks, "micro_batch_sessions": mb, "model_batch": 2 * mb, "flip_tta": True, "device_sessions": True,
            "click_forwards": fwd, "seconds": ms * 1e-3, "host_wall_seconds": wall, "rank0_network_calls": stats["network_calls"],
            "rank0_launches": int(launches), "collective": "one all_gather of row counts + one of the [rows, %d] fp32 IoU table "
            "(%s), inside the timed region" % (clicks, "NCCL" if world > 1 else "single rank: no-op"),
            "iou_table_shape": list(table.shape), "noc@80/85/90": [float(x) for x in noc],
            "step_frac_of_tensor_peak": rate / world * GFLOP_PER_CLICK_FORWARD[arch] * 1e9 / 1e12 / pk["tf_sustained"],
            "timing": "CUDA events on the launching stream around the whole loop (host work of the loop included), max over ranks"}


def gpu_eager_bar(args, dev, b200_value, b200_dtype="bf16"):
    """The GPU bar SURVEY.md 8(d) / BASELINE.md 3 name: the same module as plain torch eager ops (cuBLAS GEMMs, cuDNN
    convolutions, unfused softmax / LayerNorm / GELU) on this B200, batch = --batch, in fp32 (no TF32), TF32 and bf16 autocast.
    It runs the restatement of the reference forward (oracle/vpu_oracle.py, pinned bit-exact against the unmodified reference)
    with its tensors on the device -- /root/reference itself does not travel to the GPU box.  Generous to the bar: the
    reference's per-point host loops that build the PPuE rows (ops.py:80-104) are done once, before the timed region.
    A comparator like cpu_baseline: nothing of this repo's product path runs here, and nothing here runs in the product."""
    import torch
    from oracle import vpu_oracle as vo
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.weights import synthetic_state_dict
    cfg = make_config(args.arch)
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(cfg, 0).items()}
    image4, points = workload(args.arch, args.batch, seed=100)
    rows = vo.ppue(points, None, 0, cfg.img_size, cfg.num_max_points).float().to(dev)
    image_d, points_d = image4.to(dev), points.to(dev)
    want_aux = not args.no_aux
    out = {"batch": args.batch, "arch": args.arch, "steps": args.eager_steps, "warmup": 3, "torch": torch.__version__,
           "what": "oracle/vpu_oracle.forward (restated reference forward) as torch eager CUDA ops, PPuE rows precomputed"}
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, tf32, autocast in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32

            def step():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    return vo.forward(sd, cfg, image_d, points_d, want_aux=want_aux, ppue_rows=rows)
            try:
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.eager_steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.eager_steps
                v = args.batch / (ms * 1e-3)
                out[name] = {"value": v, "unit": METRIC, "ms_per_step": ms, "b200_over_eager": b200_value / v}
            except Exception as ex:            # an eager mode that cannot run (e.g. out of memory) is reported, not hidden
                out[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
                torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    del sd
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    from pvpuformer_b200 import lib as L
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.model import build_model
    from pvpuformer_b200.weights import synthetic_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL logs (its version banner included) go to STDOUT by default:
                                                                  # keep the one JSON line alone there
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(args.arch)
    model = build_model(args.arch, state_dict=synthetic_state_dict(cfg, 0), device=dev)
    model.want_aux = not args.no_aux
    B = args.batch
    image_h, points_h = workload(args.arch, B, seed=100 + rank)       # every rank: its own click sessions
    image_h, points_h = image_h.pin_memory(), points_h.pin_memory()
    image_d, points_d = image_h.to(dev), points_h.to(dev)
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return model(image_d, points_d)

    out_h = torch.empty(B, 1, cfg.img_size, cfg.img_size, dtype=torch.float32)      # shape of the per-step D2H result

    from pvpuformer_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, dev, depth=3)

    # e2e operands as a data loader hands them over: decoded uint8 HWC images + the fp32 previous masks + the click rows,
    # all in pinned host memory; the predictor's ToTensor (x / 255) runs on the device (vpu_image_from_u8)
    img_u8_h = (image_h[:, :3] * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    prev_h = image_h[:, 3].contiguous().pin_memory()

    def step_e2e():
        # every step: H2D of this step's images, previous masks and click rows (pinned), ToTensor + forward on the device, D2H of
        # 'instances' into pinned host memory; the pipeline keeps up to 3 steps in flight so that the copies of neighbouring
        # steps overlap the forward
        pipe.submit(img_u8_h, points_h, prev_mask_host=prev_h)

    def timed(fn, steps, warmup, sample_clocks=False, drain=None):
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        barrier()
        if sampler:
            sampler.begin()
        l0 = lib.vpu_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()                                    # every step's result has landed on the host
        e1.record()
        barrier()
        if sampler:
            sampler.end()
        launches = lib.vpu_launch_count() - l0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, clocks

    ms, launches, clocks = timed(step_resident, args.steps, max(args.warmup, 3), sample_clocks=True)
    value = world * B * args.steps / (ms * 1e-3)
    e2e = None
    value_no_aux = None
    if not args.no_e2e:
        # the e2e call returns only 'instances' to the host (what the NoBRS predictor reads, base.py:177), so it does not compute
        # the 48-channel aux upsample; `value_no_aux` is the device-timed rate of exactly that workload
        aux0 = model.want_aux
        model.want_aux = False
        ms_na, _, _ = timed(step_resident, args.steps, 3)
        value_no_aux = world * B * args.steps / (ms_na * 1e-3)
        ms_e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3), drain=pipe.drain)
        model.want_aux = aux0
        e2e = {"value": world * B * args.steps / (ms_e * 1e-3), "unit": METRIC,
               "h2d_bytes_per_step": img_u8_h.numel() + prev_h.numel() * 4 + points_h.numel() * 8,
               "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": ms_e / args.steps, "want_aux": False,
               "value_resident_same_outputs": value_no_aux,
               "call": "pipeline.HostPipeline(model).submit(images_u8_nhwc, points, prev_mask_host=prev): pinned host tensors -> "
                       "H2D -> vpu_image_from_u8 (ToTensor) -> VitMultiGaussianVector_ed_Model.forward (instances only) -> D2H of "
                       "'instances' into pinned host memory, 3 steps in flight"}

    # one NoBRS click with flip TTA: batch 2, instances only (BASELINE.json configs[0] shape; launch-latency-bound)
    latency_b2 = None
    if rank == 0:
        aux0 = model.want_aux
        model.want_aux = False
        i2, p2 = image_d[:2].contiguous(), points_d[:2].contiguous()
        for _ in range(5):
            model(i2, p2)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(50):
            model(i2, p2)
        eb.record()
        torch.cuda.synchronize()
        latency_b2 = ea.elapsed_time(eb) / 50
        model.want_aux = aux0

    classes = profile_classes(model, L, args.profile_steps, step_resident) if (rank == 0 and args.profile_steps > 0) else []
    pk = peaks()
    # release the ViT-B job before the ViT-H loop
    del pipe
    model._ws = {}
    torch.cuda.empty_cache()
    noc = None
    if not args.no_noc:
        noc = run_noc_loop(args, rank, world, dev, pk)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    eager = None
    if world == 1 and not args.no_eager:
        eager = gpu_eager_bar(args, dev, value)

    pk = peaks()
    gemm = [c for c in classes if c["name"].startswith("gemm.")]
    attn = [c for c in classes if c["name"].startswith("attn.")]
    tot_ms = sum(c["ms"] for c in classes) or 1.0

    def tf(cs):
        t = sum(c["ms"] for c in cs)
        return (sum(c["flops"] for c in cs) / (t * 1e-3) / 1e12) if t > 0 else 0.0

    g_tf = tf(gemm)
    # DRAM bytes per launch of the GEMM class: ncu cannot run inside this process, so the figure is read from the newest
    # committed ncu capture of this same command (profiles/gemm_traffic_r*.json, written by tools/ncu_traffic.py) and stamped
    # with its file name; the algorithmic bytes per launch (operands + outputs once) are booked live by the C ABI
    traffic, traffic_src = None, None
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "gemm_traffic_r*.json")), key=os.path.getmtime)
    if cands and args.arch == "vit_base" and B == 64 and not args.no_aux:
        traffic = json.load(open(cands[-1])).get("traffic_bytes_per_launch")
        traffic_src = "profiles/" + os.path.basename(cands[-1])
    n_gemm = sum(c["launches"] for c in gemm) or 1
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel<BN> (tcgen05.mma + TMA, all GEMMs of the forward)",
                "achieved": g_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": g_tf / pk["tf_sustained"],
                "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % pk["src"],
                "share_of_step": sum(c["ms"] for c in gemm) / tot_ms,
                "launches_per_step": sum(c["launches"] for c in gemm), "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, class average)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": sum(c["bytes"] for c in gemm) / n_gemm,
                "algorithmic_flops_per_launch": sum(c["flops"] for c in gemm) / n_gemm}
    attn_info = {}
    for c in attn:
        attn_info[c["name"]] = {"ms_per_step": c["ms"], "tflops": tf([c]), "frac_of_tensor_peak": tf([c]) / pk["tf_sustained"],
                                "launches": c["launches"]}
    kernels = [{"name": c["name"], "ms_per_step": round(c["ms"], 4), "share": round(c["ms"] / tot_ms, 4),
                "tflops": round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 1) if c["flops"] and c["ms"] > 0 else None,
                "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1) if c["ms"] > 0 else None, "launches": c["launches"]}
               for c in sorted(classes, key=lambda c: -c["ms"])]

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_forward_rate(args.arch, args.ref_batch, 3, 1, budget_s=25.0)

    step_tflops = world * B * GFLOP_PER_CLICK_FORWARD[args.arch] * 1e9 * args.steps / (ms * 1e-3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "VPUFormer %s 448 batched per-click forward, batch %d synthetic images per GPU "
                                   "(BASELINE.json configs[1]); random-init weights; outputs instances%s" %
                                   (args.arch, B, "" if args.no_aux else " + instances_aux"),
                       "arch": args.arch, "batch_per_gpu": B, "clicks": "1..20 per image", "parallelism": "dp%d (independent click sessions, no collective)" % world,
                       "l2": "inputs + workspace per step (>4 GB) exceed the 126 MB L2; no explicit flush"},
            "step_tflops_algorithmic": step_tflops, "step_frac_of_tensor_peak": step_tflops / world / pk["tf_sustained"],
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "attention": attn_info,
            "kernels": kernels,
            "kernels_note": "per-class times are CUDA-event brackets around every launch of %d profiled steps; the brackets serialise "
                            "launches and add ~0.4 us each, so the classes sum to a few %% more than ms_per_step" % args.profile_steps,
            "latency_ms_b2": latency_b2, "noc_loop": noc, "gpu_eager_bar": eager, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
