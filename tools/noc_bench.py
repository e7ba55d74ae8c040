"""Batched NoC evaluation loop on synthetic images, sharded over ranks.

    python tools/noc_bench.py --arch vit_huge --images 64 --clicks 20 --micro-batch 32
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/noc_bench.py ...   (N GPUs)

Every rank evaluates a contiguous block of images in lock-step micro-batches (model batch = 2 x sessions with flip
TTA); the only collective is the final all_gather of the [images, clicks] IoU table.  Prints one JSON line on rank 0:
click-forwards/s of the whole job including all host work of the loop (clicker, ZoomIn, metrics, uploads); the synthetic
images are generated before the clock starts."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200.config import make_config  # noqa: E402
from pvpuformer_b200.inference import compute_noc_metric  # noqa: E402
from pvpuformer_b200.inference.datasets import SyntheticEllipseDataset  # noqa: E402
from pvpuformer_b200.inference.evaluation import evaluate_lockstep, evaluate_sharded  # noqa: E402
from pvpuformer_b200.model import build_model  # noqa: E402
from pvpuformer_b200.weights import synthetic_state_dict  # noqa: E402


class _Materialised:
    """This rank's samples generated before the timed region (the synthetic images stand for decoded images in host memory;
    drawing 600k random numbers per image is not part of the evaluation loop being measured)."""

    def __init__(self, ds, rank, world):
        from pvpuformer_b200.inference.evaluation import shard_range
        self.ds = ds
        a, b = shard_range(len(ds), rank, world)
        self.cache = {i: ds.get_sample(i) for i in range(a, b)}

    def __len__(self):
        return len(self.ds)

    def get_sample(self, index):
        return self.cache[index]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vit_huge")
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--clicks", type=int, default=20)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--host-clicker", action="store_true", help="cv2 distance-transform clicker + IoU on the host (default: csrc/noc.cu)")
    ap.add_argument("--host-transforms", action="store_true",
                    help="per-session Python predictors (ZoomIn / flip / sigmoid as torch ops) instead of the device sessions of csrc/session.cu")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL logs (its version banner included) go to STDOUT by default:
                                                                  # keep the one JSON line alone there
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(args.arch)
    model = build_model(args.arch, state_dict=synthetic_state_dict(cfg, 0), device=dev)
    model.want_aux = False                        # NoBRS reads only ['instances'] (reference predictors/base.py:177)
    ds = _Materialised(SyntheticEllipseDataset(args.images), rank, world)
    # warm-up: one small shard-independent pass (weights packed, workspaces allocated)
    wds = SyntheticEllipseDataset(2, seed0=10_000)
    evaluate_lockstep([(wds.get_sample(i).image, wds.get_sample(i).gt_mask(1)) for i in range(2)], model, dev, 1.01, max_clicks=2,
                      micro_batch=2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    table, local_s, stats = evaluate_sharded(ds, model, dev, rank, world, 1.01, max_clicks=args.clicks,
                                             micro_batch=args.micro_batch, gather_device=dev if world > 1 else None,
                                             device_clicker=not args.host_clicker,
                                             device_session=not (args.host_clicker or args.host_transforms))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([total_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_s = t.item()
    if rank == 0:
        noc, _, over = compute_noc_metric([r[np.isfinite(r)] for r in table], [0.8, 0.85, 0.9], max_clicks=args.clicks)
        fwd = 2 * args.images * args.clicks
        print(json.dumps({"metric": "click-forwards/sec (NoC loop, host work included)", "value": fwd / total_s, "n_gpus": world,
                          "arch": args.arch, "images": args.images, "clicks": args.clicks, "micro_batch": args.micro_batch, "clicker": "host" if args.host_clicker else "device",
                          "transforms": "host" if (args.host_clicker or args.host_transforms) else "device",
                          "seconds": total_s, "rank0_loop_seconds": local_s, "rank0_network_calls": stats["network_calls"],
                          "noc@80/85/90": [float(x) for x in noc], "iou_table_shape": list(table.shape)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
