"""Stage-by-stage parity report: CUDA path (workspace taps) vs the CPU oracle on one small batch.
Debug/evidence tool (imports oracle/ => test infrastructure, not product).  Usage on the GPU box:
    python tools/stage_report.py [vit_base|vit_large|vit_huge] [B]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, vpu_oracle as vo  # noqa: E402
from pvpuformer_b200 import lib as L  # noqa: E402
from pvpuformer_b200.config import make_config  # noqa: E402
from pvpuformer_b200.model import build_model  # noqa: E402
from pvpuformer_b200.weights import synthetic_state_dict  # noqa: E402


def err(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    print("%-22s max|d|=%.3e  mean|d|=%.3e  ref std=%.3e max=%.3e  rel=%.3e %s" % (
        name, d.max().item(), d.mean().item(), ref.std().item(), ref.abs().max().item(),
        d.max().item() / max(ref.std().item(), 1e-12), "NaN!" if torch.isnan(got).any() else ""), flush=True)


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "vit_base"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = make_config(arch)
    sd = synthetic_state_dict(cfg, 0)
    dev = torch.device("cuda:0")
    m = build_model(arch, state_dict=sd, device=dev)
    m.graph_max_batch = 0      # the debug early exit below changes what a forward launches: no CUDA-graph replay here
    image4 = cases.images(B, seed=1)
    pts = cases.random_clicks(B, seed=6, dtype=torch.float64)
    taps = {}
    with torch.no_grad():
        ref = vo.forward(sd, cfg, image4, pts, taps=taps)
    C, N, g, Q = cfg.embed_dim, cfg.num_tokens, cfg.grid, cfg.num_queries
    img_d = image4.to(dev)
    # tokens after selected ViT blocks via the debug early exit
    lib = L.load()
    m._ensure_ready(dev)
    for k, name in ((0, "tokens_embed"), (1, "tokens_block1"), (cfg.blocks_per_group, "tokens_block%d" % cfg.blocks_per_group)):
        L.check(lib.vpu_set_scalar(m._handle, b"debug.stop_after_block", float(k)))
        m(img_d, pts)
        torch.cuda.synchronize()
        err(name, m.tap("X", B, torch.float32, (B, N, C)), taps[name])
    L.check(lib.vpu_set_scalar(m._handle, b"debug.stop_after_block", -1.0))
    out = m(img_d, pts)
    torch.cuda.synchronize()
    err("backbone_features", m.tap("X", B, torch.float32, (B, N, C)), taps["backbone_features"])
    err("ppue", m.tap("ppue", B, torch.float32, (B, Q, cfg.ppue_dim)), taps["ppue"])
    err("q_ffn", m.tap("Q0", B, torch.float32, (B, Q, C)), taps["q_ffn"])
    err("dma_q_final", m.tap("qfin", B, torch.float32, (B, Q, C)), taps["dma_q_final"])
    err("dma_k_final", m.tap("Kb", B, torch.bfloat16, (B, N, C)), taps["dma_k_final"])
    err("q_out", m.tap("qout", B, torch.float32, (B, Q, C)), taps["q_out"])
    od = cfg.out_dims
    for name, tapn, r, c in (("pyr4", "P4", 4 * g, od[0]), ("pyr8", "P8", 2 * g, od[1]), ("pyr16", "P16", g, od[2]),
                             ("pyr32", "P32", g // 2, od[3])):
        err(name, m.tap(tapn, B, torch.bfloat16, (B, r, r, c)).permute(0, 3, 1, 2), taps[name])
    # "head_feat" (the fused 256-channel feature map) has no tap since round 2: head_tail.cu keeps it in TMEM; its two consumers
    # (seg_lowres, aux_lowres) are compared below
    err("seg_lowres", m.tap("seg_low", B, torch.float32, (B, 1, 4 * g, 4 * g)), taps["seg_lowres"])
    err("aux_lowres", m.tap("aux_low", B, torch.float32, (B, Q, 4 * g, 4 * g)), taps["aux_lowres"])
    err("instances", out["instances"], ref["instances"])
    err("instances_aux", out["instances_aux"], ref["instances_aux"])
    a = torch.sigmoid(out["instances"].cpu()) > 0.49
    b = torch.sigmoid(ref["instances"]) > 0.49
    print("mask IoU @0.49: %.6f" % ((a & b).sum().item() / max((a | b).sum().item(), 1)))
    med = ref["instances"].median()
    a, b = out["instances"].cpu() > med, ref["instances"] > med
    print("mask IoU @median logit: %.6f" % ((a & b).sum().item() / max((a | b).sum().item(), 1)))


if __name__ == "__main__":
    main()
