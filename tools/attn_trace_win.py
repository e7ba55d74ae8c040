"""Timeline of CTA 0 of the ViT-H tcgen05 window-attention kernel (clock64 stamps logged by a -DVPU_ATTN_DEBUG build).
Usage on the GPU box: VPU_LIB_PATH=<debug build> python tools/attn_trace_win.py [events to print]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import lib as L, ops  # noqa: E402

NAMES = {0: {1: "qk_empty ok", 2: "v_empty ok"},
         1: {1: "p_full[0]", 9: "v_full", 2: "PV0 issued", 10: "qk_full next", 3: "s_empty[0]", 4: "S0' issued", 5: "p_full[1]", 6: "PV1 issued",
             7: "s_empty[1]", 8: "S1' issued"},
         2: {1: "s_full", 2: "max done", 3: "P stored + arrive", 4: "o_full", 5: "epilogue done", 6: "O in regs", 7: "s_empty arrive", 8: "staging tile free"}}
NAMES[3] = NAMES[2]


def main(nshow=80):
    dev = torch.device("cuda:0")
    B, heads, hd, grid, win = (32, 16, 80, 32, 16) if os.environ.get("ARCH", "h") == "h" else (64, 12, 64, 28, 14)
    N, C = grid * grid, heads * hd
    qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
    cap = 4096
    buf = torch.zeros(4 * cap, dtype=torch.int64, device=dev)
    run = lambda: ops.attention(qkv, qkv, qkv, win * win, win * win, heads, hd, B * 4, hd ** -0.5, 0, C, 2 * C, window=win, grid=grid)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    L.check(L.load().vpu_debug_attention_trace(L.ptr(buf), cap))
    run()
    torch.cuda.synchronize()
    L.check(L.load().vpu_debug_attention_trace(None, 0))
    t = buf.cpu().view(4, cap)
    ev = []
    for role in range(4):
        for x in t[role].tolist():
            if x == 0:
                break
            ev.append((x & ((1 << 56) - 1), role, (x >> 56) & 0xFF))
    ev.sort()
    if not ev:
        print("no events: the library was not built with -DVPU_ATTN_DEBUG")
        return
    t0 = ev[0][0]
    role_name = ["TMA ", "MMA ", "WG0 ", "WG1 "]
    skip = int(os.environ.get("SKIP", "150"))
    for clk, role, code in ev[skip:skip + nshow]:
        print("%8d  %s %s%s" % (clk - t0, role_name[role], "    " * role, NAMES[role].get(code, str(code))))
    print("total span %d clk, %d events" % (ev[-1][0] - t0, len(ev)))

    def gaps(role, ca, cb, label):
        seq = [(c, k) for c, r, k in ev if r == role]
        ds = []
        last = None
        for c, k in seq:
            if k == ca:
                last = c
            elif k == cb and last is not None:
                ds.append(c - last)
                last = None
        if ds:
            ds.sort()
            print("%-46s median %6d  p90 %6d  n %d" % (label, ds[len(ds) // 2], ds[9 * len(ds) // 10], len(ds)))
    wg = [c for c, r, k in ev if r == 2 and k == 1]
    d = sorted(b - a for a, b in zip(wg, wg[1:]))
    if d:
        print("WG0 problem period: median %d clk, p10 %d, p90 %d over %d problems" % (d[len(d) // 2], d[len(d) // 10], d[9 * len(d) // 10], len(d)))
    for role in (2, 3):
        n = role_name[role]
        gaps(role, 1, 2, n + "s_full -> max done")
        gaps(role, 2, 3, n + "max done -> P stored")
        gaps(role, 3, 4, n + "P stored -> o_full (PV)")
        gaps(role, 4, 6, n + "o_full -> O in registers")
        gaps(role, 6, 7, n + "O in registers -> s_empty arrive")
        gaps(role, 4, 7, n + "o_full -> s_empty arrive")
        gaps(role, 7, 5, n + "s_empty arrive -> stores issued")
        gaps(role, 7, 8, n + "  s_empty arrive -> staging tile free")
        gaps(role, 8, 5, n + "  scale + pack + st.shared + hand-off")
        gaps(role, 5, 1, n + "epilogue done -> next s_full")
    gaps(1, 1, 2, "MMA p_full[0] -> PV0 issued")
    gaps(1, 1, 9, "MMA p_full[0] -> v_full")
    gaps(1, 9, 2, "MMA v_full -> PV0 issued")
    gaps(1, 2, 10, "MMA PV0 issued -> qk_full next")
    gaps(1, 10, 3, "MMA qk_full next -> s_empty[0]")
    gaps(1, 3, 4, "MMA s_empty[0] -> S0' issued")
    gaps(1, 4, 5, "MMA S0' issued -> p_full[1]")
    gaps(1, 5, 6, "MMA p_full[1] -> PV1 issued")
    gaps(1, 6, 7, "MMA PV1 issued -> s_empty[1]")
    gaps(1, 7, 8, "MMA s_empty[1] -> S1' issued")
    gaps(1, 8, 1, "MMA S1' issued -> next p_full[0]")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 80)
