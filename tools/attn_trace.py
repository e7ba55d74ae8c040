"""Timeline of CTA 0 of the tcgen05 global-attention kernel (clock64 stamps logged by the kernel itself).
Usage on the GPU box: python tools/attn_trace.py [blocks]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import lib as L, ops  # noqa: E402

NAMES = {0: {1: "kv_empty ok"}, 1: {1: "s_free[0]", 2: "s_free[1]", 3: "issue S[0]", 4: "issue S[1]", 5: "p_full[0]", 6: "p_full[1]", 7: "S[0] issued", 8: "S[1] issued", 9: "PV[0] issued", 10: "PV[1] issued"},
         2: {1: "s_full", 2: "S in regs", 3: "exps done", 4: "P stored", 5: "p arrive"}}
NAMES[3] = NAMES[2]


def main(nshow=60):
    dev = torch.device("cuda:0")
    B, heads, hd, N = 64, 12, 64, 784
    C = heads * hd
    qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
    cap = 4096
    buf = torch.zeros(4 * cap, dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, hd ** -0.5, 0, C, 2 * C)
    torch.cuda.synchronize()
    L.check(L.load().vpu_debug_attention_trace(L.ptr(buf), cap))
    ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, hd ** -0.5, 0, C, 2 * C)
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
    t0 = ev[0][0]
    role_name = ["TMA ", "MMA ", "WG0 ", "WG1 "]
    skip = int(os.environ.get("SKIP", "400"))
    for clk, role, code in ev[skip:skip + nshow]:
        print("%8d  %s %s%s" % (clk - t0, role_name[role], "    " * role, NAMES[role].get(code, str(code))))
    # per-block period of WG0
    wg0 = [c for c, r, k in ev if r == 2 and k == 1]
    d = [b - a for a, b in zip(wg0, wg0[1:])]
    if d:
        d.sort()
        print("WG0 s_full period: median %d clk, p10 %d, p90 %d over %d blocks" % (d[len(d) // 2], d[len(d) // 10], d[9 * len(d) // 10], len(d)))
    for role, code_a, code_b, label in [(2, 1, 2, "WG0 s_full -> S in regs"), (2, 2, 3, "WG0 S in regs -> exps done"),
                                        (2, 3, 4, "WG0 exps done -> P stored (incl. pv_done wait)"), (2, 4, 5, "WG0 P stored -> arrive"),
                                        (2, 5, 1, "WG0 arrive -> next s_full")]:
        seq = [(c, k) for c, r, k in ev if r == role]
        ds = [b[0] - a[0] for a, b in zip(seq, seq[1:]) if a[1] == code_a and b[1] == code_b]
        if ds:
            ds.sort()
            print("%-50s median %5d  p90 %5d" % (label, ds[len(ds) // 2], ds[9 * len(ds) // 10]))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60)
