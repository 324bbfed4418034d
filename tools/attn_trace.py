"""Phase timeline of the two-tile FMHA's softmax warps (warps 0 and 4 of CTA 0: the two query tiles' warps that share
scheduler 0 and its MUFU lanes).  Answers: do the two warps' exponential phases overlap (MUFU shared) or alternate, how
long are the non-MUFU phases, does a start offset ("fmha_stagger") persist?
   python tools/attn_trace.py [--impls 2,4] [--staggers 0,900]  ->  gpurun_out/attn_trace.json
Events per key block j (clock64 low 32 bits): 0 loop top, 1 S(j) ready, 2 S(j) in registers, 3 row max known,
4 exponentials done (impl 4: of the first quarter), 5 PV(j-1) retired, 6 P stored, 7 p_ready arrived."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impls", default="2,4")
ap.add_argument("--staggers", default="0,900")
ap.add_argument("--S", type=int, default=9216)
ap.add_argument("--bk", type=int, default=128, help="key block of the traced kernel (impl 7: 64; it also records its two MMA warps)")
a = ap.parse_args()
S, heads, imgs = a.S, 5, 25
C = heads * 64
n_kv = (S + a.bk - 1) // a.bk
NW = 4 if a.bk == 64 else 2
torch.manual_seed(0)
qkv = torch.randn(imgs * S, 3 * C, device="cuda", dtype=torch.float16)
out = torch.empty(imgs * S, C, device="cuda", dtype=torch.float16)
lib = native.load()
res = []
for impl in [int(x) for x in a.impls.split(",")]:
    for st in [int(x) for x in a.staggers.split(",")]:
        native.set_tuning("fmha_stagger", st)
        buf = torch.zeros(NW * n_kv * 8, dtype=torch.int32, device="cuda")
        for _ in range(2):   # second launch: warm
            lib.svdpp_debug_attn_trace(buf.data_ptr())
            native.attn_spatial(out, qkv, n_img=imgs, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125, impl=impl)
            torch.cuda.synchronize()
        lib.svdpp_debug_attn_trace(None)
        t = buf.cpu().numpy().astype("int64").reshape(NW, n_kv, 8) & 0xFFFFFFFF
        t0 = int(t[:2, 0, 0].min())
        rel = ((t - t0) & 0xFFFFFFFF)
        per_block = [float((rel[w, -1, 7] - rel[w, 0, 0]) / n_kv) for w in range(2)]
        # mean duration of each phase (event k-1 -> k) over the steady blocks
        ph = [[float((rel[w, 8:-8, k] - rel[w, 8:-8, k - 1]).mean()) for k in range(1, 8)] for w in range(2)]
        off = [int(rel[1, j, 3] - rel[0, j, 3]) for j in range(0, n_kv, 8)]
        row = dict(impl=impl, stagger=st, clocks_per_block=per_block, phase_means=ph, tile1_minus_tile0_at_max=off,
                   first_blocks=rel[:, :12, :].tolist(), mid_blocks=rel[:, 40:46, :].tolist())
        if NW == 4:   # MMA warps: 0 loop top, 1 s_free seen, 2 S(j+2) issued, 3 V ready, 4 p_ready seen, 5 PV issued
            row["mma_phase_means"] = [[float((rel[w, 8:-8, k] - rel[w, 8:-8, k - 1]).mean()) for k in range(1, 6)] for w in (2, 3)]
        res.append(row)
        print(json.dumps({k: v for k, v in row.items() if k not in ("first_blocks", "mid_blocks")}), flush=True)
native.set_tuning("fmha_stagger", 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "attn_trace.json"), "w"))
