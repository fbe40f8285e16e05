"""Profiling target: `--iters` full ST-pipeline calls at the bench workload (no oracle compute).

Default: just runs (ncu target).  With --table every C-ABI call is bracketed with CUDA events (in situ, warm
caches) and a per-(op, shape) time table of the LAST iteration is printed."""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops  # noqa: E402
import benchdata as O  # noqa: E402  (seeded weight / input generator, no model math)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--table", action="store_true")
a = ap.parse_args()
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
m = S.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
m.load_state_dict(O.make_state_dict(man["wnet_em"], seed=1), strict=True)
inf = S.SwinWNetInference(m, "cuda:0")
x = O.synthetic_diffractions(min(a.batch, 4), seed=1, two_channel=False)
x = x.repeat((a.batch + 3) // 4, 1, 1, 1)[:a.batch].to("cuda:0")

records = []
if a.table:
    def wrap(name):
        fn = getattr(ops, name)

        def timed(*args, **kw):
            if name == "rowgemm":
                key = f"a{kw['a_mode']} e{kw['e_mode']} M={kw['M']} K={kw['K']} N={kw['nchunks'] * kw['n_valid']}"
            elif name == "mlp":
                key = f"M={args[2]} C={args[3]}"
            elif name == "window_attention":
                key = f"B={args[4]} {args[5]}x{args[6]} C={args[7]} nH={args[8]}"
            elif name == "swin_block_fused":
                key = f"B={args[2]} {args[3]}x{args[4]} C={args[5]} nH={args[6]}"
            elif name == "swin_block_warp":
                key = f"B={args[2]} {args[3]}x{args[4]} C={args[5]} depth={args[10] if len(args) > 10 else kw.get('depth', 1)}"
            elif name == "swin_block_small":
                key = f"B={args[2]} {args[3]}x{args[4]} C={args[5]}"
            elif name == "cross_attention":
                key = f"Lq={args[4]} Lk={args[5]} C={args[6]}"
            else:
                key = ""
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            records.append((name, key, e0, e1))
            return r
        setattr(ops, name, timed)
    for n in ("rowgemm", "mlp", "window_attention", "swin_block_small", "swin_block_fused", "swin_block_warp", "cross_attention", "patch_embed", "seg_head",
              "recon_head", "copy_cols", "sigmoid_mask", "normalize"):
        wrap(n)

for it in range(a.iters):
    records.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    inf(x)
    e1.record()
    torch.cuda.synchronize()
print("ok", S.ops.LAUNCH_COUNT, f"last iteration {e0.elapsed_time(e1):.2f} ms")
if a.table:
    agg = collections.OrderedDict()
    for name, key, s, e in records:
        k = (name, key)
        t = agg.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += s.elapsed_time(e)
    total = sum(v[1] for v in agg.values())
    print(f"sum of op times {total:.2f} ms over {len(records)} calls")
    for (name, key), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:8.3f} ms {100 * ms / total:5.1f}%  n={n:3d} avg {ms / n:7.3f}  {name:18s} {key}")
