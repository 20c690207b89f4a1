"""Phase-level timing of the persistent token kernel from %globaltimer stamps.
usage: python trace_token.py [workload=7b]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module("rwkv-cpp-accelerated_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "7b"
L, E = bench.SHAPES[workload]
eng = pkg.Engine(bench.model_path(workload, pkg))
eng.set_option("trace", 1)
tok = bench.SEED_TOKEN
for _ in range(4):
    tok = eng.forward_greedy(tok)
tr = eng.read_trace().astype(np.int64)
n = int((tr[0] > 0).sum())
tr = tr[:, :n]
t0 = tr[:, 0].min()
tr = tr - t0
print("stamps per CTA", n, "token time us", (tr[:, -1].max()) / 1e3)
# stamp layout: 0 start | 1 E(embed) 2 B | per layer: [slice: E B][kvr: G C E B][out: G C E B][ffn slice: E B][rk: G C E B][v: G C E B] | head...
names = []
names += ["embed.epi", "embed.bar"]
for l in range(L):
    for ph in ("kvr", "out", "rk", "v"):
        if ph in ("kvr", "rk"):
            names += [("att_ln" if ph == "kvr" else "ffn_ln") + ".epi", ("att_ln" if ph == "kvr" else "ffn_ln") + ".bar"]
        names += [ph + ".g_loads", ph + ".g_reduce", ph + ".g_quant", ph + ".g_sync", ph + ".gemv", ph + ".epi", ph + ".bar"]
names += ["head_ln.epi", "head_ln.bar", "head.g_loads", "head.g_reduce", "head.g_quant", "head.g_sync", "head.gemv"]
d = np.diff(tr, axis=1)  # [cta, n-1]
agg = {}
for i, nm in enumerate(names[:d.shape[1]]):
    agg.setdefault(nm, []).append(d[:, i])
print("%-12s %8s %8s %8s %8s   (us; mean over CTAs and layers, then min/max of per-CTA means)" % ("segment", "mean", "median", "min", "max"))
tot = 0
for nm, lst in agg.items():
    a = np.stack(lst, 1) / 1e3  # [cta, layers]
    per_layer = a.mean()
    tot += a.mean(0).sum()
    print("%-12s %8.2f %8.2f %8.2f %8.2f   x%d" % (nm, per_layer, np.median(a), a.mean(1).min(), a.mean(1).max(), a.shape[1]))
print("sum of means (us):", tot)
# arrival skew at barriers: spread of the stamp just before each barrier
bar_idx = [i for i, nm in enumerate(names[:d.shape[1]]) if nm.endswith(".bar")]
skew = [(tr[:, i].max() - tr[:, i].min()) / 1e3 for i in bar_idx]
print("arrival skew at barriers (us): mean %.2f median %.2f max %.2f" % (np.mean(skew), np.median(skew), np.max(skew)))
lat = [(tr[:, i + 1].min() - tr[:, i].max()) / 1e3 for i in bar_idx]
print("barrier release latency after LAST arrival (us): mean %.2f median %.2f" % (np.mean(lat), np.median(lat)))

# ---- tile-level: how far ahead of the consumers does the producer run? -------------------------
tt = eng.read_tile_trace().astype(np.int64)
n_cta = 148
for cta in (0, 77):
    issue, ready = tt[0, cta], tt[1, cta]
    n_t = int((issue > 0).sum())
    issue, ready = (issue[:n_t] - t0) / 1e3, (ready[:n_t] - t0) / 1e3
    d_t = np.diff(ready)
    inside = d_t[d_t < 2.0]  # consecutive tiles of one streaming phase (gaps are phase boundaries)
    print("CTA %d: tiles %d; tile period inside a phase (us): median %.3f mean %.3f p90 %.3f" % (
        cta, n_t, np.median(inside), inside.mean(), np.percentile(inside, 90)))
    per_layer = 48 if workload == "7b" else None
    if per_layer:
        base = 2 * per_layer
        print("layer-2 tiles of CTA %d: idx issue(us) started(us) lead(us)" % cta)
        for i in range(base, min(base + per_layer, n_t)):
            print("%4d %9.2f %9.2f %7.2f" % (i - base, issue[i], ready[i], ready[i] - issue[i]))
print("layer-2 stamps (CTA 0, us):", np.round(tr[0, 3 + 2 * 32:3 + 3 * 32] / 1e3, 2))
