"""Per-GEMM-shape efficiency from an ncu launch list of one headline step: python scripts/gemm_shapes.py launches.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
g = []
for r in rows[hi + 1:]:
    if len(r) > vi and 'gemm_split3' in r[ki]:
        g.append(float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[ui], 1))
N, T = 6400, 16; Rp = N * T; Rv = 64 * T; Ra = Rp + Rv; Mc = Rp * 4
seq = [("corr_fc1", Mc, 384, 2401), ("corr_fc2", Mc, 256, 384), ("in_tr", Rp, 384, 1110)]
for i in range(3):
    seq += [("t.qkv", Ra, 1152, 384), ("t.out", Ra, 384, 384), ("t.fc1", Ra, 1536, 384), ("t.fc2", Ra, 384, 1536)]
    seq += [("v2p.q", Rv, 384, 384), ("v2p.kv", Rp, 768, 384), ("v2p.out", Rv, 384, 384), ("v2p.fc1", Rv, 1536, 384), ("v2p.fc2", Rv, 384, 1536)]
    seq += [("vs.qkv", Rv, 1152, 384), ("vs.out", Rv, 384, 384), ("vs.fc1", Rv, 1536, 384), ("vs.fc2", Rv, 384, 1536)]
    seq += [("p2v.q", Rp, 384, 384), ("p2v.kv", Rv, 768, 384), ("p2v.out", Rp, 384, 384), ("p2v.fc1", Rp, 1536, 384), ("p2v.fc2", Rp, 384, 1536)]
it = g[57:114]
agg = {}
for (name, M, Nn, K), us in zip(seq, it):
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += 2 * M * Nn * K
tot = sum(a[1] for a in agg.values()); small = 0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] < 100: small += a[1]; continue
    print(f"{k:9s} n={a[0]} {a[1]:9.1f} us {100*a[1]/tot:5.1f}%  {a[2]/a[1]/1e6:7.1f} TFLOP/s (x3 = {3*a[2]/a[1]/1e6:6.0f} bf16)")
print("small virtual-token GEMMs us", round(small, 1), "total us", round(tot, 1))
