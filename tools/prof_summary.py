"""Turns gpurun_out/<round>_launches.csv (ncu launch list) and profiles/<round>_*_ncu_raw.csv into the markdown summaries."""
import collections, csv, re, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01b"
rows = list(csv.reader(open(f"gpurun_out/{R}_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, gi, bi = (hdr.index(n) for n in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("unnamed>::", "")
    a = agg.setdefault((name, r[gi], r[bi]), [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", "")) / 1000.0
skip = ("fill_normal", "permute_conv_weight", "quantize_kernel")
tot = sum(v[1] for k, v in agg.items() if not any(s in k[0] for s in skip))
with open(f"profiles/{R}_launches_dit_step_summary.md", "w") as f:
    f.write(f"# ncu launch list, `LTX_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg5` (B200, first 2600 launches)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 2600`; per-launch times are cold-cache and serialised -- compare SHARES.\n"
            "Weight-init kernels (`fill_normal_kernel`, `permute_conv_weight_kernel`) are excluded. The window covers model init, the text-cache build and ~4 denoise steps.\n\n"
            "| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if any(s in k[0] for s in skip) or v[1] / tot < 0.0005: continue
        f.write(f"| `{k[0]}` | {k[1]} | {k[2]} | {v[0]} | {v[1]:.1f} | {v[1]/v[0]:.1f} | {100*v[1]/tot:.1f}% |\n")
    cls = collections.Counter()
    for k, v in agg.items():
        if any(s in k[0] for s in skip): continue
        c = "gemm" if "gemm" in k[0] else "attention" if "attention" in k[0] else "rows" if ("norm" in k[0]) else "other"
        cls[c] += v[1]
    f.write("\nClass shares under ncu: " + ", ".join(f"{c} {100*t/tot:.1f}%" for c, t in cls.most_common()) + "\n")
print(open(f"profiles/{R}_launches_dit_step_summary.md").read())
