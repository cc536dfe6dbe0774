"""Per-kernel shares of the last two wavefronts (= the two timed steps of `bench.py --quick --steps 2`) in an
`ncu --metrics gpu__time_duration.sum --csv` launch list, next to the CUDA-event stage times of the same command run without ncu:
    python scripts/ncu_launch_shares.py launches.csv bench_quick.json > profiles/rNN_ncu_launches_bench.md"""
import collections, csv, json, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[data[0][iu].lower()]
names = [re.sub(r"\(.*", "", d[ik]).replace("void ", "").replace("jpbrt::", "") for d in data]
gen = [i for i, n in enumerate(names) if n.startswith("k_generate")]
agg = collections.OrderedDict()
for n, d in zip(names[gen[-2]:], data[gen[-2]:]):
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(d[iv].replace(",", "")) * scale
ours = {n: a for n, a in agg.items() if n.startswith("k_") and not n.startswith(("k_gather", "k_l2_stream", "k_ffma"))}
tot = sum(a[1] for a in ours.values())
q = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
st = q["stages_ms_per_step"]
print(f"# ncu launch list of the bench command (round 2, final kernels)\n")
print("Command (run first without ncu, exit 0; then under ncu in the same gpurun call):")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file … python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline`\n")
print(f"{len(data)} launches captured (the run's own ceiling micro-benchmark `peaks_l2` included: `k_gather`, `k_l2_stream`, `k_ffma` are not the renderer's).")
print(f"Without ncu the same command printed {q['value']:.0f} Msamples/s, {q['ms_per_step']:.2f} ms per step.\n")
print("The two timed steps (last two wavefronts), ncu per-launch durations (cold-cache, serialised) summed per kernel:\n")
print("| kernel | launches | ncu sum (ms, 2 steps) | share | CUDA events without ncu (ms per step) |\n|---|---|---|---|---|")
ev = {"k_extend": f"{st['extend']:.2f}", "k_connect": f"{st['connect']:.2f}", "k_generate": f"{st['generate']:.2f}", "k_logic": f"shade stage = logic + material kernels: {st['shade']:.2f}"}
for n, a in ours.items():
    key = next((k for k in ev if n.startswith(k)), None)
    print(f"| `{n}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % | {ev.get(key, '') if key else ''} |")
print(f"| total | {sum(a[0] for a in ours.values())} | {tot:.2f} ({tot / 2:.2f} per step) | | {q['ms_per_step']:.2f} (event-bracketed step) |")
ke = next(a for n, a in ours.items() if n.startswith("k_extend"))
print(f"\nThe dominant kernel's share ({100 * ke[1] / tot:.1f} %) and its time per step ({ke[1] / 2:.2f} ms under ncu, {st['extend']:.2f} ms by CUDA events) agree.")
