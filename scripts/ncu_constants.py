"""Distil an `ncu --set full` report of scripts/profile_render.py into the per-ray constants bench.py quotes in its
`roofline` (traffic, DRAM fraction, issue-slot x lane efficiency):

    python scripts/ncu_constants.py <config> <report.ncu-rep> <spp> [existing.json]  ->  JSON on stdout

Rays per captured k_extend launch come from the report itself: thread-level executions of the hit-record store."""
import csv, io, json, subprocess, sys

cfg, rep, spp = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else {}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
def col(r, name, default=0.0):
    try:
        return float(r[h.index(name)].replace(",", ""))
    except (ValueError, IndexError):
        return default
unit = {n: rows[1][i] for i, n in enumerate(h)}
def to_bytes(v, name):
    u = unit.get(name, "byte").lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
def to_ms(v, name):
    u = unit.get(name, "ms").lower()
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}.get(u, 1)
# rays per k_extend launch: thread executions of STG.E.EF.64 (the hit store) from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_extend"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address":
        cur = {"h": r, "rows": []}; blocks.append(cur)
    elif r and r[0].startswith("0x") and cur is not None:
        cur["rows"].append(r)
rays, seen = [], set()
for b in blocks:
    iT, iS, iI = b["h"].index("Thread Instructions Executed"), b["h"].index("Source"), b["h"].index("Instructions Executed")
    tot = sum(int(r[iI]) for r in b["rows"])
    if tot in seen or tot == 0:
        continue
    seen.add(tot)
    rays.append(sum(int(r[iT]) for r in b["rows"] if "STG.E" in r[iS] and ".64" in r[iS]))
agg = {}
k = 0
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    kind = "k_extend" if "k_extend" in name else "k_connect" if "k_connect" in name else "k_shade" if "k_shade" in name else None
    if kind is None:
        continue
    a = agg.setdefault(kind, {"ms": 0.0, "dram": 0.0, "l2": 0.0, "winst": 0.0, "tinst": 0.0, "cycles": 0.0, "launches": 0, "rays": 0.0})
    a["ms"] += to_ms(col(r, "gpu__time_duration.sum"), "gpu__time_duration.sum")
    a["dram"] += to_bytes(col(r, "dram__bytes_read.sum"), "dram__bytes_read.sum") + to_bytes(col(r, "dram__bytes_write.sum"), "dram__bytes_write.sum")
    a["l2"] += 32.0 * col(r, "lts__t_sectors.sum")  # 32-byte sectors looked up in L2
    a["winst"] += col(r, "sm__inst_executed.sum") or col(r, "smsp__inst_executed.sum")
    a["tinst"] += (col(r, "sm__inst_executed.sum") or col(r, "smsp__inst_executed.sum")) * col(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    a["cycles"] += col(r, "sm__cycles_active.avg")
    a["ipc_last"] = col(r, "sm__inst_executed.avg.per_cycle_active")
    a["launches"] += 1
    if kind == "k_extend" and k < len(rays):
        a["rays"] += rays[k]; k += 1
res = {"source": f"profiles: ncu --set full --clock-control none of `scripts/profile_render.py {cfg} {spp}` ({rep.split('/')[-1]}), launches of bounces 0-1", "hbm_peak_gbs": 6545.3}
for kind, a in agg.items():
    lanes = a["tinst"] / a["winst"] if a["winst"] else None
    ipc = a["winst"] / a["cycles"] / 148 if a["cycles"] else None  # sm__inst_executed.sum over all SMs / (avg active cycles x SMs)
    e = {"launches_captured": a["launches"], "ms": a["ms"], "dram_bytes": a["dram"], "l2_bytes": a["l2"],
         "dram_frac": a["dram"] / (a["ms"] * 1e-3) / 6545.3e9 if a["ms"] else None,
         "l2_gbs": a["l2"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] else None,
         "ipc": ipc, "lanes_per_inst": lanes, "issue_lane_eff": (ipc / 4.0) * (lanes / 32.0) if ipc and lanes else None}
    if kind == "k_extend" and a["rays"]:
        e.update({"rays": a["rays"], "dram_bytes_per_ray": a["dram"] / a["rays"], "l2_bytes_per_ray": a["l2"] / a["rays"],
                  "thread_inst_per_ray": a["tinst"] / a["rays"], "warp_inst_per_ray": a["winst"] / a["rays"]})
    res[kind] = e
out[cfg] = res
print(json.dumps(out, indent=1))
