"""Per-instruction active-lane profile of the traversal kernels from an `ncu --set full --import-source on` report:
python scripts/ncu_lane_profile.py prof.ncu-rep [kernel-regex]   (needs ncu on PATH)

For every captured launch of the kernel: warp instructions, thread instructions, average active lanes, and for the
load / vote / branch instructions (the skeleton of the refill -> node phase -> leaf phase loop) how often they ran and
with how many lanes.  This is the view that showed the node phase waiting for its slowest ray."""
import csv, io, subprocess, sys

rep = sys.argv[1]
regex = sys.argv[2] if len(sys.argv) > 2 else "k_extend|k_connect"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
blocks, cur, name = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    elif r and r[0] == "Address":
        cur = {"name": name, "header": r, "data": []}
        blocks.append(cur)
    elif r and r[0].startswith("0x") and cur is not None:
        cur["data"].append(r)
seen = set()
for b in blocks:
    h, data = b["header"], b["data"]
    iI, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
    tot_i = sum(int(r[iI]) for r in data)
    tot_t = sum(int(r[iT]) for r in data)
    key = (b["name"], tot_i)
    if key in seen or tot_i == 0:
        continue  # ncu prints the SASS view twice per launch
    seen.add(key)
    print(f"== {b['name'][:60]}  warp instr {tot_i}  thread instr {tot_t}  avg lanes {tot_t / tot_i:.2f}")
    for i, r in enumerate(data):
        src = r[iS].strip()
        n = int(r[iI])
        if n and any(k in src for k in ("LDG", "VOTE", "ATOMG", "STG", "STL", "LDL", "RED")):
            print(f"  {i:4d} {src[:64]:64s} execs {n:10d} lanes {int(r[iT]) / n:5.1f}")
