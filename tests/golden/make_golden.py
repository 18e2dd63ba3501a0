"""Extracts the reference's only pinned PCG results for this path -- the DICPCG lines of
cases/steckler/original/linux64/log.fireFoam (listed by cases/steckler/testFiles:1) -- into
steckler_log.json.  Run in the build container (needs /root/reference); the JSON is committed
because /root/reference does not exist on the GPU box."""
import json
import os
import re

LOG = "/root/reference/cases/steckler/original/linux64/log.fireFoam"
pat = re.compile(r"^(\w+):\s+Solving for (\w+), Initial residual = (\S+), Final residual = (\S+), "
                 r"No Iterations (\d+)")
out = {"source": "cases/steckler/original/linux64/log.fireFoam", "ph_rgh": [], "p_rgh": [],
       "variation": []}
with open(LOG) as f:
    for n, line in enumerate(f, 1):
        m = pat.match(line)
        if m and m.group(2) in ("ph_rgh", "p_rgh"):
            out[m.group(2)].append({"line": n, "solver": m.group(1), "initial": float(m.group(3)),
                                    "final": float(m.group(4)), "iters": int(m.group(5))})
        m = re.match(r"^Hydrostatic pressure variation of internalField \(gMax-gMin\): (\S+)", line)
        if m:
            out["variation"].append({"line": n, "value": float(m.group(1))})
with open(os.path.join(os.path.dirname(__file__), "steckler_log.json"), "w") as f:
    json.dump(out, f, indent=1)
print({k: len(v) for k, v in out.items() if isinstance(v, list)})
