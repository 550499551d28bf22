"""profiles/executed.json from ncu captures: executed FP32 flops and MUFU.RCP operations per ALGORITHMIC flop of the same launch.

    python profiles/scripts/executed_from_hist.py key=histogram.txt:bench_line.json [...]

histogram.txt = profiles/scripts/sass_hist.py of the `ncu --set full --import-source on` source page (executed warp
instructions per SASS opcode); bench_line.json = the bench.py line of the same command without the profiler (its
roofline.achieved x ms_per_step = the algorithmic flops of one launch, a timing-independent product).
Warp-level count: every executed FP32 instruction is charged for all 32 lanes (FFMA2 = 64 FMAs = 128 flops), so the
fraction it yields is pipe occupancy - at T = 50 only 25 of the 32 lanes carry time samples.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out_path = os.path.join(ROOT, "profiles", "executed.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
for spec in sys.argv[1:]:
    key, rest = spec.split("=")
    hist, line = rest.split(":")
    tot = {}
    for l in open(hist):
        m = re.match(r"(\S+)\s+(\d+)\s", l)
        if m:
            tot[m.group(1)] = int(m.group(2))
    fp32 = 128 * tot.get("FFMA2", 0) + 64 * (tot.get("FMUL2", 0) + tot.get("FADD2", 0)) + 64 * tot.get("FFMA", 0) \
        + 32 * (tot.get("FADD", 0) + tot.get("FMUL", 0))
    mufu = 32 * sum(v for k, v in tot.items() if k.startswith("MUFU.RCP"))
    j = json.loads([l for l in open(line) if l.startswith("{")][-1])
    alg = j["roofline"]["achieved"] * 1e12 * j["ms_per_step"] * 1e-3
    out[key] = {"fp32_flop_per_alg_flop": fp32 / alg, "mufu_per_alg_flop": mufu / alg, "executed_fp32_flop": fp32,
                "mufu_rcp": mufu, "algorithmic_flop": alg, "warp_instructions": sum(tot.values()),
                "source": f"profiles/{os.path.basename(hist)} + profiles/{os.path.basename(line)}"}
    print(key, out[key])
json.dump(out, open(out_path, "w"), indent=1)
