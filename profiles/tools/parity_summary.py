"""gpurun_out/parity_report.jsonl (written by the GPU tests through conftest.report) -> profiles/rNN_parity.md:
the achieved errors behind the pass/fail bars.   python profiles/tools/parity_summary.py > profiles/r02_parity.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")
rows = [json.loads(l) for l in open(path) if l.strip()]
print("# Achieved parity figures (GPU suite on B200; bars: image max-abs 1e-4, gradients rel-L2 1e-3, integers bit-exact)\n")
print(f"{len(rows)} records from `{os.path.relpath(path, ROOT)}` of the final full run of the round.\n")
by = {}
for r in rows:
    by.setdefault(r["test"] if isinstance(r["test"], str) else "misc", []).append(r)
for test, rs in by.items():
    keys = [k for k in rs[0] if k != "test"]
    num = [k for k in keys if all(isinstance(r.get(k), (int, float)) and not isinstance(r.get(k), bool) for r in rs)]
    tags = [k for k in keys if k not in num]
    print(f"## {test} ({len(rs)} cases)\n")
    if len(rs) <= 12:
        print("| " + " | ".join(tags + num) + " |\n|" + "---|" * (len(tags) + len(num)))
        for r in rs:
            print("| " + " | ".join([str(r.get(k)) for k in tags] + [f"{r.get(k):.3g}" for k in num]) + " |")
    else:
        print("| figure | max over cases | min |\n|---|---|---|")
        for k in num:
            v = [r[k] for r in rs]
            print(f"| {k} | {max(v):.3g} | {min(v):.3g} |")
    print()
