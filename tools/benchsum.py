"""One-line summary of bench.py JSON lines:  python tools/benchsum.py <label>=<file> ..."""
import json
import sys

for a in sys.argv[1:]:
    label, _, path = a.partition("=")
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        e = d["e2e"]
        print(label, d["config"]["kmers"], "resident", round(d["ms_per_step"], 1),
              {x: round(k[x]["ms"], 2) for x in k}, "e2e", round(e["ms_per_step"], 1),
              "e2e(class strings)", round(e.get("class_strings", {}).get("ms_per_step", 0), 1),
              "flips", d["parity_sample"]["flips"], "bad", d["reads_with_errors"],
              "cli", (d.get("cli") or {}).get("wall_s"))
    except Exception as ex:                                     # noqa: BLE001
        print(label, "failed:", ex)
