"""Print the headline numbers of a bench.py JSON line (last line starting with '{' in the given file)."""
import json, sys
b = json.loads([x for x in open(sys.argv[1]) if x.startswith("{")][-1])
print(b["ms_per_step"], b.get("phases_ms"), "mufu frac", b["roofline"]["frac"], "e2e", (b.get("e2e") or {}).get("ms_per_step"),
      "iter", (b.get("iteration") or {}).get("ms"))
