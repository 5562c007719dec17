"""Compare tools/variant_check.py outputs: python tools/pick_variant.py base.log var1.log ...  -> prints the name
of the fastest build (bench workload) whose unweighted counts are all identical to the first log's."""
import sys, json

def load(p):
    for line in open(p):
        if line.startswith("JSON "):
            return json.loads(line[5:])
    return None

logs = [(p, load(p)) for p in sys.argv[1:]]
base = logs[0][1]
best, best_ms = sys.argv[1], base["c2_f_smu_fma:00"]["ms"] if base else 1e30
for p, d in logs[1:]:
    if not d or not base:
        print(f"# {p}: no result", file=sys.stderr); continue
    ok = True
    for k, v in base.items():
        if k == "lib": continue
        if v["digest"] is not None:
            if d[k]["digest"] != v["digest"]: ok = False; print(f"# {p}: {k} counts differ", file=sys.stderr)
        else:
            rel = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(d[k]["wsums"], v["wsums"]))
            if rel > 1e-12: ok = False; print(f"# {p}: {k} weighted sums differ by {rel:.2e}", file=sys.stderr)
    ms = d["c2_f_smu_fma:00"]["ms"]
    print(f"# {p}: parity {'ok' if ok else 'FAIL'}; c2 {ms} ms (base {base['c2_f_smu_fma:00']['ms']} ms)", file=sys.stderr)
    if ok and ms < best_ms * 0.995:
        best, best_ms = p, ms
print(best)
