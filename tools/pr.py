import json,sys
for f in sys.argv[1:]:
    try:
        line=[l for l in open(f).read().splitlines() if l.startswith("{")][-1]
        d=json.loads(line); print(f, "value %.4g e2e %.4g frac %.4f launch_ms %.3f ms/step %.3f e2e_ms %.3f hops %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["stats"]["hops_per_substep"]), d["stats"])
    except Exception as e: print(f, "ERR", e)
