import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, "value %.4g e2e %.4g frac %.4f launch_ms %.3f ms/step %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
