import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "ade/fde", d["config"].get("ade_px"), d["config"].get("fde_px"), "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "clk", d["clocks"])
    for g in d["roofline"]["by_group"][:16]:
        print("   ", g)
