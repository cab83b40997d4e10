import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    c = d["cg"]; print(round(d["value"],2), c["its"], c["reason"], c["rnorm_rel"], round(c["time_s"],2), d["clocks"])
