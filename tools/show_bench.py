import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("headline", round(d["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "e2e", round(d["e2e"]["value"],1))
for k,v in d["suite"].items(): print(k, round(v.get("value",0),1), v.get("unit"), "cpu:", v.get("cpu_baseline",{}).get("value"), v.get("error",""))
