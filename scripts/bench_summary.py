import json, sys
for fn in sys.argv[1:]:
    d = json.load(open(fn))
    print(fn, "ms/step %.4f" % d["ms_per_step"], "vol/s %.1f" % d["value"], "roofline.frac", d["roofline"]["frac"], "e2e ms %.2f" % d["e2e"]["ms_per_step"], d["clocks"])
    for k in d["kernels"]:
        print("   %-45s %.4f ms  %7.1f GB/s  %.3f" % (k["name"][:45], k["ms"], k["gbs"], k["frac_of_peak"]))
