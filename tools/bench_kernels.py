import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.3f value %.1f clocks %s"%(d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"]))
for k in d["kernels"]:
    print("  %-28s %7.4f ms  %s TF %s GB/s x%d"%(k["name"],k["ms_per_step"],k["tflops"],k["gbs"],k["launches"]))
