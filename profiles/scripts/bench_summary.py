import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f"{f}"))
        r=d["roofline"]
        print(f, "cold",round(d["value"]), "warm",round(d["warm"]["value"]), "e2e",round(d["e2e"]["value"]), "cold_parts",[round(v*1e3,1) for v in [r["launch_ms_all"]["primal_or_kernel_projection"]]+r["launch_ms_all"]["sweeps_in_launch_order"]+r["launch_ms_all"]["dual_kernels_branching_chain_leaves"]], "warm_parts",[round(v*1e3,1) for v in r["launch_ms_all_warm"]["dual_kernels_branching_chain_leaves"]], "frac",round(r["frac"],3))
    except Exception as e: print(f, "ERR", e)
