#!/usr/bin/env python
"""prints the few numbers of a bench.py JSON line that a GPU-box visit is usually about"""
import json
import sys

for ln in open(sys.argv[1]):
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    m = d.get("metrics") or {}
    print("impl %s N=%s value %.4g e2e %.4g ms/step %.3f" % (d.get("impl", "ours"), d.get("n_gpus"), d["value"], d["e2e"]["value"], d["ms_per_step"]))
    if "roofline" in d:
        r = d["roofline"]
        print("  roofline %s %.1f GB/s frac %.3f ms %.4f share %.3f" % (r["kernel"], r["achieved"], r["frac"], r["ms_per_launch"], r["share_of_step"]))
    print("  assemble_ms %s el/s %.4g" % (m.get("assemble_ms"), m.get("elements_assembled_per_s", 0)))
    if "assembly_roofline" in d:
        a = d["assembly_roofline"]
        print("  assembly fp64 frac %.3f hbm frac %.3f (fp64 peak %.1f)" % (a["fp64"]["frac"], a["hbm"]["frac"], a["fp64"]["peak"]))
    for k, v in (m.get("time_to_solution") or {}).items():
        print("  tts %s: %.3f s, %d its, conv %s, oracle_residual %s floor %s vs_ml %s" % (k, v["seconds"], v["iterations"], v["converged"], v.get("oracle_residual"), v.get("oracle_residual_eval_floor"), v.get("rel_l2_vs_multilevel")))
    for k in ("time_to_solution_bounded", "time_to_first_solution", "dist_parity"):
        if m.get(k):
            print("  %s: %s" % (k, json.dumps(m[k])[:300]))
    if m.get("c3"):
        c = m["c3"]
        print("  c3: tts %.3f s (asm %.4f + solve %.3f), %d its, first %.2f s, set_mesh %.2f s, el/s %.4g, spmv frac %.3f, defl %.5f, conv %s"
              % (c["time_to_solution_s"], c["assemble_s"], c["solve_s"], c["iterations"], c["time_to_first_solution_s"], c["set_mesh_s"],
                 c["elements_assembled_per_s"], c["spmv_frac_of_hbm_peak"], c["deflection_ratio"], c["converged"]))
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print("  cpu %.4g (%d cores) el/s %.4g  -> ratio %.1f" % (c["value"], c["cores"], c.get("elements_per_s", 0), d["value"] / c["value"]))
    print("  clocks %s" % json.dumps(d.get("clocks")))
