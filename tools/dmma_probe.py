"""north_star (a): FP64 DMMA vs the FMA pipe on the plate contraction -- prints fs_bench_contraction's numbers
(ncu: `ncu --set full -k regex:k_contract python tools/dmma_probe.py`)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_shell_b200 as fsb

s = fsb.FemShell(device=0)
r = s.bench_contraction(n_elem=1 << 20, reps=int(os.environ.get("REPS", "20")))
r["fp64_fma_peak_tflops"] = s.bench_fp64_peak()
r["note"] = "useful = 2*12^3 flops per element; DMMA executes 12 m8n8k4 per element (56 % of its MACs useful)"
print(json.dumps(r))
