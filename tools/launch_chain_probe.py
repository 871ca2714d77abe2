"""kernel-to-kernel latency of a graph-replayed chain of short dependent kernels, plain vs programmatic dependent launch"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fem_shell_b200 as fsb
s = fsb.FemShell()
out = {}
for n in (1024, 65536, 1 << 20, 6 << 20):
    out[str(n)] = s.bench_launch_chain(links=200, n=n, reps=20)
print(json.dumps(out))
