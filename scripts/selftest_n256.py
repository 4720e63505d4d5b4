"""One self-test pass at the bench batch (N=256): every conv through SIMT and tcgen05 kernels (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from test_gpu_tc_selftest import run_selftest
rows = run_selftest(int(os.environ.get("N", "256")))
bad = [(n, r) for n, r in rows if any(x is not None and x > 4e-3 for x in r)]
print("rows", len(rows), "bad", bad)
