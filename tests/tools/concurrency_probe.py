"""Runs the mixed-family concurrency test body with a watchdog that dumps every Python thread's stack after 40 s."""
import faulthandler, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
faulthandler.dump_traceback_later(40, exit=True)
import importlib.util, types
import numpy as np
from algo_dsp_b200 import conv
from oracle import oracle
spec = importlib.util.spec_from_file_location("rb", os.path.join(ROOT, "tests", "test_gpu_robustness.py"))
m = importlib.util.module_from_spec(spec)
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec.loader.exec_module(m)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    m.test_concurrent_mixed_families_on_one_context(conv, oracle)
    print("round", i, "ok", flush=True)
