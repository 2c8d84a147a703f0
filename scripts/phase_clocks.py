"""Development: per-phase clock64 stamps of CTA 0 of k_hs_rows (library built with -DDBMM_PHASE_TIMERS)."""
import ctypes as C, os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DBMM_GRAPH", "0"); os.environ.setdefault("DBMM_TAIL", "serial"); os.environ.setdefault("N", "8192"); os.environ.setdefault("EPOCHS", "1")
exec(open(os.path.join(os.path.dirname(__file__), "train_only.py")).read())
from dbmm import _lib
lib = _lib.load()
buf = (C.c_longlong * 32)()
lib.dbmm_debug_phase_clocks.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.dbmm_debug_phase_clocks(buf))
v = list(buf)
print("hs_rows phases (cycles since tick 0):", [v[i] - v[0] for i in range(11)])
print("hs_w2 phases (cycles since tick 0):", [v[16 + i] - v[16] for i in range(8)])
