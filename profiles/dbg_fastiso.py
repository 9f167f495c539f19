import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, ctypes as C
import util
from covid19_demography_b200 import engine
name = sys.argv[1] if len(sys.argv) > 1 else "italy_fastiso"
c = util.Case(name)
params = c.no_tracing_params()
tup, ex, home, init = c.keyed_oracle("native", params)
want = util.pack_history(tup[:9])
for launch in ("persistent", "per_day"):
    eng = c.gpu_engine(engine, params, "native", launch=launch)
    eng.run([c.seed], [home], [init])
    packed = eng.history("packed")
    bad = np.argwhere(packed != want)
    ho = np.zeros(8, dtype=np.int32)
    eng.lib.seir_get_handover.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    eng.lib.seir_get_handover(eng._h, ho.ctypes.data, 8)
    tal = eng.tallies(0)
    act = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))
    print(launch, "diffs", len(bad), bad[:3].tolist(), "handover", ho.tolist(), "counters", eng.counters())
    print("   active per day", act.tolist())
    eng.close()
