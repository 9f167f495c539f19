import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle')); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, util
from covid19_demography_b200 import engine
for name, trac in (("italy_default", True), ("italy_policy", False)):
    c = util.Case(name)
    params = dict(c.params) if trac else c.no_tracing_params()
    params["T"] = 45.0 if trac else 30.0
    tup, ex, home, init = c.keyed_oracle("native", params)
    eng = c.gpu_engine(engine, params, "native")
    eng.run([c.seed], [home], [init])
    ok = np.array_equal(eng.history("packed"), util.pack_history(tup[:9])) and np.array_equal(eng.tallies(0), util.tallies_from_history(tup[:9], c.age, int(params["T"])))
    print(name, "tracing" if trac else "no tracing", "bit-exact" if ok else "MISMATCH", flush=True)
    eng.close()
