import sys; sys.path.insert(0, '.')
import numpy as np
from sipnet_b200 import _abi as A, api, synth
sites, P, ms, flags = synth.config_c3(nsites=3, members_per_site=45, nyears=1)
for math in (A.MATH_FAST, A.MATH_VALIDATION):
    ens = api.Ensemble(sites, P, ms, flags, outputs=A.OUT_FULL | A.OUT_EVENTS | A.OUT_LOGLIK | A.OUT_MOMENTS | A.OUT_QUANTILES, math=math, summary_cols=[11], quantiles=[0.5], max_event_records=256)
    ens.run(0, 300); ens.run(300, 730)
    o = ens.output(); m = ens.mean(); q = ens.quantiles(); ens.close()
print("done", np.isfinite(o).mean())
