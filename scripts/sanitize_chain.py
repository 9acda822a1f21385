"""Short chains through every step-kernel variant, sized so that each warp of the TMA-staged kernel re-arms its
pipeline stages several times.  Run under compute-sanitizer (one tool per call):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python scripts/sanitize_chain.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python scripts/sanitize_chain.py

The chains are also compared with each other (TMA-staged == direct == generic) so a silent corruption shows up even
where the tool reports nothing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmbayes_b200 import api  # noqa: E402
from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort  # noqa: E402

N_BIG = int(os.environ.get("SAN_N", "8000"))


def chain(c, variant, ms=False, iters=12):
    if variant:
        os.environ["JMB_KERNEL"] = variant
    else:
        os.environ.pop("JMB_KERNEL", None)
    ctl = dict(c["control"]); ctl.update(n_burnin=iters - 4, n_iter=4, n_block=4, n_thin=4)
    with api.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"], multiState=ms) as fit:
        name = fit.layout_info()["kernel"]
        fit.set_draw(c["Data"])
        r = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    return name, r


for cfg, n in (("config4", N_BIG), ("config3", N_BIG), ("config2", 3000)):
    c = make_cohort(cfg, n=n)
    ref = None
    for variant in (None, "direct", "generic"):
        name, r = chain(c, variant)
        tr = r["extras"]["trace_surv"]
        if ref is None:
            ref = tr
        err = float(np.max(np.abs(tr - ref) / np.maximum(np.abs(ref), 1e-12)))
        print(f"{cfg} n={n} kernel={name}: trace rel diff vs first variant {err:.2e}", flush=True)
        assert err < 1e-9, (cfg, name, err)
os.environ.pop("JMB_KERNEL", None)
c = make_multistate_cohort(n=2000)
name, r = chain(c, None, ms=True)
print(f"multistate n=2000 kernel={name}: ok, accept_surv={float(r['extras']['accept_surv'][0]):.3f}", flush=True)
print("sanitize_chain: done")
