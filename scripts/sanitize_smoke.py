import numpy as np, sys
sys.path.insert(0, '.')
from jmbayes_b200 import api, survfit as sv
from jmbayes_b200.cohorts import make_cohort
for cfg in ("config4", "config2"):
    c = make_cohort(cfg, n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=10, n_block=10, n_thin=10)
    with api.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"], basis_from_knots=True) as fit:
        fit.set_xdesigns(c["Data"]); fit.set_draw_betas(c["Data"]["betas"], c["Data"]["sigmas"], c["Data"]["invD"])
        r = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    d = sv.newdata_designs(c, L=3, dense_basis=False)
    means, draws = sv.posterior_draws(c, M=10)
    with sv.SurvFit(d) as h:
        g = h.survfitJM(means, draws)
    print(cfg, "ok", float(np.nanmean(g["summaries"])))
import os
os.environ["JMB_SVFT_KERNEL"] = "generic"
c = make_cohort("config3", n=100); d = sv.newdata_designs(c, L=3); means, draws = sv.posterior_draws(c, M=6)
with sv.SurvFit(d) as h:
    g = h.survfitJM(means, draws)
print("generic ok", float(np.nanmean(g["summaries"])))
