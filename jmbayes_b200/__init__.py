"""jmbayes_b200 - B200-native MCMC hot path of JMbayes' mvJointModelBayes().

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/jmb200.h``), ``api.py`` (host-side mirror of the reference's Rcpp entry points),
``cohorts.py`` (seeded synthetic cohorts in the reference's ``Data`` contract) and
``sharding.py`` (contiguous subject shards for one-process-per-GPU runs).
"""
from .cohorts import make_cohort  # noqa: F401

__all__ = ["make_cohort"]
