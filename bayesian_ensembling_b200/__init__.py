"""B200-native fit -> weight -> barycentre hot path of mattramos/bayesian_ensembling."""
__version__ = "0.1.0"
