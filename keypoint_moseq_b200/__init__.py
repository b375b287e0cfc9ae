"""B200-native keypoint-SLDS Gibbs sweep (drop-in for the jax_moseq `resample_model`
path driven by keypoint_moseq.fit_model / apply_model)."""
