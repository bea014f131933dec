"""dpi_b200 - sm_100a implementation of the DPI variational-posterior training step.

Sub-modules mirror the reference's import surface (DPItorch/):
    dpi_b200.generative_model.realnvpfc_model   RealNVP
    dpi_b200.interferometry_helpers             Obs_params_torch, eht_observation_pytorch, Loss_*
    dpi_b200.MRI_helpers                        Loss_kspace_diff*, Loss_l1/TSV/TV, fft2c_torch
    dpi_b200.trainer                            fused training steps (the benchmarked hot path)
"""
__version__ = "0.1.0"
