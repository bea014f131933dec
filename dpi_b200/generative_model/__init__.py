"""Drop-in namespace for DPItorch/generative_model (RealNVP implemented; Glow is a later row)."""
from . import realnvpfc_model  # noqa: F401
