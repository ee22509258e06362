"""Host-side mirror of the reference's `guided_diffusion` package for the guided-sampling hot path."""
