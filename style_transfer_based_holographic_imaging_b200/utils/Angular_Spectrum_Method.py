from ..Angular_Spectrum_Method import ASM, torch_fft, torch_ifft, center_crop  # noqa: F401
