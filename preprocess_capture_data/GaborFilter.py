"""Drop-in module name of the reference (preprocess_capture_data/GaborFilter.py)."""
from monohair_b200.gabor import batch_generate, calOrientationGabor, calculate_orientation  # noqa: F401
