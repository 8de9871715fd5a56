"""Drop-in module name of the reference (preprocess_capture_data/calc_orientation_maps.py)."""
from monohair_b200.gabor import calc_confidences, calc_orients, generate_gabor_filters, rgb2gray  # noqa: F401
