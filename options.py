"""Drop-in module name of the reference (options.py)."""
from monohair_b200.options import *  # noqa: F401,F403
from monohair_b200.options import load_options, override_options, parse_arguments, process_options, save_options_file, set  # noqa: F401
