"""Drop-in module name of the reference (Utils/Utils.py): the functions the hot-path scripts import from it -- the .hair
codecs (Utils.py:25-66, 1246-1262) and strand smoothing (:1148-1198).  The rest of the reference's Utils.py (mesh / asset
preparation) is outside this repository's scope."""
from monohair_b200.hairgrow import load_strand, save_hair_strands, smooth_strands  # noqa: F401
