"""Drop-in module name of the reference's loader: ``import LoadData as DATA``."""
from cffm_b200.data import LoadData  # noqa: F401
