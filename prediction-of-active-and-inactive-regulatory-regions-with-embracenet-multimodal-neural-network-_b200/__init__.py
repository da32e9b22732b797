"""B200-native EmbraceNet engine: CUDA kernels + C ABI (csrc/, lib/) and the host-side mirror of the
reference's BIOINF_tesi.models interface (BIOINF_tesi/).

The directory name is not a Python identifier; import it through the `embrace_b200` alias package at
the repository root (it points its __path__ here)."""
from ._native import EmbError, build, lib          # noqa: F401
from .archspec import ArchSpec                      # noqa: F401
from .engine import Engine                          # noqa: F401
