"""admm-library_b200: B200-native batched ADMM solver for convex optimal-control QPs
(host-side mirror of the MATLAB surface over the C ABI of include/admm_b200.h).

The directory name is not a valid Python identifier; import it through
`__graft_entry__.load_pkg()` (registers it as `admm_library_b200`)."""
from . import dist, problems  # noqa: F401
from ._lib import AdmmError, load  # noqa: F401
from .solver import Solver, admm_solve  # noqa: F401
