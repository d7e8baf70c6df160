"""cupyx.scipy.sparse stand-in: SciPy CSR on the host (test infrastructure only)."""
from scipy.sparse import csr_matrix  # noqa: F401
from . import linalg  # noqa: F401
