"""`sparse.linalg` -> legate.sparse_b200.linalg (see sparse/__init__.py)."""
from legate.sparse_b200.linalg import *  # noqa: F401,F403
from legate.sparse_b200.linalg import LinearOperator, IdentityOperator, cg, cg_axpby, make_linear_operator  # noqa: F401
