"""`sparse.linalg` -> legate.sparse_b200.linalg (see sparse/__init__.py)."""
from legate.sparse_b200.linalg import *  # noqa: F401,F403
from legate.sparse_b200.linalg import (LinearOperator, IdentityOperator, bicg, bicgstab, cg, cg_axpby, cgs, eigsh,  # noqa: F401
                                       gmres, lsqr, make_linear_operator)
