"""`sparse.io` -> legate.sparse_b200.io (see sparse/__init__.py)."""
from legate.sparse_b200.io import mmread  # noqa: F401
