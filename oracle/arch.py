"""Architecture descriptions used by the oracle and the tests (TEST INFRASTRUCTURE).

The reference reads ``cfg.ARCH`` (a yacs node, ``lib/config/defaults.py:35-100``); yacs is not installed here, so
the oracle hands it a ``SimpleNamespace`` with the same fields.  The definitions live in the product package
(``iodine_b200/config.py``: plain data, no arithmetic) and are re-exported here.
"""
from iodine_b200.config import ALL_ENCODINGS, arch_by_name, make_arch  # noqa: F401
