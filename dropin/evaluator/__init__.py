"""Drop-in replacement for the reference's ``evaluator`` package (``evaluator/__init__.py``)."""
from elimrec_b200.evaluator import AbstractEvaluator, ProxyEvaluator, UniEvaluator  # noqa: F401
