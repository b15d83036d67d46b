"""Backend selection point of the reference (``evaluator/backend/__init__.py:1-6``): the CUDA backend, no fallback."""
from elimrec_b200.evaluator import UniEvaluator  # noqa: F401
print("Evaluate model with elimrec_b200 (CUDA, sm_100a)")
