"""Drop-in replacement for the reference's ``models`` package (``models/__init__.py:2-3``).

Put this directory ahead of the reference checkout on ``sys.path`` (or copy it over ``models/``):
``main.py:12`` does ``from models import *`` and then ``EliMRec(config, dataset).to(config.device)``.
"""
from elimrec_b200.model import BasicModel, EliMRec  # noqa: F401

__all__ = ["BasicModel", "EliMRec"]
