"""Import alias: ``import icd10_b200`` == importlib.import_module("rag-project-icd10_b200")."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("rag-project-icd10_b200")
sys.modules[__name__] = _pkg
