"""Import alias: the package directory is named ``multimodal-sentiment-analyzer_b200`` (a hyphen
is not importable), so ``import msa_b200`` loads that directory as the package ``msa_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodal-sentiment-analyzer_b200")
_spec = importlib.util.spec_from_file_location("msa_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["msa_b200"] = _mod
_spec.loader.exec_module(_mod)
