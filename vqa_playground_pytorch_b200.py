"""Import alias.  The package directory is `vqa-playground-pytorch_b200/` (the repo's layout name);
a hyphen cannot appear in a Python identifier, so `import vqa_playground_pytorch_b200` resolves
to this file, which loads that directory as the package of the same (underscored) name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vqa-playground-pytorch_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
