"""Import shim: the package directory is `battlezips-halo2_b200/` (hyphen, as the layout contract names it),
which Python cannot import by name; this module loads it as `battlezips_halo2_b200`."""
import importlib.util, os, sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "battlezips-halo2_b200")
_spec = importlib.util.spec_from_file_location(
    "battlezips_halo2_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["battlezips_halo2_b200"] = _mod
_spec.loader.exec_module(_mod)
