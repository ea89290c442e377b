"""Loader for the `geordd.jl_b200/` package (a dot is not importable with a plain `import`).

    import georddb200
    G = georddb200.load()          # module object, registered as sys.modules["geordd_jl_b200"]
    gp = G.GPE(x, y, G.MeanZero(), G.SEIso(ll, ls) + G.Const(lc), logNoise)
"""
import importlib.util
import os
import sys

_NAME = "geordd_jl_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geordd.jl_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force=False, verbose=False):
    spec = importlib.util.spec_from_file_location("_geordd_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)
