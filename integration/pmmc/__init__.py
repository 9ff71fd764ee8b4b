"""pmmc front end on the mmc_b200 engine: the package layout of the reference's pmmc/pmmc/__init__.py (run, gpuinfo, version from
the compiled module _pmmc), with _pmmc built from integration/pmmc_b200.cpp over the C-ABI of include/mmc_b200.h.
    import sys; sys.path.insert(0, "<repo>/integration"); import pmmc; res = pmmc.run(cfg)"""
try:
    from ._pmmc import gpuinfo, run, version  # noqa: F401
except ImportError as e:  # pragma: no cover
    raise ImportError("the pmmc binary extension (_pmmc) is not compiled: run `python -m mmc_b200.build` (%s)" % e)
