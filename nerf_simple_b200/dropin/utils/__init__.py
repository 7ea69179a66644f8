"""`utils` package shim: put this directory's parent first on sys.path (nerf_simple_b200.run does)."""
