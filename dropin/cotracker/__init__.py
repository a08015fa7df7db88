"""Drop-in shim: put `<repo>/dropin` (and `<repo>`) on PYTHONPATH and the reference's import paths
(`cotracker.predictor`, `cotracker.models.build_cotracker`, ...) resolve to cotracker_b200."""
