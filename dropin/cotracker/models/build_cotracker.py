from cotracker_b200.build import build_cotracker  # noqa: F401
