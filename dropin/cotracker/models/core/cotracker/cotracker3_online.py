from cotracker_b200.model import CoTrackerThreeBase, CoTrackerThreeOnline  # noqa: F401
