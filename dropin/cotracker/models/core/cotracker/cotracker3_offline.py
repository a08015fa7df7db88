from cotracker_b200.model import CoTrackerThreeOffline  # noqa: F401
