from cotracker_b200.predictor import CoTrackerOnlinePredictor, CoTrackerPredictor  # noqa: F401
