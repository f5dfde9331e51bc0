"""keypoints_detector — B200-native drop-in for the inference hot path of
sandyz1000/face-landmark-detector (same module and function names as the reference package)."""
