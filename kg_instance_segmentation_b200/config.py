"""Keypoint-graph constants — same names and values as the reference's config.py:2-28."""
EDGES = [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4)]
NUM_KPS = 5
NUM_EDGES = len(EDGES)
KP_RADIUS = 5
KEYPOINTS = ["tl", "tr", "bl", "br", "center"]
# decode constants the reference hard-codes at the call sites
PEAK_THRESH = 0.004      # postprocessing.py:145
GAUSS_SIGMA = 2          # postprocessing.py:144
BOX_SCALES = (1, 2, 4, 8)  # postprocessing.py:256-259
