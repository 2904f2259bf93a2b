"""Constants of reference data/utils.py that the correction hook needs (markerset_ssm67_smplh
:232-238, marker2bodypart hand ids :252-253)."""
from ..engine import HAND_MARKERS, MARKERSET_SSM67_SMPLH

markerset_ssm67_smplh = list(MARKERSET_SSM67_SMPLH)
marker2bodypart = {"left_hand_ids": HAND_MARKERS[:9], "right_hand_ids": HAND_MARKERS[9:]}
