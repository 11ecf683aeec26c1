"""medmamba_b200 -- B200-native (sm_100a) SS2D hot path of MedMamba behind the reference's API."""
from .selective_scan_interface import selective_scan_fn  # noqa: F401
