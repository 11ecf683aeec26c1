"""medmamba_b200 -- B200-native (sm_100a) SS2D hot path of MedMamba behind the reference's API."""
from .selective_scan_interface import selective_scan_fn  # noqa: F401
from .model import (SS2D, SS_Conv_SSM, VSSLayer, VSSM, PatchEmbed2D, PatchMerging2D, channel_shuffle,  # noqa: F401
                    medmamba_t, medmamba_s, medmamba_b)
from .infer import GraphedForward, InferencePipeline  # noqa: F401
from . import trainer  # noqa: F401,E402
