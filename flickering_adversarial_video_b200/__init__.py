"""flickering_adversarial_video_b200 — B200-native engine for the flickering-attack optimisation
loop of roiponytch/Flickering_Adversarial_Video.  The hot path lives in libfav.so
(hand-written sm_100a CUDA behind the C-ABI in include/fav.h); this package is the host-side mirror
of the reference's operator surface."""

__all__ = ["_lib", "engine", "synthetic"]
