"""gameplay_vision_llm_b200 — B200-native per-frame perception embedding path.

Drop-in for ONE hot path of chasemetoyer/gameplay-vision-llm: decoded frames -> SigLIP2-so400m vision
tower -> ProjectorBank MLP -> timeline index (+ cosine top-k), behind the reference's own call
signatures (`SigLIPSemanticEncoder.encode_image`, `ProjectorBank.project_region`, `FeatureCache`).
Python host code + a C-ABI library of hand-written sm_100a CUDA kernels (`include/gvl.h`,
`gameplay_vision_llm_b200/csrc`).  No CPU fallback: every compute call raises without the extension.
"""
__version__ = "0.1.0"
