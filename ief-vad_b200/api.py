"""Public names of the package (the reference's own names for this path)."""
from .imf_vad import MMFMIL, MultiModal_Fusion_Attn_Iter  # noqa: F401

__all__ = ["MMFMIL", "MultiModal_Fusion_Attn_Iter"]
