"""Import shim (test infrastructure only) for the three timm symbols the reference's
EVA tower imports (eva_vit_model.py:10-13, eva_clip/transformer.py:11-14, eva_clip/loss.py:18)."""
__version__ = "0.0.0-shim"
