"""The C-ABI library loads on a CPU-only box and exports every symbol include/clipself_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "clipself_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("cs_extract_rois", "cs_roi_align_fwd", "cs_roi_align_bwd", "cs_cosine_loss_fwd", "cs_cosine_loss_bwd",
                 "cs_mask_pool_fwd", "cs_gemm_bf16", "cs_attention_fwd", "cs_attention_bwd", "cs_layernorm_fwd",
                 "cs_adamw_step"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from clipself_b200 import build, _lib
    path = build.build()                      # no-op when the in-tree .so is current
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes prototypes cover the same set
    assert set(_lib.PROTOTYPES) | {"cs_last_error"} == set(_declared())
    assert _lib.lib().cs_abi_version() >= 1


def test_no_cuda_device_fails_loudly():
    """On a box without a GPU every product entry point must raise, never fall back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clipself_b200 import _lib
    with pytest.raises(_lib.ClipselfB200Error):
        _lib.require_device()
    from clipself_b200.factory import create_model
    m = create_model("EVA02-CLIP-B-16", "eva", cache_dir="")
    with pytest.raises(_lib.ClipselfB200Error), torch.no_grad():
        m.encode_image(torch.zeros(1, 3, 224, 224))


def test_sass_is_blackwell_native():
    import subprocess
    from clipself_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnem in out, mnem
