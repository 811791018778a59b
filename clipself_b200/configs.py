"""Architecture registry: the two EVA02-CLIP configs the reference's scripts use
(src/open_clip/eva_clip/model_configs/EVA02-CLIP-B-16.json, EVA02-CLIP-L-14-336.json)."""

MODEL_CONFIGS = {
    "EVA02-CLIP-B-16": {
        "embed_dim": 512,
        "vision_cfg": {"image_size": 224, "layers": 12, "width": 768, "head_width": 64, "patch_size": 16,
                       "mlp_ratio": 2.6667, "eva_model_name": "eva-clip-b-16-X", "drop_path_rate": 0.0,
                       "xattn": True, "fusedLN": True, "rope": True, "pt_hw_seq_len": 16, "intp_freq": True,
                       "naiveswiglu": True, "subln": True},
        "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": 512, "heads": 8, "layers": 12,
                     "xattn": True, "fusedLN": True},
    },
    "EVA02-CLIP-L-14-336": {
        "embed_dim": 768,
        "vision_cfg": {"image_size": 336, "layers": 24, "width": 1024, "drop_path_rate": 0, "head_width": 64,
                       "mlp_ratio": 2.6667, "patch_size": 14, "eva_model_name": "eva-clip-l-14-336",
                       "xattn": True, "fusedLN": True, "rope": True, "pt_hw_seq_len": 16, "intp_freq": True,
                       "naiveswiglu": True, "subln": True},
        "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": 768, "heads": 12, "layers": 12,
                     "xattn": False, "fusedLN": True},
    },
}


def list_models():
    return sorted(MODEL_CONFIGS)


def get_model_config(name: str):
    key = name.replace("/", "-")
    if key not in MODEL_CONFIGS:
        raise RuntimeError(f"Model config for {name} not found; available: {list_models()}")
    import copy
    return copy.deepcopy(MODEL_CONFIGS[key])
