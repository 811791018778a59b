"""Drop-in namespace: put `<repo>/compat` on PYTHONPATH (after the repo root) and code written against
the reference's `import open_clip` resolves to the B200 implementation (src/open_clip/__init__.py)."""
from clipself_b200 import ClipLoss, create_model, create_model_and_transforms, list_models  # noqa: F401
from clipself_b200.factory import get_cast_dtype, load_checkpoint  # noqa: F401
from clipself_b200.model import CustomCLIP  # noqa: F401


def get_tokenizer(model_name):
    raise NotImplementedError("the text tower / tokenizer are outside the CLIPSelf hot path (SURVEY.md §2)")
