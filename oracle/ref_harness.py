"""Import harness for the UNMODIFIED reference (test infrastructure only).

The reference tree at /root/reference imports itself as ``point_e`` and pulls in
third-party modules that are absent from this image (open3d, timm, clip).  This
module builds the import-time shims described in SURVEY.md section 8c so that the
reference's own modules can be executed on CPU to (a) pin the oracle restatement
in ``oracle/`` and (b) generate the golden vectors under ``tests/golden/``.

It only works in the authoring container (``/root/reference`` does not exist on
the GPU box); nothing in the product path, ``smoke()``, ``bench.py`` or the
``-m gpu`` tests imports it.
"""
import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("PCD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffusion"))


class _StubFrozenCLIP:
    """Stand-in for ``FrozenImageCLIP`` (reference models/pretrained_clip.py:219-259).

    Exposes the three shape properties (pretrained_clip.py:46-65) and the
    ``embeddings=`` pass-through of ``ImageCLIP.forward`` with
    ``ensure_used_params=False`` (pretrained_clip.py:95-107): a zero row for a
    ``None`` entry, the given vector otherwise.  The CLIP network itself is
    third-party (OpenAI CLIP ViT-L/14, unpinned) and not on the per-step path.
    """

    feature_dim = 768
    grid_size = 16
    grid_feature_dim = 1024

    def __init__(self, device, **kwargs):
        self.device = device

    def __call__(self, batch_size, images=None, texts=None, embeddings=None):
        import torch

        assert images is None and texts is None, "stub CLIP only supports embeddings="
        result = torch.zeros((batch_size, self.feature_dim), device=self.device)
        if embeddings is not None:
            embeddings = list(embeddings)
            assert len(embeddings) == batch_size
            for i, emb in enumerate(embeddings):
                if emb is not None:
                    result[i] = emb.to(result)
        return result

    def embed_images_grid(self, xs):
        raise RuntimeError("stub CLIP cannot embed images; pass embeddings=")


def load_reference():
    """Return the reference package imported as ``point_e`` (with shims)."""
    if "point_e" in sys.modules and getattr(sys.modules["point_e"], "_pcd_harness", False):
        return sys.modules["point_e"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # 1. `point_e` -> /root/reference via a symlink in a temp dir on sys.path.
    link_dir = tempfile.mkdtemp(prefix="pcd_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(link_dir, "point_e"))
    sys.path.insert(0, link_dir)
    sys.dont_write_bytecode = True  # /root/reference is read-only

    # 2. absent third-party modules.
    if "open3d" not in sys.modules:  # models/util.py:6, unused on the path
        sys.modules["open3d"] = types.ModuleType("open3d")
    try:
        importlib.import_module("timm")
    except Exception:
        import torch.nn as nn

        class Mlp(nn.Module):  # timm.models.vision_transformer.Mlp stand-in (TwoStream only)
            def __init__(self, in_features, hidden_features=None, out_features=None,
                         act_layer=nn.GELU, drop=0.0, **kw):
                super().__init__()
                out_features = out_features or in_features
                hidden_features = hidden_features or in_features
                self.fc1 = nn.Linear(in_features, hidden_features)
                self.act = act_layer()
                self.drop1 = nn.Dropout(drop)
                self.norm = nn.Identity()
                self.fc2 = nn.Linear(hidden_features, out_features)
                self.drop2 = nn.Dropout(drop)

            def forward(self, x):
                return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0, **kw):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        vt = types.ModuleType("timm.models.vision_transformer")
        vt.Mlp, vt.DropPath = Mlp, DropPath
        timm.models, models.vision_transformer = models, vt
        sys.modules.update({"timm": timm, "timm.models": models,
                            "timm.models.vision_transformer": vt})

    pkg = importlib.import_module("point_e")
    tr = importlib.import_module("point_e.models.transformer")
    # 3. CLIP: replace the constructor the transformer classes call.
    tr.FrozenImageCLIP = _StubFrozenCLIP
    tr.ImageCLIP = _StubFrozenCLIP
    pkg._pcd_harness = True
    return pkg


def ref_modules():
    """Convenience: the reference modules on the hot path."""
    load_reference()
    return types.SimpleNamespace(
        transformer=importlib.import_module("point_e.models.transformer"),
        perceiver=importlib.import_module("point_e.models.perceiver"),
        rotary=importlib.import_module("point_e.models.rotaryencoderpcd"),
        model_configs=importlib.import_module("point_e.models.configs"),
        diffusion_configs=importlib.import_module("point_e.diffusion.configs"),
        k_diffusion=importlib.import_module("point_e.diffusion.k_diffusion"),
        gaussian_diffusion=importlib.import_module("point_e.diffusion.gaussian_diffusion"),
        sampler=importlib.import_module("point_e.diffusion.sampler"),
        util=importlib.import_module("point_e.models.util"),
    )
