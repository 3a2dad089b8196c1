"""Input contract of the CLIP conditioning wrapper (reference models/pretrained_clip.py).

CLIP itself (OpenAI ViT-L/14, third-party weights) runs once per stage, is not on the
per-step hot path and is out of scope.  This class keeps the
shape contract (pretrained_clip.py:46-65) and the ``embeddings=`` pass-through of
``ImageCLIP.forward`` (pretrained_clip.py:95-107).  Assign a real CLIP wrapper with
the same methods to ``model.clip`` to condition on images or text.
"""
from typing import Iterable, Optional

import torch


class EmbeddingCLIP:
    feature_dim = 768
    grid_size = 16
    grid_feature_dim = 1024

    def __init__(self, device, **kwargs):
        self.device = device

    def __call__(self, batch_size: int, images=None, texts=None,
                 embeddings: Optional[Iterable[Optional[torch.Tensor]]] = None) -> torch.Tensor:
        if images is not None or texts is not None:
            raise RuntimeError("CLIP weights are not bundled: pass precomputed `embeddings=` or set "
                               "model.clip to a FrozenImageCLIP-compatible object")
        result = torch.zeros((batch_size, self.feature_dim), device=self.device)
        if embeddings is not None:
            if torch.is_tensor(embeddings):
                assert embeddings.shape[0] == batch_size, "number of embeddings should match batch size"
                return embeddings.to(result)
            embeddings = list(embeddings)
            assert len(embeddings) == batch_size, "number of embeddings should match batch size"
            for i, emb in enumerate(embeddings):
                if emb is not None:
                    result[i] = emb.to(result)
        return result

    def embed_images_grid(self, xs):
        raise RuntimeError("CLIP weights are not bundled: pass precomputed grid `embeddings=`")


FrozenImageCLIP = EmbeddingCLIP
ImageCLIP = EmbeddingCLIP
