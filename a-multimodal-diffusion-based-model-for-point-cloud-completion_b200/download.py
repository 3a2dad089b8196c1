"""Checkpoint lookup with the reference's names and cache layout (models/download.py:14-78), local files
only: this package never opens a network connection.  ``load_checkpoint("base40M-imagevec", device)`` reads
``<cache_dir>/base_40m_imagevec.pt`` -- the file the reference's downloader would have left there -- and
returns the ``state_dict`` to pass to ``model.load_state_dict`` (same keys, transformer.py of this shim)."""
import os
from typing import Dict, Optional

import torch

# checkpoint name -> file name inside the cache directory (the basenames of the reference's URLs)
CHECKPOINT_FILES = {
    "base40M-imagevec": "base_40m_imagevec.pt",
    "base40M-textvec": "base_40m_textvec.pt",
    "base40M-uncond": "base_40m_uncond.pt",
    "base40M": "base_40m.pt",
    "base300M": "base_300m.pt",
    "base1B": "base_1b.pt",
    "upsample": "upsample_40m.pt",
}


def default_cache_dir() -> str:
    return os.path.join(os.path.abspath(os.getcwd()), "point_e_model_cache")


def checkpoint_path(checkpoint_name: str, cache_dir: Optional[str] = None) -> str:
    if checkpoint_name not in CHECKPOINT_FILES:
        raise ValueError(f"Unknown checkpoint name {checkpoint_name}. Known names are: {sorted(CHECKPOINT_FILES)}.")
    return os.path.join(cache_dir or default_cache_dir(), CHECKPOINT_FILES[checkpoint_name])


def load_checkpoint(checkpoint_name: str, device: torch.device, progress: bool = True,
                    cache_dir: Optional[str] = None, chunk_size: int = 4096) -> Dict[str, torch.Tensor]:
    """Same signature as the reference (``progress`` / ``chunk_size`` only matter to a downloader)."""
    path = checkpoint_path(checkpoint_name, cache_dir)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: place the reference checkpoint there (no network access from this package)")
    return torch.load(path, map_location=device)
