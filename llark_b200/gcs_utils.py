"""Local-filesystem half of m2t/gcs_utils.py: the `.npy` embedding files the Jukebox front end writes and the training
data builders read. `gs://` paths are the reference's Google Cloud Storage I/O, out of scope here:
they raise NotImplementedError instead of importing google.cloud.

    list_files_with_extension   m2t/gcs_utils.py:162-172 (local branch)
    write_npy                   :175-198
    read_audio_encoding         :201-248
"""
import logging
import os
from typing import Sequence, Union

import numpy as np
import torch


def _no_gcs(path):
    raise NotImplementedError(f"{path}: Google Cloud Storage I/O is outside llark_b200's scope (DESIGN.md §7)")


def list_files_with_extension(input_dir: str, extension: str) -> Sequence[str]:
    if input_dir.startswith("gs://"):
        _no_gcs(input_dir)
    input_paths = [os.path.join(input_dir, fp) for fp in os.listdir(input_dir)]
    return sorted([x for x in input_paths if x.endswith(extension)])


def write_npy(filepath: str, ary: np.ndarray) -> None:
    assert filepath.endswith(".npy")
    assert isinstance(ary, np.ndarray)
    if filepath.startswith("gs://"):
        _no_gcs(filepath)
    output_dir = os.path.dirname(filepath)
    if not os.path.exists(output_dir):
        os.makedirs(output_dir)
    np.save(filepath, ary)


def read_audio_encoding(uri: str, representations_dir: str, numpy_to_torch=True) -> Union[None, torch.Tensor, np.ndarray]:
    """`<representations_dir>/<uri>.npy` -> tensor (or ndarray); a missing file is a warning and None."""
    if not isinstance(uri, str):
        logging.debug(f"casting uri {uri} of type {type(uri)} to string {str(uri)}")
        uri = str(uri)
    encoding_fp = os.path.join(representations_dir, uri + ".npy")
    if encoding_fp.startswith("gs://"):
        _no_gcs(encoding_fp)
    audio_encoding = None
    try:
        logging.debug(f"reading local encoding file from {encoding_fp}")
        audio_encoding = np.load(encoding_fp)
    except FileNotFoundError:
        logging.warning(f"no encodings found for {encoding_fp}; skipping")
    if audio_encoding is not None and numpy_to_torch:
        return torch.from_numpy(audio_encoding)
    return audio_encoding
