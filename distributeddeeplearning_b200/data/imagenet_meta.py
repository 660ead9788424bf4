"""ILSVRC2012 lookup data.

The reference ships three lookup files — ``scripts/imagenet_class_index.json`` (class index -> [wnid, noun]),
``scripts/imagenet_val_maps.csv`` (validation file name -> wnid, consumed by ``scripts/prepare_imagenet.py:58-71``) and
``TensorFlow_imagenet/src/imagenet_nounid_to_class.json`` (wnid -> class index).  All three are views of the same
devkit facts, which this package keeps ONCE, compactly, in ``ilsvrc2012_meta.json`` (sorted wnids, nouns, and the
50,000 validation labels as zlib+base85 of uint16) and materialises on demand in the reference's formats.
"""
from __future__ import annotations

import base64
import functools
import json
import os
import struct
import zlib
from typing import Dict, List, Tuple

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ilsvrc2012_meta.json")


@functools.lru_cache(maxsize=None)
def _meta() -> dict:
    with open(_PATH) as f:
        return json.load(f)


def wnids() -> List[str]:
    """The 1000 WordNet ids in class-index order (sorted — the order torchvision's ImageFolder produces)."""
    return list(_meta()["wnids"])


def class_index() -> Dict[str, Tuple[str, str]]:
    """``{"0": ("n01440764", "tench"), ...}`` — the content of the reference's imagenet_class_index.json."""
    m = _meta()
    return {str(i): (w, n) for i, (w, n) in enumerate(zip(m["wnids"], m["nouns"]))}


def nounid_to_class() -> Dict[str, int]:
    """``{"n01440764": 0, ...}`` — the content of the reference's imagenet_nounid_to_class.json."""
    return {w: i for i, w in enumerate(_meta()["wnids"])}


@functools.lru_cache(maxsize=None)
def val_labels() -> Tuple[int, ...]:
    """Class index of ILSVRC2012_val_00000001.JPEG ... ILSVRC2012_val_00050000.JPEG."""
    raw = zlib.decompress(base64.b85decode(_meta()["val_labels_z85"]))
    return struct.unpack("<%dH" % (len(raw) // 2), raw)


def val_filename(i: int) -> str:
    return "ILSVRC2012_val_%08d.JPEG" % (i + 1)


def val_map() -> Dict[str, str]:
    """Validation file name -> wnid: what the reference reads from imagenet_val_maps.csv."""
    w = _meta()["wnids"]
    return {val_filename(i): w[c] for i, c in enumerate(val_labels())}


def write_reference_files(directory: str) -> List[str]:
    """Materialise the three lookup files in the reference's own formats (CSV header ``class,filename``)."""
    os.makedirs(directory, exist_ok=True)
    out = []
    p = os.path.join(directory, "imagenet_class_index.json")
    with open(p, "w") as f:
        json.dump({k: list(v) for k, v in class_index().items()}, f, indent=1)
    out.append(p)
    p = os.path.join(directory, "imagenet_val_maps.csv")
    with open(p, "w") as f:
        f.write("class,filename\n")
        for name, wn in val_map().items():
            f.write(f"{wn},{name}\n")
    out.append(p)
    p = os.path.join(directory, "imagenet_nounid_to_class.json")
    with open(p, "w") as f:
        json.dump(nounid_to_class(), f)
    out.append(p)
    return out
