"""Un-tar the ILSVRC2012 archives into ImageFolder layout.

Parity: reference ``scripts/prepare_imagenet.py`` — SHA1 check of the two archives (``:18-35``),
``ILSVRC2012_img_train.tar`` is a tar of per-class tars -> ``train/<wnid>/*.JPEG`` (``:38-55``),
``ILSVRC2012_img_val.tar`` is flat -> ``validation/<wnid>/`` using the val filename->wnid map
(``:58-71``; the map ships as ``scripts/imagenet_val_maps.csv`` in the reference — here it comes from the
packaged ``data/imagenet_meta`` lookup unless the caller passes a CSV via ``--val-map`` or drops one next to
the archives).
"""
from __future__ import annotations

import csv
import hashlib
import os
import tarfile
from typing import Dict, Optional

TRAIN_TAR = "ILSVRC2012_img_train.tar"
VAL_TAR = "ILSVRC2012_img_val.tar"
SHA1 = {TRAIN_TAR: "43eda4fe35c1705d6606a6a7a633bc965d194284", VAL_TAR: "5f3f73da3395154b60528b2b2a2caf2374f5f178"}


def sha1_of(path: str, chunk: int = 1 << 22) -> str:
    h = hashlib.sha1()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(chunk), b""):
            h.update(block)
    return h.hexdigest()


def check_sha1(path: str, expected: Optional[str]) -> bool:
    if not expected:
        return True
    got = sha1_of(path)
    if got != expected:
        raise ValueError(f"SHA1 mismatch for {path}: {got} != {expected}")
    return True


def extract_train(tar_path: str, target_dir: str) -> int:
    out = os.path.join(target_dir, "train")
    os.makedirs(out, exist_ok=True)
    n = 0
    with tarfile.open(tar_path) as outer:
        for member in outer:
            if not member.isfile() or not member.name.endswith(".tar"):
                continue
            wnid = os.path.splitext(os.path.basename(member.name))[0]
            cls_dir = os.path.join(out, wnid)
            os.makedirs(cls_dir, exist_ok=True)
            inner = outer.extractfile(member)
            with tarfile.open(fileobj=inner) as cls_tar:
                cls_tar.extractall(cls_dir, filter="data")
                n += len(cls_tar.getnames())
    return n


def load_val_map(path: str) -> Dict[str, str]:
    m: Dict[str, str] = {}
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if len(row) < 2 or row[0].lower() in ("filename", "file", "class"):
                continue
            if row[0].upper().endswith((".JPEG", ".JPG", ".PNG")):
                m[os.path.basename(row[0])] = row[1]            # filename,wnid
            else:
                m[os.path.basename(row[1])] = row[0]            # class,filename (the reference's column order)
    return m


def extract_val(tar_path: str, target_dir: str, val_map: Dict[str, str]) -> int:
    out = os.path.join(target_dir, "validation")
    os.makedirs(out, exist_ok=True)
    n = 0
    with tarfile.open(tar_path) as tf:
        for member in tf:
            if not member.isfile():
                continue
            name = os.path.basename(member.name)
            wnid = val_map.get(name)
            if wnid is None:
                continue
            d = os.path.join(out, wnid)
            os.makedirs(d, exist_ok=True)
            with tf.extractfile(member) as src, open(os.path.join(d, name), "wb") as dst:
                dst.write(src.read())
            n += 1
    return n


def main(download_dir: str, target_dir: str, check: bool = True, val_map: Optional[str] = None) -> Dict[str, int]:
    counts = {}
    train_tar, val_tar = os.path.join(download_dir, TRAIN_TAR), os.path.join(download_dir, VAL_TAR)
    if os.path.isfile(train_tar):
        if check:
            check_sha1(train_tar, SHA1[TRAIN_TAR])
        counts["train"] = extract_train(train_tar, target_dir)
    if os.path.isfile(val_tar):
        if check:
            check_sha1(val_tar, SHA1[VAL_TAR])
        vm_path = val_map or os.path.join(download_dir, "imagenet_val_maps.csv")
        if os.path.isfile(vm_path):
            vm = load_val_map(vm_path)          # a user-supplied ``filename,wnid`` / ``class,filename`` CSV wins
        else:
            from . import imagenet_meta

            vm = imagenet_meta.val_map()        # the packaged ILSVRC2012 ground truth (reference: scripts/imagenet_val_maps.csv)
        counts["validation"] = extract_val(val_tar, target_dir, vm)
    print(counts)
    return counts
