"""preprocessing.py of the reference (preprocessing.py:7-85) without its file-format dependency.

The reference reads NIfTI volumes with medpy.io.load (absent from this image; no network) and writes
  * HU-clipped float32 volumes              `proprecessing`       (preprocessing.py:7-20:  img[img<-200]=-200; img[img>250]=250)
  * per-volume index lists of liver / tumour voxels  `generate_livertxt` / `generate_tumortxt` (:22-64: np.where(label == 1 | 2),
    "x y z" rows with fmt "%d", one trailing blank line)
  * per-volume liver bounding boxes          `generate_txt`        (:66-79: min / max of the liver index list, fmt "%d")
which train_hybrid.py / train_2ddense.py read back (train_hybrid.py:160-197).  The array functions below are exact
restatements on numpy arrays; the file functions take any loader `load(path) -> (array, header)` (medpy's signature) and
default to numpy .npy volumes, so the scripts' directory layout is produced unchanged when a NIfTI loader is available."""
import os

import numpy as np


def clip_hu(img, lo=-200, hi=250):
    """preprocessing.py:15-17."""
    img = np.array(img, copy=True)
    img[img < lo] = lo
    img[img > hi] = hi
    return np.array(img, dtype="float32")


def voxel_index_list(label, value):
    """preprocessing.py:31-36 / :52-58: (n, 3) int array of the voxels where label == value, in np.where (C) order."""
    idx = np.where(np.asarray(label) == value)
    return np.c_[idx[0], idx[1], idx[2]]


def write_index_txt(path, coords):
    """np.savetxt(f, np.c_[x, y, z], fmt="%d") followed by one newline, as the reference writes it."""
    with open(path, "w") as f:
        np.savetxt(f, coords, fmt="%d")
        f.write("\n")


def liver_box(coords):
    """preprocessing.py:73-77: [min x, min y, min z, max x, max y, max z]."""
    coords = np.asarray(coords).reshape(-1, 3)
    return np.append(np.min(coords, axis=0), np.max(coords, axis=0), axis=0)


def _npy_load(path):
    return np.load(path), None


def _npy_save(arr, path):
    np.save(path, arr)


def proprecessing(image_path, save_folder, load=_npy_load, save=_npy_save, root="data/"):
    """preprocessing.py:7-20 (the reference's spelling)."""
    out = os.path.join(root, save_folder)
    os.makedirs(out, exist_ok=True)
    for name in sorted(f for f in os.listdir(image_path) if "volume" in f):
        img, _ = load(os.path.join(image_path, name))
        save(clip_hu(img), os.path.join(out, name))


def generate_label_txt(image_path, save_folder, n, value, sub, stem, load=_npy_load, root="data/", ext=".npy"):
    out = os.path.join(root, save_folder, sub)
    os.makedirs(out, exist_ok=True)
    for i in range(n):
        lab, _ = load(os.path.join(image_path, "segmentation-%d%s" % (i, ext)))
        write_index_txt(os.path.join(out, "%s_%d.txt" % (stem, i)), voxel_index_list(lab, value))


def generate_livertxt(image_path, save_folder, n=131, **kw):
    generate_label_txt(image_path, save_folder, n, 1, "LiverPixels", "liver", **kw)        # preprocessing.py:22-41


def generate_tumortxt(image_path, save_folder, n=131, **kw):
    generate_label_txt(image_path, save_folder, n, 2, "TumorPixels", "tumor", **kw)        # preprocessing.py:43-64


def generate_txt(save_folder, n=131, root="data/"):
    """preprocessing.py:66-79: liver bounding boxes from the LiverPixels lists."""
    out = os.path.join(root, save_folder, "LiverBox")
    os.makedirs(out, exist_ok=True)
    for i in range(n):
        values = np.loadtxt(os.path.join(root, save_folder, "LiverPixels", "liver_%d.txt" % i), delimiter=" ", usecols=[0, 1, 2])
        np.savetxt(os.path.join(out, "box_%d.txt" % i), liver_box(values), fmt="%d")
