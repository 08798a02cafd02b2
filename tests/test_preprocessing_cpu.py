"""preprocessing.py:7-85 restated without medpy (h_denseunet_b200/preprocessing.py): HU clipping, voxel index lists in the
reference's text format, liver boxes -- on a synthetic labelled volume, read back the way train_hybrid.py:160-197 does."""
import os

import numpy as np

from h_denseunet_b200 import preprocessing as pp
from h_denseunet_b200.synthetic import synthetic_slab


def test_directory_layout_and_contents(tmp_path):
    src = tmp_path / "TrainingData"
    os.makedirs(str(src))
    vols = []
    for i in range(2):
        vol, lab = synthetic_slab(1, 32, 12, seed=50 + i)
        v = (vol[0, :, :, :, 0] * 3.0).astype(np.float32)              # stretch beyond [-200, 250]
        l = lab[0, :, :, :, 0]
        np.save(str(src / ("volume-%d.npy" % i)), v)
        np.save(str(src / ("segmentation-%d.npy" % i)), l)
        vols.append((v, l))
    root = str(tmp_path / "data") + "/"
    pp.proprecessing(str(src) + "/", "myTrainingData/", root=root)
    pp.generate_livertxt(str(src) + "/", "myTrainingDataTxt/", n=2, root=root)
    pp.generate_tumortxt(str(src) + "/", "myTrainingDataTxt/", n=2, root=root)
    pp.generate_txt("myTrainingDataTxt/", n=2, root=root)
    for i, (v, l) in enumerate(vols):
        c = np.load(root + "myTrainingData/volume-%d.npy" % i)
        assert c.dtype == np.float32 and c.min() >= -200 and c.max() <= 250
        assert np.array_equal(c, np.clip(v, -200, 250))
        txt = open(root + "myTrainingDataTxt/LiverPixels/liver_%d.txt" % i).read()
        assert txt.endswith("\n\n")                                   # savetxt rows + the reference's extra newline
        rows = np.loadtxt(root + "myTrainingDataTxt/LiverPixels/liver_%d.txt" % i, delimiter=" ", usecols=[0, 1, 2]).astype(int)
        assert np.array_equal(rows, np.argwhere(l == 1))
        trows = np.loadtxt(root + "myTrainingDataTxt/TumorPixels/tumor_%d.txt" % i, delimiter=" ", usecols=[0, 1, 2]).astype(int)
        assert np.array_equal(trows.reshape(-1, 3), np.argwhere(l == 2))
        box = np.loadtxt(root + "myTrainingDataTxt/LiverBox/box_%d.txt" % i).astype(int)
        assert np.array_equal(box, np.concatenate([np.argwhere(l == 1).min(0), np.argwhere(l == 1).max(0)]))
        # the line format train_hybrid.py:52 parses: np.fromstring(line, dtype=int, sep=' ')
        first = txt.splitlines()[0]
        assert np.array_equal(np.array(first.split(), dtype=int), rows[0])
