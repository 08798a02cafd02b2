"""Golden description of the one HDF5 file the reference ships (Keras-2.0.8/examples/mymodel.h5, written by the real
HDF5 library through h5py): group tree, attributes and a digest of every dataset, as read by h5lite.  Run in the build
container (python tests/golden/make_h5_golden.py); tests/test_h5lite.py checks h5lite against it when the file exists."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SRC = "/root/reference/Keras-2.0.8/examples/mymodel.h5"


def describe(path=SRC):
    from h_denseunet_b200 import h5lite
    f = h5lite.File(path)
    out = {"datasets": {}, "attrs": {}}

    def walk(node, prefix):
        for k, v in node.attrs.items():
            if hasattr(v, "tolist"):
                v = v.tolist()
            if isinstance(v, (list, tuple)):
                v = [x.decode() if isinstance(x, bytes) else x for x in v]
            elif isinstance(v, bytes):
                v = v.decode()
            if isinstance(v, str) and len(v) > 200:
                v = "sha256:" + hashlib.sha256(v.encode()).hexdigest()
            out["attrs"][prefix + "@" + k] = v
        if node.is_group:
            for k in node.keys():
                walk(node[k], prefix + "/" + k)
        else:
            a = node.value()
            out["datasets"][prefix] = {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
                                      "mean": float(a.astype("float64").mean()) if a.size else 0.0}

    walk(f, "")
    return out


if __name__ == "__main__":
    d = describe()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "h5_mymodel_golden.json"), "w") as fh:
        json.dump(d, fh, indent=1, sort_keys=True)
    print(len(d["datasets"]), "datasets", len(d["attrs"]), "attributes")
