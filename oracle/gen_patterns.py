"""Generates tests/golden/patterns_512x432.npz: the reference's three projector patterns (default / kinect / real)
pushed through the reference's OWN pipeline -- data.data_manipulation.read_pattern_file (imported, unmodified) and the
remap + post_process steps of data/create_syn_data.py:283-330 (restated here line by line: that script's main() also
needs the renderer and ShapeNet) -- exactly what the reference stores in settings.pkl as `pattern` and `K`.

TEST INFRASTRUCTURE / DATA PREPARATION: run in the build container (needs /root/reference and cv2); the .npz travels.

    python oracle/gen_patterns.py
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "patterns_512x432.npz")


def remapped_pattern(pattern_type):
    import cv2
    sys.path.insert(0, os.path.join(REF, "data"))
    sys.path.insert(0, REF)
    import data_manipulation as dm   # the reference module, unmodified
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "data"))   # read_pattern_file opens '<type>_pattern.png' relative to the cwd
    try:
        # camera / projector parameters: data/create_syn_data.py:283-311
        if pattern_type == "real":
            fl_proj = fl = 1112.1806640625
            imsize_proj, imsize = (1280, 1080), (1280, 1080)
            K_proj = np.array([[fl_proj, 0, 517.0896606445312], [0, fl_proj, 649.6329956054688], [0, 0, 1]], dtype=np.float32)
            K = np.array([[fl, 0, 517.0896606445312], [0, fl, 649.6329956054688], [0, 0, 1]], dtype=np.float32)
            baseline = 0.0246
        else:
            fl_proj, fl = 1582.06005876, 435.2
            imsize_proj, imsize = (4096, 4096), (512, 432)
            K_proj = np.array([[fl_proj, 0, 2047.5], [0, fl_proj, 2047.5], [0, 0, 1]], dtype=np.float32)
            K = np.array([[fl, 0, 216], [0, fl, 256], [0, 0, 1]], dtype=np.float32)
            baseline = 0.025
        pattern = dm.read_pattern_file(pattern_type, imsize_proj)
        # data/create_syn_data.py:316-330
        x_mesh, y_mesh = np.meshgrid(np.arange(0, imsize[1]), np.arange(0, imsize[0]))
        grid_points = np.stack([x_mesh.reshape(-1), y_mesh.reshape(-1), np.ones(x_mesh.size, x_mesh.dtype)], axis=0)
        mapped = K_proj.dot(np.linalg.inv(K).dot(grid_points))
        mapped = mapped / mapped[2, :]
        x_map = mapped[0, :].reshape(x_mesh.shape).astype("float32")
        y_map = mapped[1, :].reshape(y_mesh.shape).astype("float32")
        mapped_pattern = cv2.remap(pattern, x_map, y_map, cv2.INTER_LINEAR)
        pattern_processed, K_processed = dm.post_process(pattern_type, mapped_pattern, K)
    finally:
        os.chdir(cwd)
    return np.ascontiguousarray(pattern_processed, np.float32), np.asarray(K_processed, np.float32), baseline


def main():
    out = {}
    for kind in ("default", "kinect", "real"):
        pat, K, baseline = remapped_pattern(kind)
        assert pat.shape == (512, 432, 3), pat.shape
        # the three channels are identical copies of the grey pattern; the training code averages them
        # (model/networks.py:344); stored once, as 8-bit-exact float32 where the pipeline kept them so
        grey = pat.mean(axis=2).astype(np.float32)
        out[kind] = grey
        out[kind + "_K"] = K
        out[kind + "_baseline"] = np.float32(baseline)
        print(kind, "mean", float(grey.mean()), "max", float(grey.max()), "K", K.tolist(), "channels equal:",
              bool(np.array_equal(pat[..., 0], pat[..., 1]) and np.array_equal(pat[..., 1], pat[..., 2])))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
