"""The drop-in C++ host classes (ExtendKF / Map / Tracking, reference method names) driven in System::TrackRunning's call order by
ransac_slam_b200/lib/rslam_replay must reproduce the direct C-ABI run bit for bit (same kernels, same order)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ransac_slam_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_classes_replay(tmp_path):
    exe = os.path.join(ROOT, "ransac_slam_b200", "lib", "rslam_replay")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    scene = synth.make_scene(N=50, seed=33)
    T = 5
    seq = synth.make_sequence(scene, T=T, seed=34)
    n = scene.x0.size
    dump = tmp_path / "dump.bin"
    with open(dump, "wb") as f:
        f.write(struct.pack("<6i", scene.N, n, T, scene.cam.nRows, scene.cam.nCols, seq.u01.shape[1]))
        f.write(scene.cam.as9().tobytes())
        f.write(scene.x0.tobytes())
        f.write(np.asfortranarray(scene.P0).tobytes(order="F"))
        f.write(scene.templates.astype(np.float64).tobytes())
        for k in range(T):
            f.write(seq.images[k].tobytes())
            f.write(seq.u01[k].tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(dump), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    rec = 13 * 8 + 12
    assert len(raw) == T * rec
    g = H.gpu_from(scene, scene.x0, scene.P0, prior=False)
    g.set_graph(False)
    for k in range(T):
        g.frame(seq.images[k][None], seq.u01[k][None])
        x13 = np.frombuffer(raw[k * rec:k * rec + 104], dtype=np.float64)
        cnt = np.frombuffer(raw[k * rec + 104:(k + 1) * rec], dtype=np.int32)
        ft = g.features()
        assert np.array_equal(x13, g.download_pose()), k
        assert list(cnt) == [int(ft["ic"].sum()), int(ft["li"].sum()), int(ft["hi"].sum())]


def test_pgm_replay_through_host_classes_matches_the_reference(tmp_path):
    """configuration C1 end to end: frames of the reference's bundled sequence written as P5 files, replayed by the C++ drop-in
    classes in System::System / System::TrackRunning order (rslam_replay_pgm), against the outputs of the reference's own sources
    (tests/golden/ref_vectors.npz) fed the same libc draws."""
    from oracle import ref_py as R  # only for the settings template (numbers of examples/Monocular/initialize_param.yaml)
    from tests import ref_cases as RC

    exe = os.path.join(ROOT, "ransac_slam_b200", "lib", "rslam_replay_pgm")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    g = RC.load()
    frames = np.load(os.path.join(RC.GOLD, "pgm_frames.npz"))["frames"]
    nf = int(g["bundled_frames"])
    yaml = tmp_path / "settings.yaml"
    yaml.write_text(R.YAML_TEMPLATE.format(**R.BUNDLED_YAML))
    names = []
    with open(tmp_path / "draws.bin", "wb") as d:
        for k in range(nf):
            p = tmp_path / f"rawoutput{k:04d}.pgm"
            with open(p, "wb") as f:
                f.write(b"P5\n# bundled frame\n320 240\n255\n")
                f.write(frames[k].tobytes())
            names.append(str(p))
            d.write(g[f"bundled_k{k}_draws_map"][:100].astype("<i4").tobytes())
            d.write(g[f"bundled_k{k}_draws_ransac"].astype("<i4").tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(yaml), str(out), "--draws", str(tmp_path / "draws.bin")] + names, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    rec = 13 * 8 + 6 * 4
    assert len(raw) == nf * rec
    for k in range(nf):
        x13 = np.frombuffer(raw[k * rec:k * rec + 104], dtype=np.float64)
        cnt = np.frombuffer(raw[k * rec + 104:(k + 1) * rec], dtype=np.int32)
        pre = f"bundled_k{k}_"
        H.assert_x_close(x13, g[pre + "x"][:13], what=f"camera state after frame {k}")
        assert cnt[0] == g[pre + "types"].size and cnt[1] == g[pre + "ic"].sum() and cnt[2] == g[pre + "li"].sum() and cnt[3] == g[pre + "hi"].sum(), (k, cnt)
