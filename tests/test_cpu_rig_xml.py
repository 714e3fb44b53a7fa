"""CPU: the rig XML reader of the C++ host mirror (calibu::ReadXmlRig stand-in; the reference reads such files
for `-model_files` warm starts, vicalib-engine.cc:188-196) on a fixture in the format WriteCameraModels writes."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build():
    subprocess.run(["make", "-C", CPP, "xml_reader"], check=True, capture_output=True, timeout=300)
    return os.path.join(CPP, "xml_reader")


def test_reader_parses_the_writer_format():
    exe = _build()
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "cameras_fixture.xml")], check=True, capture_output=True,
                         text=True, timeout=60).stdout.strip().splitlines()
    assert len(out) == 2
    a, b = (line.split() for line in out)
    assert a[:4] == ["calibu_fu_fv_u0_v0_k1_k2_k3", "0", "640", "480"] and b[:4] == ["calibu_fu_fv_u0_v0_w", "1", "752", "480"]

    def fields(tok):
        i, j, k = tok.index("params"), tok.index("rdf"), tok.index("T_wc")
        return (np.array(tok[i + 1:j], float), np.array(tok[j + 1:k], float).reshape(3, 3), np.array(tok[k + 1:], float).reshape(3, 4))

    pa, ra, Ta = fields(a)
    pb, rb, Tb = fields(b)
    assert np.array_equal(pa, [321.5, 322.25, 318.75, 241.125, -0.15, 0.03, -0.004])
    assert np.array_equal(pb, [300, 300.5, 376, 240, 0.9])
    # right / down / forward are the COLUMNS of the RDF matrix: camera 0 is RdfRobotics, camera 1 RdfVision
    assert np.array_equal(ra, [[0, 1, 0], [0, 0, 1], [1, 0, 0]]) and np.array_equal(rb, np.eye(3))
    assert np.allclose(Ta, [[0, 0, 1, 0.02], [1, 0, 0, -0.01], [0, 1, 0, 0.005]], atol=1e-15)
    assert np.allclose(Tb, [[1, 0, 0, 0.1], [0, 1, 0, 0], [0, 0, 1, 0]], atol=1e-15)


def test_reader_rejects_garbage(tmp_path):
    exe = _build()
    bad = tmp_path / "bad.xml"
    bad.write_text("<rig><camera><camera_model type=\"calibu_fu_fv_u0_v0_w\"><width> 1 </width></camera_model></camera></rig>")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and r.stdout.startswith("error")
    r = subprocess.run([exe, str(tmp_path / "missing.xml")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "cannot open" in r.stdout
