"""the batched force factory (admm-elastic-sca_b200/host/scene/ForceBuilderBatched.cpp) against the
reference's own src/ForceBuilder.cpp.  Both are linked into a headless runner with the reference's SimContext + mclscene
(tests/dropin/Makefile); `dump-forces` loads an XML scene and prints the force list element by element -- no GPU involved.
The two lists must be the same text: same elements, same vertex order, same order in the list, same parameters, same
ForceBuilder::bend_index, for the four shipped scenes and for synthetic scenes that walk the XML vocabulary
(ForceBuilder.cpp:106-435) including its error paths.  The load times are printed: the reference deduplicates hinges with a
linear scan per candidate (ForceBuilder.cpp:55-74)."""
import os
import re
import subprocess
import tarfile

import pytest

from util import ROOT

DROPIN = os.path.join(ROOT, "tests", "dropin", "build")
RUNNERS = [os.path.join(DROPIN, "ref_scene_runner"), os.path.join(DROPIN, "ref_scene_runner_batched")]
XML = dict(bunnyexpand="bunnyexpand.xml", windyflag="cloth.xml", poordillo="poordillo.xml", plinkopony="plinko.xml")

SOLVER = "<solver><iterations value=\"5\" /><timestep value=\"0.04\" /></solver>"


def cloth_xml(w, l, forces, defs):
    return ("<?xml version=\"1.0\"?>\n<mclScene><Object name=\"c\" type=\"plane\"><width value=\"%d\" /><length value=\"%d\" /><Mass value=\"1\" />%s</Object></mclScene>\n"
            "<admmelastic>%s%s</admmelastic>\n" % (w, l, "".join(f"<Force value=\"{f}\" />" for f in forces), defs, SOLVER))


def tet_xml(forces, defs):
    return ("<?xml version=\"1.0\"?>\n<mclScene><Object name=\"b\" type=\"tetmesh\"><File value=\"bunny_1124\" /><Mass value=\"1\" />%s</Object></mclScene>\n"
            "<admmelastic>%s%s</admmelastic>\n" % ("".join(f"<Force value=\"{f}\" />" for f in forces), defs, SOLVER))


SYNTHETIC = {
    # every triangle-mesh force type at once, in an order that interleaves the batches
    "cloth_all": ("windyflag", cloth_xml(24, 16, ["s", "b", "t", "s2"],
                                         "<Force name=\"t\" type=\"TriangleStrain\"><Stiffness value=\"80\" /><limit value=\".9 1.1\" /></Force>"
                                         "<Force name=\"b\" type=\"Bend\"><Stiffness value=\"3\" /></Force>"
                                         "<Force name=\"s\" type=\"Spring\"><Stiffness value=\"11\" /></Force>"
                                         "<Force name=\"s2\" type=\"LinearTriangleStrain\"><Stiffness value=\"7\" /></Force>")),
    "cloth_errors": ("windyflag", cloth_xml(6, 4, ["nostiff", "limited", "bendnostiff", "unknown", "const", "ok"],
                                            "<Force name=\"nostiff\" type=\"TriangleStrain\"></Force>"
                                            "<Force name=\"limited\" type=\"Spring\"><Stiffness value=\"1\" /><limit value=\"0.5 2\" /></Force>"
                                            "<Force name=\"bendnostiff\" type=\"Bend\"></Force>"
                                            "<Force name=\"unknown\" type=\"NoSuchForce\"><Stiffness value=\"1\" /></Force>"
                                            "<Force name=\"const\" type=\"ConstForce\"></Force>"
                                            "<Force name=\"ok\" type=\"Spring\"><Stiffness value=\"2\" /></Force>")),
    "tets_all": ("bunnyexpand", tet_xml(["lin", "nh", "vol", "sv"],
                                        "<Force name=\"lin\" type=\"LinearTetStrain\"><Stiffness value=\"1000\" /><weight_scale value=\"2\" /></Force>"
                                        "<Force name=\"nh\" type=\"NeoHookeanTet\"><mu value=\"50\" /><lambda value=\"70\" /></Force>"
                                        "<Force name=\"vol\" type=\"VolPres\"><Stiffness value=\"9\" /><range_min value=\"0.9\" /><range_max value=\"1.2\" /></Force>"
                                        "<Force name=\"sv\" type=\"StVKTet\"><mu value=\"5\" /><lambda value=\"6\" /><max_iterations value=\"3\" /></Force>")),
    "tets_errors": ("bunnyexpand", tet_xml(["nostiff", "volhalf", "unknown", "const", "ok"],
                                           "<Force name=\"nostiff\" type=\"LinearTetStrain\"></Force>"
                                           "<Force name=\"volhalf\" type=\"VolPres\"><Stiffness value=\"9\" /><range_min value=\"0.9\" /></Force>"
                                           "<Force name=\"unknown\" type=\"Anisotropic\"><Stiffness value=\"1\" /></Force>"
                                           "<Force name=\"const\" type=\"ConstForce\"></Force>"
                                           "<Force name=\"ok\" type=\"LinearTetStrain\"><Stiffness value=\"3\" /></Force>")),
    # large enough that the reference's hinge scan shows (at width 120 x length 80, 95 800 elements, it takes 76 s against
    # 0.075 s batched; the test uses a quarter of that to stay short)
    "cloth_large": ("windyflag", cloth_xml(60, 40, ["t", "b"],
                                           "<Force name=\"t\" type=\"TriangleStrain\"><Stiffness value=\"80\" /></Force>"
                                           "<Force name=\"b\" type=\"Bend\"><Stiffness value=\"3\" /></Force>")),
}


def _need_runners():
    if not all(os.path.exists(r) for r in RUNNERS) or not os.path.exists(os.path.join(DROPIN, "scenes.tar")):
        pytest.skip("tests/dropin not built (needs the reference tree at build time: __graft_entry__.build())")


def _dump_both(xml, tmp_path, tag):
    out = []
    for k, runner in enumerate(RUNNERS):
        txt = str(tmp_path / f"{tag}.{k}.txt")
        r = subprocess.run([runner, "dump-forces", xml, txt], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        m = re.search(r"load ([0-9.]+) s, (\d+) force objects, (\d+) elements", r.stdout)
        assert m, r.stdout
        out.append((open(txt).read(), float(m.group(1)), int(m.group(2)), int(m.group(3))))
    return out


def _unpack(tmp_path, scene):
    with tarfile.open(os.path.join(DROPIN, "scenes.tar")) as tf:
        tf.extractall(tmp_path, members=[m for m in tf.getmembers() if m.name.startswith(scene + "/")], filter="data")
    return tmp_path / scene


@pytest.mark.parametrize("name", list(XML))
def test_batched_factory_builds_the_shipped_scenes_identically(name, tmp_path):
    _need_runners()
    d = _unpack(tmp_path, name)
    (ref, t_ref, o_ref, e_ref), (bat, t_bat, o_bat, e_bat) = _dump_both(str(d / XML[name]), tmp_path, name)
    print(f"{name}: {e_ref} elements; reference factory {o_ref} heap objects in {t_ref:.3f} s, batched {o_bat} objects in {t_bat:.3f} s")
    assert e_ref == e_bat and e_ref > 0
    assert ref == bat
    assert o_bat < 8


@pytest.mark.parametrize("name", list(SYNTHETIC))
def test_batched_factory_matches_over_the_xml_vocabulary(name, tmp_path):
    _need_runners()
    scene, xml = SYNTHETIC[name]
    d = _unpack(tmp_path, scene)
    path = d / f"{name}.xml"
    path.write_text(xml)
    (ref, t_ref, o_ref, e_ref), (bat, t_bat, o_bat, e_bat) = _dump_both(str(path), tmp_path, name)
    print(f"{name}: {e_ref} elements; reference factory {o_ref} heap objects in {t_ref:.3f} s, batched {o_bat} objects in {t_bat:.3f} s ({t_ref / max(t_bat, 1e-9):.0f} x)")
    assert e_ref == e_bat
    assert ref == bat
    if name.endswith("_errors"):
        assert o_bat == 1          # only the last, well-formed force produced anything
    if name == "cloth_large":
        assert e_ref > 20000
        assert t_bat < t_ref
