"""Host-side mirror of the reference interface: yaml init, calibration-file parsing, errors."""
import os

import numpy as np

import panob200
from golden import calib

st_mod = panob200.pkg.stitcher

NEW_LAYOUT = """2022-04-24-11-20-00:
1,0,2,0,1,3,0,0,1,1,0,0,0,1,0,0,0,1,
1,0,2,0,1,3,0,0,1,1,0,0,0,1,0,0,0,1,
77.5
2022-04-24-11-24-08:
5093.54,0,320,0,5093.54,180,0,0,1,0.990251,-0.0289489,-0.13625,-0.0162866,0.947392,-0.31966,0.138336,0.318763,0.937685,
5062.47,0,320,0,5062.47,180,0,0,1,0.998381,-0.0367373,-0.043432,0.0209781,0.947459,-0.319187,0.0528762,0.317759,0.946696,
5019.69
"""
OLD_LAYOUT = """2021-11-17-10-25-21:
391.047,0,240,0,391.047,135,0,0,1,
0.999932,0.0115248,0.0018935,-0.0115405,0.999897,0.00851308,-0.00179527,-0.00853435,0.999962,
0.732011,0.0131828,-0.681166,-0.00411094,0.99988,0.0149332,0.681281,-0.00813103,0.731977,
381.719
"""


def test_parse_camera_params_layouts(tmp_path):
    p = tmp_path / "cameraparaout_0.txt"
    p.write_text(NEW_LAYOUT)
    Ks, Rs, sc = st_mod.parse_camera_params_file(str(p), 2)
    assert len(Ks) == 2 and Ks[0][0, 0] == np.float32(5093.54) and Rs[1][2, 2] == np.float32(0.946696)
    assert sc == float(np.float32(5019.69))
    p.write_text(OLD_LAYOUT)
    Ks, Rs, sc = st_mod.parse_camera_params_file(str(p))
    assert len(Rs) == 2 and Ks[1][0, 2] == 240 and sc == float(np.float32(381.719))
    Ks2, sc2 = st_mod.scale_intrinsics(Ks, sc, 4)
    assert Ks2[0][0, 0] == np.float32(391.047) * np.float32(4) and Ks2[0][2, 2] == 1


def test_yaml_init_matches_structure(tmp_path):
    cams = tmp_path / "cameras.yaml"
    cams.write_text("""
structures:
 -
  vendor: lijing
  sensor: imx390
  sttype: 4cam-black
  undistor: true
  fov: 120
  inputsz: 720
  params:
   -
    cams: [400,0,360,0,400,202,0,0,1, 1,0,0,0,1,0,0,0,1,
           410,0,360,0,410,202,0,0,1, 0.8,0,0.6,0,1,0,-0.6,0,0.8,
           395.5]
    cut: [3, 58, 900, 300]
""")
    cfg = tmp_path / "stitcher.yaml"
    cfg.write_text("""
vendor: lijing
sensor: imx390
sttype: 4cam-black
undistor: true
fov: 120
outPutWidth: 720
outPutHeight: 405
num_images: 2
stitcherMatchConf: 0.3
stitcherAdjusterConf: 0.7
stitcherBlenderStrength: 1
stitcherCameraExThres: 30e2
stitcherCameraInThres: 500e2
camcfgpath: "%s/"
initMode: 2
cameraparams: "%s"
""" % (tmp_path, cams))
    panob200.ocvStitcher.m_num = 0
    s = panob200.ocvStitcher()
    assert s.init(str(cfg)) == st_mod.RET_OK
    c = s.m_cfg
    assert (c.width, c.height, c.num_images, c.initMode) == (720, 405, 2, 2)
    assert c.Ks[1][0, 0] == 410 and c.Rs[1][0, 2] == np.float32(0.6) and c.cut == [3, 58, 900, 300]
    assert c.warped_image_scale == float(np.float32(395.5))
    # unmatched rig -> RET_ERR like the reference (defaultCamParams empty, :428-432)
    cfg.write_text(cfg.read_text().replace("fov: 120", "fov: 60"))
    s2 = panob200.ocvStitcher()
    assert s2.init(str(cfg)) == st_mod.RET_ERR
    assert panob200.ocvStitcher().init(str(tmp_path / "missing.yaml")) == st_mod.RET_ERR


def test_blender_rule_follows_reference():
    """:1188-1195: blend_width = sqrt(area)*strength/100; <1 -> NO; bands = ceil(log2(bw)) - 1."""
    Ks, Rs, sc = calib.rig("2222", 1920)
    for strength, want in ((1, 4), (3, 6), (5, 6)):   # 4/6/6 at the config-1 dst size
        s = panob200.ocvStitcher(panob200.StitcherConfig(width=1920, height=1080, num_images=4, Ks=Ks, Rs=Rs,
                                                         warped_image_scale=sc, blendStrength=strength))
        kind, nb, _ = s._blender_choice(5336, 1025)
        assert (kind, nb) == ("multiband", want)
    s = panob200.ocvStitcher(panob200.StitcherConfig(blendStrength=0))
    assert s._blender_choice(5336, 1025)[0] == "no"


def test_calibration_rejects_feature_matching_mode():
    s = panob200.ocvStitcher(panob200.StitcherConfig(initMode=1, num_images=2))
    assert s.calibration([]) == st_mod.RET_ERR and "out of scope" in s.last_error
