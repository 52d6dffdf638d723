"""Calibration data used by the parity configurations (values are DATA taken from the
reference's calibration stores):

* RIG_2222: last block (2021-11-17-10-25-21) of /root/reference/2222/cameraparaout_1.txt --
  shared K for 480x270 inputs, four R rows, warped_image_scale.  Coherent with 2222/1-4.png.
* RIG_424: last block (2022-04-24-11-24-08) of /root/reference/cfg/424camcfg/cameraparaout_1.txt --
  per-camera K (640x360 inputs) + R, scale.
* CAM_LIJING_390_FOV60_1920: cfg/cameras.yaml:80-88 (undistort K / distortion / crop rect).
"""
import numpy as np

RIG_2222 = dict(
    base_width=480,
    K=[391.047, 0, 240, 0, 391.047, 135, 0, 0, 1],
    R=[
        [0.999932, 0.0115248, 0.0018935, -0.0115405, 0.999897, 0.00851308, -0.00179527, -0.00853435, 0.999962],
        [0.732011, 0.0131828, -0.681166, -0.00411094, 0.99988, 0.0149332, 0.681281, -0.00813103, 0.731977],
        [-0.0130466, 0.00948596, -0.99987, 0.0120215, 0.999884, 0.00932925, 0.999843, -0.0118982, -0.0131591],
        [-0.724781, 0.0204802, -0.688675, 0.0178955, 0.999781, 0.0108983, 0.688747, -0.00442533, -0.724989],
    ],
    scale=381.719,
)

RIG_424 = dict(
    base_width=640,
    K=[
        [5093.54, 0, 320, 0, 5093.54, 180, 0, 0, 1],
        [5062.47, 0, 320, 0, 5062.47, 180, 0, 0, 1],
        [4976.91, 0, 320, 0, 4976.91, 180, 0, 0, 1],
        [4947.03, 0, 320, 0, 4947.03, 180, 0, 0, 1],
    ],
    R=[
        [0.990251, -0.0289489, -0.13625, -0.0162866, 0.947392, -0.31966, 0.138336, 0.318763, 0.937685],
        [0.998381, -0.0367373, -0.043432, 0.0209781, 0.947459, -0.319187, 0.0528762, 0.317759, 0.946696],
        [0.998971, 0.0103404, 0.04415, 0.00423479, 0.948122, -0.317878, -0.0451466, 0.317738, 0.947103],
        [0.989246, 0.05499, 0.135532, -0.0091451, 0.948075, -0.317916, -0.145977, 0.313257, 0.938382],
    ],
    scale=5019.69,
)

CAM_LIJING_390_FOV60_1920 = dict(
    K=[2.075765787574657e+03, 0, 9.479666200437899e+02, 0, 2.066538110898970e+03, 5.677805443267157e+02, 0, 0, 1],
    distorParams=[-0.6183, 0.3355, 0, 0],
    rect=[69, 103, 1782, 889],
    size=(1920, 1080),
)


def rig(name, width):
    """-> (Ks, Rs, scale) as float32, re-targeted to frames `width` pixels wide the way the
    reference data is meant to be used (multiply K rows 0-1 and the scale)."""
    d = {"2222": RIG_2222, "424": RIG_424}[name]
    n = len(d["R"])
    Ks = d["K"] if isinstance(d["K"][0], list) else [d["K"]] * n
    f = np.float32(width / d["base_width"])
    out_k = []
    for k in Ks:
        K = np.array(k, np.float32).reshape(3, 3)
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        out_k.append(K)
    Rs = [np.array(r, np.float32).reshape(3, 3) for r in d["R"]]
    return out_k, Rs, float(np.float32(d["scale"]) * f)


def ring(n, width, height, hfov_deg, step_deg, focal=None):
    """Synthetic cylindrical ring of BASELINE config 4: R_i = R_y((i-(n-1)/2)*step)."""
    f = np.float32(focal if focal is not None else (width / 2.0) / np.tan(np.radians(hfov_deg) / 2.0))
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], np.float32)
    Rs = []
    for i in range(n):
        a = np.radians((i - (n - 1) / 2.0) * step_deg)
        c, s = np.cos(a), np.sin(a)
        Rs.append(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32))
    return [K.copy() for _ in range(n)], Rs, float(f)
