"""Host-side logic of the reference-facing mirror (no GPU): config -> native constants, pose conventions."""
import math

import numpy as np
import pytest

import scenes


def test_native_config_from_reference_yaml_keys():
    from threedgut_tracer.tracer import _native_config

    conf = {"render": {"particle_kernel_degree": 2, "particle_kernel_min_response": 0.0113, "particle_kernel_min_alpha": 1 / 255,
                       "particle_kernel_max_alpha": 0.99, "min_transmittance": 1e-4, "enable_kernel_timings": True,
                       "splat": {"ut_alpha": 1.0, "ut_beta": 2.0, "ut_kappa": 0.0, "ut_in_image_margin_factor": 0.1, "rect_bounding": True,
                                 "tight_opacity_bounding": True, "tile_based_culling": True, "k_buffer_size": 0, "global_z_order": True}}}
    cfg = _native_config(conf)
    assert cfg.kernel_degree == 2 and cfg.enable_timings == 1
    assert abs(cfg.ut_delta - math.sqrt(3.0)) < 1e-6
    assert _native_config({"render": {"splat": {"k_buffer_size": 16}}}).k_buffer_size == 16  # sorted 3DGUT (configs/paper/3dgut/sorted_*.yaml)
    with pytest.raises(NotImplementedError):
        _native_config({"render": {"splat": {"k_buffer_size": 64}}})

    class Obj:  # attribute-style (OmegaConf-like) access works too
        class render:
            particle_kernel_degree = 4
            min_transmittance = 0.001

    cfg = _native_config(Obj)
    assert cfg.kernel_degree == 4 and abs(cfg.min_transmittance - 1e-3) < 1e-9


def test_pose_convention_matches_reference_tracer():
    """C2W -> [t, q.xyzw] world->sensor, the convention of threedgut_tracer/tracer.py:404-423."""
    from threedgut_tracer.tracer import Tracer

    sc = scenes.scene_c1()
    for i in range(5):
        c2w = sc.camera(i, 5)
        pose = Tracer._pose_from_c2w(np.asarray(c2w, np.float32))
        assert np.allclose(pose, scenes.pose7_from_c2w(c2w), atol=1e-6)
        x, y, z, w = pose[3:]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        W2C = np.linalg.inv(c2w)
        assert np.allclose(R, W2C[:3, :3], atol=1e-5) and np.allclose(pose[:3], W2C[:3, 3], atol=1e-5)


def test_camera_parameters_from_batch_intrinsics():
    torch = pytest.importorskip("torch")
    from threedgut_tracer.tracer import Tracer

    class B:
        T_to_world = torch.eye(4)[None]
        T_to_world_end = None
        rays_in_world_space = False
        intrinsics = [1111.0, 1111.0, 400.0, 400.0]

    sensor, poses = Tracer._create_camera_parameters(B)
    assert list(sensor.resolution) == [800, 800]
    assert np.allclose(sensor.focal_length, [1111.0, 1111.0], rtol=1e-5) and np.allclose(sensor.principal_point, [400, 400])
    assert np.allclose(poses.T_world_sensors[0], [0, 0, 0, 0, 0, 0, 1], atol=1e-7)


def test_camera_factories_fill_the_native_camera():
    """bindings.cpp:50-101 factories -> gutb200_camera fields (model, rolling_shutter, fisheye / f-theta parameters)."""
    from threedgut_tracer.tracer import (PolynomialType, ShutterType, SplatRaster, Tracer, fromFThetaCameraModelParameters,
                                         fromOpenCVFisheyeCameraModelParameters, fromOpenCVPinholeCameraModelParameters)

    pose = np.array([0, 0, 0, 0, 0, 0, 1], np.float32)
    pin = fromOpenCVPinholeCameraModelParameters(np.array([64, 48]), ShutterType.ROLLING_LEFT_TO_RIGHT, [32, 24], [50, 51], np.arange(6) * 0.01,
                                                 [0.001, 0.002], [0.1, 0.2, 0.3, 0.4])
    cam = SplatRaster._camera(pin, pose, pose, 64, 48)
    assert cam.model == 0 and cam.rolling_shutter == 2 and abs(cam.radial[5] - 0.05) < 1e-7 and abs(cam.thin_prism[3] - 0.4) < 1e-7
    glob = fromOpenCVPinholeCameraModelParameters(np.array([64, 48]), ShutterType.GLOBAL, [32, 24], [50, 51], np.zeros(6), np.zeros(2), np.zeros(4))
    assert SplatRaster._camera(glob, pose, pose, 64, 48).rolling_shutter == 0
    fe = fromOpenCVFisheyeCameraModelParameters(np.array([64, 48]), ShutterType.ROLLING_TOP_TO_BOTTOM, [32, 24], [40, 40], [0.1, 0.2, 0.3, 0.4], 1.5)
    cam = SplatRaster._camera(fe, pose, pose, 64, 48)
    assert cam.model == 1 and cam.rolling_shutter == 1 and abs(cam.max_angle - 1.5) < 1e-7 and abs(cam.radial[3] - 0.4) < 1e-7 and cam.radial[4] == 0
    ft = fromFThetaCameraModelParameters(np.array([64, 48]), ShutterType.GLOBAL, [31.5, 23.5], PolynomialType.ANGLE_TO_PIXELDIST,
                                         [0, 0.02, 0, 0, 0, 0], [0, 50, 0, 0, 0, 0], 1.2, [1.0, 0.0, 0.0])
    cam = SplatRaster._camera(ft, pose, pose, 64, 48)
    assert cam.model == 2 and cam.ftheta_reference_poly == 1 and abs(cam.ftheta_fw[1] - 50) < 1e-6 and abs(cam.ftheta_cde[0] - 1.0) < 1e-7
    assert abs(cam.principal[0] - 31.5) < 1e-6

    class Batch:  # the intrinsics dictionaries of threedgrut/datasets/protocols.py:24-43 as the tracer receives them
        rays_in_world_space = False
        T_to_world = np.eye(4, dtype=np.float32)[None]
        T_to_world_end = None
        intrinsics = None
        intrinsics_OpenCVPinholeCameraModelParameters = None
        intrinsics_OpenCVFisheyeCameraModelParameters = dict(resolution=np.array([64, 48]), shutter_type="ROLLING_BOTTOM_TO_TOP",
                                                             principal_point=np.array([32, 24], np.float32), focal_length=np.array([40, 40], np.float32),
                                                             radial_coeffs=np.zeros(4, np.float32), max_angle=1.0)

    sensor, poses = Tracer._create_camera_parameters(Batch)
    assert sensor.model == 1 and sensor.shutter_type == ShutterType.ROLLING_BOTTOM_TO_TOP
    assert SplatRaster._camera(sensor, poses.T_world_sensors[0], poses.T_world_sensors[1], 64, 48).rolling_shutter == 3


def test_densify_step_schedule_and_optimizer_groups_are_consistent():
    import densify

    assert densify.GROUPS == ("positions", "density", "rotation", "scale", "features_albedo", "features_specular")
    c = densify.DensifyConfig()
    fired = [s for s in range(1, 2001) if densify.check_step_condition(s, c.densify_start, c.densify_end, c.densify_frequency)]
    assert fired == [600, 900, 1200, 1500, 1800]  # configs/strategy/gs.yaml: start 500, every 300


def test_camera_struct_is_reused_only_for_the_same_unchanged_frame():
    """SplatRaster._camera_cached: trace_bwd re-uses trace's native.Camera when the same sensor / pose objects come back unchanged; a new
    pose object, a pose mutated in place, another resolution or a tensor pose (the reference's binding takes tensors) build a fresh struct."""
    import torch
    from threedgut_tracer.tracer import ShutterType, SplatRaster, fromOpenCVPinholeCameraModelParameters

    raster = object.__new__(SplatRaster)  # no CUDA context needed for the host-side struct
    raster._last_camera = None
    sensor = fromOpenCVPinholeCameraModelParameters(np.array([64, 48]), ShutterType.GLOBAL, [32, 24], [50, 51], np.zeros(6), np.zeros(2), np.zeros(4))
    pose = np.array([0.1, 0.2, 0.3, 0, 0, 0, 1], np.float32)
    a = raster._camera_cached(sensor, pose, pose, 64, 48)
    assert raster._camera_cached(sensor, pose, pose, 64, 48) is a
    assert raster._camera_cached(sensor, pose.copy(), pose, 64, 48) is not a            # another object
    b = raster._camera_cached(sensor, pose, pose, 64, 48)
    pose[0] = 0.5                                                                         # same object, mutated in place
    c = raster._camera_cached(sensor, pose, pose, 64, 48)
    assert c is not b and abs(c.pose_start[0] - 0.5) < 1e-7 and abs(b.pose_start[0] - 0.1) < 1e-7
    assert raster._camera_cached(sensor, pose, pose, 32, 48) is not c                    # another resolution
    t = torch.tensor([0.0, 0, 0, 0, 0, 0, 1])
    d = raster._camera_cached(sensor, t, t, 64, 48)
    assert raster._camera_cached(sensor, t, t, 64, 48) is d and abs(d.pose_start[6] - 1.0) < 1e-7
