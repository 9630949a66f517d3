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
    with pytest.raises(NotImplementedError):
        _native_config({"render": {"splat": {"k_buffer_size": 16}}})

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
