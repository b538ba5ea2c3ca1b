"""Parameter surface: names/defaults of VARIABLES_TO_EXPOSE and the JSON name-matching rule."""
import numpy as np

from silver2_isaacsim_b200 import params as P
from silver2_isaacsim_b200 import workloads as W


def test_exposed_variables_match_reference_surface():
    # hydrodynamics_behavior.py:28-46 / README.md:128-141
    assert [k for k, _ in P.EXPOSED_VARIABLES] == [
        "waterDensity", "gravity", "xDimension", "yDimension", "zDimension", "linearDragCoefficient",
        "angularDragCoefficient", "linearDamping", "angularDamping", "linearAddedMassCoefficient",
        "angularAddedMassCoefficient", "liftCoefficient"]
    assert P.EXPOSED_DEFAULTS == {
        "waterDensity": 1025.0, "gravity": 9.81, "xDimension": 1.0, "yDimension": 1.0, "zDimension": 1.0,
        "linearDragCoefficient": 1.2, "angularDragCoefficient": 0.8, "linearDamping": 300.0,
        "angularDamping": 150.0, "linearAddedMassCoefficient": 0.05, "angularAddedMassCoefficient": 0.02,
        "liftCoefficient": 1.0}
    p = P.HydroParams()
    # wrapper ctor order (numba_hydrodynamics_wrapper.py:9-10), README default cube
    assert p.ctor_row() == [1, 1, 1, 1.2, 0.8, 300, 150, 1025, 9.81, 0.05, 0.02, 1.0]
    assert p.coeff_record(512.5) == [1, 1, 1, 1.2, 0.8, 300, 150, 0.05, 0.02, 1.0, 512.5]
    assert not p.set("noSuchAttribute", 1.0)


def test_part_matching_rule():
    cfg = P.load_config()
    assert P.match_part("Coxa_0", cfg) == "coxa"
    assert P.match_part("Femur_5", cfg) == "femur"
    assert P.match_part("TIBIA_3", cfg) == "tibia"
    assert P.match_part("Body", cfg) == "body"
    assert P.match_part("Obsea_Buoy", cfg) is None  # falls back to the exposed defaults
    assert P.match_part("my_body_part", {"parts": {}}) == "body"
    p, part = P.params_for_prim("Obsea_Buoy", cfg)
    assert part is None and p.xDimension == 1.0 and p.waterDensity == 1025.0
    p, part = P.params_for_prim("Tibia_2", cfg)
    assert part == "tibia" and (p.xDimension, p.yDimension, p.zDimension) == (0.06, 0.09, 0.06)
    assert p.linearDamping == 20.0 and p.liftCoefficient == 0.1


def test_hexapod_table():
    table, slot_type, rho, g = P.hexapod_table()
    assert table.shape == (4, 11) and slot_type.shape == (19,)
    assert list(slot_type) == [0] + [1] * 6 + [2] * 6 + [3] * 6
    assert (rho, g) == (1025.0, 9.81)
    assert list(table[:, 10]) == [18.0, 0.45, 0.75, 0.8]  # masses, SURVEY.md Appendix D
    rows = P.coeff_to_ctor_rows(table, rho, g)
    assert list(rows[0]) == [0.26, 0.26, 0.30, 1.2, 0.8, 300, 150, 1025, 9.81, 0.2, 0.1, 0.5]


def test_workloads_are_seeded_and_shaped():
    a, b = W.hexapod_envs(8), W.hexapod_envs(8)
    assert a.n == 152 and a.bodies_per_robot == 19 and (a.pos == b.pos).all()
    assert a.pos.dtype == np.float32 and a.transforms().shape == (152, 7) and a.velocities().shape == (152, 6)
    c = W.heterogeneous_boxes(1000)
    assert c.coeff.shape == (1000, 11) and c.table is None
    assert np.allclose(np.linalg.norm(c.quat_xyzw, axis=1), 1, atol=1e-6)
    d = W.sharded_robots(4)
    assert d.coeff.shape == (76, 11) and d.bodies_per_robot == 19
    e = W.uniform_small_batch()
    assert e.n == 1024 and e.table.shape == (1, 11)
    assert W.readme_buoy().meta["steps"] == 10000


def test_trace_and_rtf_utilities(tmp_path):
    """CSV column order of the reference's LogVelocity and the RTF arithmetic of BenchmarkRtf."""
    import csv
    from silver2_isaacsim_b200.trace import LOG_HEADER, RtfMeter, VelocityTrace

    path = tmp_path / "velocity_log.csv"
    tr = VelocityTrace(str(path))
    tr.sample(np.array([[1.0, 2.0, 3.0]]), np.array([[4.0, 5.0, 6.0]]), np.array([[7.0, 8.0, 9.0]]), timestamp="t0")
    rows = list(csv.reader(open(path)))
    assert rows[0] == LOG_HEADER and rows[0][1:4] == ["z_position", "linear_velocity_z", "angular_velocity_z"]
    assert rows[1] == ["t0", "3.0", "6.0", "9.0", "1.0", "4.0", "7.0", "2.0", "5.0", "8.0"]
    lines = []
    m = RtfMeter(report_every=600, printer=lines.append)
    m.on_physics_step(1 / 120, n_steps=1200)
    r = m.report()
    assert r["steps"] == 1200 and abs(r["sim_time_s"] - 10.0) < 1e-9 and r["rtf"] > 0 and len(lines) == 1
