"""VehicleParams — field-for-field the reference dataclass (multi_layer_planner_v3.py:29-39)."""
from dataclasses import dataclass


@dataclass
class VehicleParams:
    working_width: float = 3.2
    min_turn_radius: float = 8.0
    max_work_speed_kmh: float = 9.0
    max_headland_speed_kmh: float = 15.0
    headland_turn_speed_kmh: float = 4.0
    max_lateral_accel: float = 2.0
    max_longitudinal_accel: float = 1.5
    safety_factor: float = 0.85
