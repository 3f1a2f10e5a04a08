"""Module-name shim for `from multi_vehicle_planner import MultiVehiclePlanner, MultiVehicleRoute`
(multi_field_planner.py:26): KMeans split (Lloyd iterations on the GPU) + one device GA per vehicle."""
from field_coverage_path_planning_b200.multi_vehicle import (  # noqa: F401
    MultiVehiclePlanner, MultiVehicleRoute, VehicleRoute)
