"""Module-name shim for `from multi_field_planner import MultiFieldPlannerV38` (the reference's
test/test_multi_field_v38.py): distance / connection matrices, GA ordering and plan lengths on the GPU."""
from field_coverage_path_planning_b200.multi_field import (  # noqa: F401
    Connection, FieldData, MultiFieldPlannerV38, OptimizedRoute)
from field_coverage_path_planning_b200.multi_vehicle import MultiVehiclePlanner, MultiVehicleRoute  # noqa: F401,E402
