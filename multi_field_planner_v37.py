"""Module-name shim for `from multi_field_planner_v37 import TSPSolver` (multi_field_planner.py:176,
multi_vehicle_planner.py:131): the reference imports this module but does not ship it.  2-opt on the GPU
(build-defined, see field_coverage_path_planning_b200/tsp.py)."""
from field_coverage_path_planning_b200.tsp import TSPSolver  # noqa: F401
