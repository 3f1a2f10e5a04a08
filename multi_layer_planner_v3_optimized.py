"""Module-name shim: the reference's callers import the planner under this name
(multi_field_planner.py:24, test/test_v37_complete.py:15, test/test_multi-layer_planner_v3.py:7-9).
Everything resolves to the CUDA-backed drop-in in field_coverage_path_planning_b200."""
from field_coverage_path_planning_b200 import (  # noqa: F401
    TwoLayerPathPlannerV35, TwoLayerPathPlannerV36, TwoLayerPathPlannerV37, TwoLayerPlannerV35,
    TwoLayerPlannerV36, VehicleParams)
