"""Module-name shim for `from genetic_algorithm_solver import GeneticAlgorithmSolver, GAConfig`
(multi_field_planner.py:26, multi_vehicle_planner.py:119): GPU tour-length fitness."""
from field_coverage_path_planning_b200 import GAConfig, GeneticAlgorithmSolver  # noqa: F401
