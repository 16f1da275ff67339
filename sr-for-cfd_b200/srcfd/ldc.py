"""Namespace mirror of PyCFD_ML_accelerated.py (lid-driven cavity + SR warm start)."""
from .kernels import (apply_bc_configured, copy_new_to_old, correct_velocity, linear_interpolation,  # noqa: F401
                      solve_momentum_quick, solve_momentum_upwind, solve_pressure, update_flux)
from .solver import (BoundaryCondition, BoundaryConditions, CFDSolver, FluidProperties, MeshParameters,  # noqa: F401
                     SolverSettings)
from .workflow import ldc_workflow as _wf

create_timestamped_output_dir = _wf.create_timestamped_output_dir
standardize_with_stats = _wf.standardize_with_stats
inverse_standardize = _wf.inverse_standardize
SuperResolutionAE = _wf.SuperResolutionAE
run_coarse_simulation = _wf.run_coarse_simulation
ml_super_resolution = _wf.ml_super_resolution
run_fine_simulation_with_ml_init = _wf.run_fine_simulation_with_ml_init
generate_coarse_mesh_solution = _wf.generate_coarse_mesh_solution
run_ml_accelerated_fine_simulation = _wf.run_ml_accelerated_fine_simulation
run_normal_simulation = _wf.run_normal_simulation
extract_centerlines = _wf.extract_centerlines
