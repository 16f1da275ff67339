"""Namespace mirror of bfs_ml_accelerated.py (backward-facing step + SR warm start)."""
from .kernels import (apply_bc_configured, copy_new_to_old, linear_interpolation, solve_momentum_quick,  # noqa: F401
                      solve_momentum_upwind, solve_pressure, under_relax_field, update_flux)
from .kernels import correct_velocity as _correct_velocity
from .solver import BFSBoundaryConditions as BoundaryConditions  # noqa: F401
from .solver import BFSCFDSolver as CFDSolver  # noqa: F401
from .solver import BFSSolverSettings as SolverSettings  # noqa: F401
from .solver import BoundaryCondition, FluidProperties  # noqa: F401
from .solver import MeshParameters as _Mesh
from .workflow import bfs_workflow as _wf


class MeshParameters(_Mesh):
    """bfs_ml_accelerated.py:180-189: default domain 10 x 3."""

    def __init__(self, nx: int = 100, ny: int = 100, lx: float = 10.0, ly: float = 3.0):
        super().__init__(nx, ny, lx, ly)


def correct_velocity(Var, VarOld, dt, rho, Nx, Ny, dx, dy):
    """bfs_ml_accelerated.py:445-464: returns (res_u, res_v, res_p)."""
    return _correct_velocity(Var, VarOld, dt, rho, Nx, Ny, dx, dy)


create_timestamped_output_dir = _wf.create_timestamped_output_dir
standardize_with_stats = _wf.standardize_with_stats
inverse_standardize = _wf.inverse_standardize
reshape_rectangular_to_square = _wf.reshape_rectangular_to_square
reshape_square_to_rectangular = _wf.reshape_square_to_rectangular
SuperResolutionAE = _wf.SuperResolutionAE
run_coarse_simulation = _wf.run_coarse_simulation
ml_super_resolution = _wf.ml_super_resolution
run_fine_simulation_with_ml_init = _wf.run_fine_simulation_with_ml_init
generate_coarse_mesh_solution = _wf.generate_coarse_mesh_solution
run_ml_accelerated_fine_simulation = _wf.run_ml_accelerated_fine_simulation
run_normal_simulation = _wf.run_normal_simulation
extract_centerlines = _wf.extract_centerlines
