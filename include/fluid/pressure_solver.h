#pragma once
// fluid::pressure_solver -- mirror of the reference class (include/fluid/pressure_solver.h:13-98).  solve() and
// apply_pressure() run on the device through lfk_pressure_solve / lfk_apply_pressure; the reference's sequential
// MIC(0) preconditioner is replaced by a GPU-parallel multigrid V-cycle, so `tau` and `sigma` are accepted for
// source compatibility but unused, and iteration counts differ.  Convergence: max |r_i| < tolerance in the
// reference's scaling of A and b (stricter than, and implying, the reference's one-sided max r_i < tolerance).
#include <tuple>
#include <vector>

#include "mac_grid.h"
#include "math/vec.h"
#include "simulation.h"

namespace fluid {
	class pressure_solver {
	public:
		struct cell_data {
			cell_data() : nonsolid_neighbors(0), fluid_xpos(0), fluid_ypos(0), fluid_zpos(0) {
			}
			std::size_t nonsolid_neighbors : 3, fluid_xpos : 1, fluid_ypos : 1, fluid_zpos : 1;
		};

		/// fluid_cells must be the cells that hold particles, in ascending raw order (what time_step passes,
		/// reference src/simulation.cpp:83-98); the device derives the same list from its sorted-cell table.
		explicit pressure_solver(simulation &sim, const std::vector<vec3s> &fluid_cells);

		/// returns (pressure per fluid cell, residual, iterations)
		[[nodiscard]] std::tuple<std::vector<double>, double, std::size_t> solve(double dt);
		void apply_pressure(double dt, const std::vector<double> &pressure) const;

		double tau = 0.97, sigma = 0.25, tolerance = 1e-6;
		std::size_t max_iterations = 200;
	protected:
		const std::vector<vec3s> &_fluid_cells;
		simulation &_sim;
	};
}
