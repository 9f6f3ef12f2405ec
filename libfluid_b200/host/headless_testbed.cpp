// Headless counterpart of the reference testbed's simulation thread (testbed/main.cpp:90-197): the same scene
// presets and per-step console diagnostics, driven through the unchanged fluid::simulation API -- but the step
// runs on the GPU.  Usage: headless_testbed [setup 0-3] [grid n] [steps]
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "fluid/simulation.h"

using fluid::vec3d;
using fluid::vec3s;

int main(int argc, char **argv) {
	int setup = argc > 1 ? std::atoi(argv[1]) : 0;
	std::size_t n = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 50;
	int steps = argc > 3 ? std::atoi(argv[3]) : 20;
	double s = static_cast<double>(n) / 50.0;

	fluid::simulation sim;
	sim.resize(vec3s(n, n, n));
	sim.grid_offset = vec3d();
	sim.cell_size = 1.0;
	sim.simulation_method = fluid::simulation::method::apic;
	sim.gravity = vec3d(0.0, -981.0, 0.0);
	sim.post_pressure_solve_callback = [](double dt, std::vector<double> &pressure, double residual, std::size_t iters) {
		double maxp = 0.0;
		for (double p : pressure) {
			maxp = p > maxp ? p : maxp;
		}
		std::cout << "  dt " << dt << "  iterations " << iters << (iters > 100 ? "  WARNING: large number of iterations" : "")
			<< "  residual " << residual << "  max pressure " << maxp << "\n";
	};
	switch (setup) {
	case 0: sim.seed_box(vec3d(15, 15, 15) * s, vec3d(20, 20, 20) * s); break;
	case 1: sim.seed_sphere(vec3d(25, 25, 25) * s, 15.0 * s); break;
	case 2:
		sim.seed_sphere(vec3d(25, 44, 25) * s, 5 * s);
		sim.seed_box(vec3d(0, 0, 0), vec3d(50, 15, 50) * s);
		break;
	default: sim.seed_box(vec3d(0, 0, 0), vec3d(10, 50, 50) * s); break;
	}
	sim.reset_space_hash();
	std::cout << "setup " << setup << ", grid " << n << "^3, " << sim.particles().size() << " particles\n";
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < steps; ++i) {
		sim.time_step();
		double maxv = 0.0, energy = 0.0;
		for (const auto &p : sim.particles()) {
			maxv = std::max(maxv, p.velocity.squared_length());
			energy += 0.5 * p.velocity.squared_length() - vec3d(0, -981.0, 0).y * p.position.y;
		}
		std::cout << "step " << i << "  max speed " << std::sqrt(maxv) << "  energy " << energy << "\n";
	}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cout << steps << " steps in " << sec << " s\n";
	return 0;
}
