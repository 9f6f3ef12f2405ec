// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Phase-replay driver around the UNMODIFIED reference library (lukedan/libfluid).  It is compiled
// together with the reference's own three hot-path sources where they lie under /root/reference
// (see oracle/Makefile) into oracle/_ref/libfluid_ref.so, and exposes every private phase of
// fluid::simulation::time_step (reference src/simulation.cpp:43-125) and of
// fluid::pressure_solver::solve (reference src/pressure_solver.cpp:19-71) through a flat C ABI so
// that tests/ and bench.py's cpu_baseline leg can (a) pin the C restatement in oracle/fluid_oracle.c
// and (b) check the CUDA path stage by stage.  No reference source is copied here: the private
// members are reached with the `#define private public` trick around the two includes.
#include <cstring>
#include <cstdint>
#include <memory>
#include <vector>
#include <tuple>
#include <algorithm>

#define private public
#define protected public
#include "fluid/simulation.h"
#include "fluid/pressure_solver.h"
#include "fluid/mesher.h"
#include "fluid/voxelizer.h"
#include "fluid/data_structures/obstacle.h"
#undef private
#undef protected

using fluid::vec3d;
using fluid::vec3s;
using sim_t = fluid::simulation;
using cell_t = fluid::mac_grid::cell;

namespace {
	struct ref_ctx {
		sim_t sim;
		std::vector<vec3s> fluid_cells;
		std::unique_ptr<fluid::pressure_solver> solver;
		std::vector<double> pressure, b;
		double residual = 0.0;
		std::size_t iters = 0;
	};
	static_assert(sizeof(sim_t::particle) == 152, "particle AoS layout");
	static_assert(sizeof(cell_t) == 32, "cell AoS layout");
}

extern "C" {
	void *ref_create(std::size_t nx, std::size_t ny, std::size_t nz, double h) {
		auto *c = new ref_ctx();
		c->sim.resize(vec3s(nx, ny, nz));
		c->sim.cell_size = h;
		c->sim.reset_space_hash();
		return c;
	}
	void ref_destroy(void *p) {
		delete static_cast<ref_ctx*>(p);
	}
	// method: 0 pic, 1 flip_blend, 2 apic (reference include/fluid/simulation.h:44-48)
	void ref_set_params(
		void *p, const double *offset, const double *gravity, int method, double blend,
		double density, double skin, double stiffness, std::size_t extrap_iters, double cfl_number
	) {
		sim_t &s = static_cast<ref_ctx*>(p)->sim;
		s.grid_offset = vec3d(offset[0], offset[1], offset[2]);
		s.gravity = vec3d(gravity[0], gravity[1], gravity[2]);
		s.simulation_method = static_cast<sim_t::method>(method);
		s.blending_factor = blend;
		s.density = density;
		s.boundary_skin_width = skin;
		s.correction_stiffness = stiffness;
		s.velocity_extrapolation_iterations = extrap_iters;
		s.cfl_number = cfl_number;
	}
	void ref_seed_box(void *p, const double *start, const double *size, const double *vel, std::size_t dens) {
		static_cast<ref_ctx*>(p)->sim.seed_box(
			vec3d(start[0], start[1], start[2]), vec3d(size[0], size[1], size[2]),
			vec3d(vel[0], vel[1], vel[2]), dens
		);
	}
	void ref_seed_sphere(void *p, const double *center, double radius, const double *vel, std::size_t dens) {
		static_cast<ref_ctx*>(p)->sim.seed_sphere(
			vec3d(center[0], center[1], center[2]), radius, vec3d(vel[0], vel[1], vel[2]), dens
		);
	}
	void ref_add_source(
		void *p, const std::size_t *cells_xyz, std::size_t ncells, const double *vel, std::size_t dens, int coerce
	) {
		auto src = std::make_unique<fluid::source>();
		for (std::size_t i = 0; i < ncells; ++i) {
			src->cells.emplace_back(cells_xyz[3 * i], cells_xyz[3 * i + 1], cells_xyz[3 * i + 2]);
		}
		src->velocity = vec3d(vel[0], vel[1], vel[2]);
		src->target_density_cubic_root = dens;
		src->coerce_velocity = coerce != 0;
		static_cast<ref_ctx*>(p)->sim.sources.emplace_back(std::move(src));
	}

	std::size_t ref_num_particles(void *p) {
		return static_cast<ref_ctx*>(p)->sim.particles().size();
	}
	void ref_get_particles(void *p, void *out152) {
		auto &v = static_cast<ref_ctx*>(p)->sim.particles();
		std::memcpy(out152, v.data(), v.size() * sizeof(sim_t::particle));
	}
	void ref_set_particles(void *p, const void *in152, std::size_t n) {
		auto &v = static_cast<ref_ctx*>(p)->sim.particles();
		v.resize(n);
		std::memcpy(static_cast<void*>(v.data()), in152, n * sizeof(sim_t::particle));
	}
	void ref_get_cells(void *p, void *out32) {
		auto &g = static_cast<ref_ctx*>(p)->sim.grid().grid();
		std::memcpy(out32, &g[0], g.get_array_size(g.get_size()) * sizeof(cell_t));
	}
	void ref_set_cells(void *p, const void *in32) {
		auto &g = static_cast<ref_ctx*>(p)->sim.grid().grid();
		std::memcpy(static_cast<void*>(&g[0]), in32, g.get_array_size(g.get_size()) * sizeof(cell_t));
	}
	void ref_get_old_cells(void *p, void *out32) {
		auto &g = static_cast<ref_ctx*>(p)->sim._old_grid.grid();
		std::memcpy(out32, &g[0], g.get_array_size(g.get_size()) * sizeof(cell_t));
	}
	void ref_set_old_cells(void *p, const void *in32) {
		auto *c = static_cast<ref_ctx*>(p);
		if (c->sim._old_grid.grid().get_size().x != c->sim._grid.grid().get_size().x ||
			c->sim._old_grid.grid().get_size().y != c->sim._grid.grid().get_size().y ||
			c->sim._old_grid.grid().get_size().z != c->sim._grid.grid().get_size().z) {
			c->sim._old_grid = c->sim._grid;
		}
		auto &g = c->sim._old_grid.grid();
		std::memcpy(static_cast<void*>(&g[0]), in32, g.get_array_size(g.get_size()) * sizeof(cell_t));
	}
	// space hash table {begin,count} per cell, 16 B each (reference include/fluid/simulation.h:193-198)
	void ref_get_space_hash(void *p, std::size_t *out_begin_count) {
		auto &g = static_cast<ref_ctx*>(p)->sim._space_hash;
		std::size_t n = g.get_array_size(g.get_size());
		for (std::size_t i = 0; i < n; ++i) {
			out_begin_count[2 * i] = g[i].begin;
			out_begin_count[2 * i + 1] = g[i].count;
		}
	}
	std::size_t ref_num_fluid_cells(void *p) {
		return static_cast<ref_ctx*>(p)->sim._fluid_cells.size();
	}
	void ref_get_fluid_cells(void *p, std::size_t *out) {
		auto &v = static_cast<ref_ctx*>(p)->sim._fluid_cells;
		std::copy(v.begin(), v.end(), out);
	}

	// ---- phases of time_step, in the order of reference src/simulation.cpp:43-125 ----
	void ref_update_and_hash(void *p) {
		static_cast<ref_ctx*>(p)->sim.update_and_hash_particles();
	}
	void ref_hash(void *p) {
		static_cast<ref_ctx*>(p)->sim.hash_particles();
	}
	void ref_reset_space_hash(void *p) {
		static_cast<ref_ctx*>(p)->sim.reset_space_hash();
	}
	void ref_advect(void *p, double dt) {
		static_cast<ref_ctx*>(p)->sim._advect_particles(dt);
	}
	void ref_collide(void *p) {
		static_cast<ref_ctx*>(p)->sim._detect_collisions();
	}
	void ref_save_old_positions(void *p) {
		for (auto &q : static_cast<ref_ctx*>(p)->sim.particles()) {
			q.old_position = q.position;
		}
	}
	void ref_update_sources(void *p) {
		static_cast<ref_ctx*>(p)->sim._update_sources();
	}
	void ref_p2g(void *p) {
		static_cast<ref_ctx*>(p)->sim._transfer_to_grid();
	}
	void ref_gravity(void *p, double dt) {
		sim_t &s = static_cast<ref_ctx*>(p)->sim;
		auto &g = s.grid().grid();
		std::size_t n = g.get_array_size(g.get_size());
		for (std::size_t i = 0; i < n; ++i) {
			g[i].velocities_posface += s.gravity * dt;
		}
	}
	// builds the solver exactly as time_step does (src/simulation.cpp:83-99); tolerance / max_iterations are
	// public members of the solver and may be overridden for large grids (negative / zero = keep default)
	void ref_solver_setup(void *p, double tolerance, std::size_t max_iterations) {
		auto *c = static_cast<ref_ctx*>(p);
		c->fluid_cells.clear();
		for (std::size_t raw : c->sim._fluid_cells) {
			c->fluid_cells.emplace_back(c->sim.grid().grid().index_from_raw(raw));
		}
		c->solver = std::make_unique<fluid::pressure_solver>(c->sim, c->fluid_cells);
		if (tolerance > 0.0) {
			c->solver->tolerance = tolerance;
		}
		if (max_iterations > 0) {
			c->solver->max_iterations = max_iterations;
		}
	}
	void ref_solve(void *p, double dt, double *residual, std::size_t *iters) {
		auto *c = static_cast<ref_ctx*>(p);
		auto [pr, res, it] = c->solver->solve(dt);
		c->pressure = std::move(pr);
		c->residual = res;
		c->iters = it;
		*residual = res;
		*iters = it;
	}
	// pieces of solve(), for the RHS / matrix-flag parity checks (src/pressure_solver.cpp:150-242)
	void ref_solver_rhs(void *p, double dt, double *out_b, unsigned char *out_flags) {
		auto *c = static_cast<ref_ctx*>(p);
		fluid::pressure_solver &s = *c->solver;
		s._compute_fluid_cell_indices();
		s._a_scale = dt / (c->sim.density * c->sim.cell_size * c->sim.cell_size);
		s._compute_a_matrix();
		std::vector<double> b = s._compute_b_vector();
		std::copy(b.begin(), b.end(), out_b);
		for (std::size_t i = 0; i < s._a.size(); ++i) {
			out_flags[i] = static_cast<unsigned char>(
				s._a[i].nonsolid_neighbors | (s._a[i].fluid_xpos << 3) | (s._a[i].fluid_ypos << 4) |
				(s._a[i].fluid_zpos << 5)
			);
		}
	}
	// out = A * v with the reference's _apply_a (src/pressure_solver.cpp:334-362); ref_solver_rhs must have run
	void ref_solver_apply_a(void *p, const double *v, double *out) {
		auto *c = static_cast<ref_ctx*>(p);
		std::size_t n = c->fluid_cells.size();
		std::vector<double> vin(v, v + n), vout(n, 0.0);
		c->solver->_apply_a(vout, vin);
		std::copy(vout.begin(), vout.end(), out);
	}
	void ref_get_pressure(void *p, double *out) {
		auto *c = static_cast<ref_ctx*>(p);
		std::copy(c->pressure.begin(), c->pressure.end(), out);
	}
	void ref_set_pressure(void *p, const double *in, std::size_t n) {
		static_cast<ref_ctx*>(p)->pressure.assign(in, in + n);
	}
	void ref_apply_pressure(void *p, double dt) {
		auto *c = static_cast<ref_ctx*>(p);
		c->solver->apply_pressure(dt, c->pressure);
	}
	void ref_correct(void *p, double dt) {
		static_cast<ref_ctx*>(p)->sim._correct_positions(dt);
	}
	void ref_extrapolate(void *p) {
		auto *c = static_cast<ref_ctx*>(p);
		c->sim._extrapolate_velocities(c->fluid_cells);
	}
	void ref_g2p(void *p) {
		static_cast<ref_ctx*>(p)->sim._transfer_from_grid();
	}
	double ref_cfl(void *p) {
		return static_cast<ref_ctx*>(p)->sim.cfl();
	}
	// the stock, un-replayed entry points
	void ref_time_step(void *p, double dt) {
		static_cast<ref_ctx*>(p)->sim.time_step(dt);
	}
	void ref_time_step_default(void *p) {
		static_cast<ref_ctx*>(p)->sim.time_step();
	}
	void ref_update(void *p, double dt) {
		static_cast<ref_ctx*>(p)->sim.update(dt);
	}

	// ---- "next" rows of SURVEY.md 8(f): mesher surface sampling (N2) and obstacle voxelisation (N4) ----
	// mesher::_sample_surface_function (src/mesher.cpp:333-376) on a grid of size^3 cells; out: (sx+1)(sy+1)(sz+1) doubles
	void ref_mesher_sample(
		const std::size_t *size, const double *offset, double cell_size, double extent, std::size_t cell_radius,
		const double *xyz, std::size_t n, double r, double *out
	) {
		fluid::mesher m;
		m.grid_offset = vec3d(offset[0], offset[1], offset[2]);
		m.cell_size = cell_size;
		m.particle_extent = extent;
		m.cell_radius = cell_radius;
		m.resize(vec3s(size[0], size[1], size[2]));
		std::vector<vec3d> pts(n);
		for (std::size_t i = 0; i < n; ++i) {
			pts[i] = vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
		}
		m._sample_surface_function(pts, r);
		vec3s gs = m._surface_function.get_size();
		std::size_t k = 0;
		for (std::size_t z = 0; z < gs.z; ++z) {
			for (std::size_t y = 0; y < gs.y; ++y) {
				for (std::size_t x = 0; x < gs.x; ++x) {
					out[k++] = m._surface_function(x, y, z);
				}
			}
		}
	}
	// voxelizer (src/voxelizer.cpp:19-126) + obstacle (src/data_structures/obstacle.cpp:9-29) for a triangle mesh.
	// Call once with voxels == nullptr to get the sizes, then again with buffers.
	void ref_voxelize(
		const double *pos, std::size_t nverts, const std::size_t *idx, std::size_t nidx, double cell_size,
		const double *ref_offset, const std::size_t *ref_size, long long *grid_min, std::size_t *vox_size,
		unsigned char *voxels, std::size_t *ncells, std::size_t *cells_xyz
	) {
		fluid::obstacle::mesh_t mesh;
		for (std::size_t i = 0; i < nverts; ++i) {
			mesh.positions.emplace_back(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		}
		mesh.indices.assign(idx, idx + nidx);
		auto [rmin, rmax] = fluid::voxelizer::get_bounding_box(mesh.positions.begin(), mesh.positions.end());
		fluid::voxelizer vox;
		fluid::vec3i off = vox.resize_reposition_grid_constrained(
			rmin, rmax, cell_size, vec3d(ref_offset[0], ref_offset[1], ref_offset[2])
		);
		vox.voxelize_mesh_surface(mesh);
		vox.mark_exterior();
		vec3s vs = vox.voxels.get_size();
		for (int d = 0; d < 3; ++d) {
			grid_min[d] = off[d];
			vox_size[d] = vs[d];
		}
		if (voxels) {
			std::size_t k = 0;
			for (std::size_t z = 0; z < vs.z; ++z) {
				for (std::size_t y = 0; y < vs.y; ++y) {
					for (std::size_t x = 0; x < vs.x; ++x) {
						voxels[k++] = static_cast<unsigned char>(vox.voxels(x, y, z));
					}
				}
			}
		}
		// obstacle::obstacle walks voxels over get_overlapping_cell_range(), whose upper corner is computed in REFERENCE-grid
		// coordinates (src/voxelizer.cpp:45-49) but used as a VOXEL-grid bound (obstacle.cpp:21-28): with a positive
		// offset on any axis the reference reads beyond its voxel array (undefined behaviour; it crashed here).  The
		// driver only runs it where it is defined, and reports SIZE_MAX otherwise.
		if (off.x > 0 || off.y > 0 || off.z > 0) {
			*ncells = static_cast<std::size_t>(-1);
			return;
		}
		fluid::obstacle obs(mesh, cell_size, vec3d(ref_offset[0], ref_offset[1], ref_offset[2]),
			vec3s(ref_size[0], ref_size[1], ref_size[2]));
		*ncells = obs.cells.size();
		if (cells_xyz) {
			for (std::size_t i = 0; i < obs.cells.size(); ++i) {
				cells_xyz[3 * i] = obs.cells[i].x;
				cells_xyz[3 * i + 1] = obs.cells[i].y;
				cells_xyz[3 * i + 2] = obs.cells[i].z;
			}
		}
	}
}
